/*
 * oracle/mpres_oracle.c -- TEST INFRASTRUCTURE ONLY (see mpres_oracle.h).
 *
 * Plain-C restatement of the reference algorithm for the mp_gemm / mp_gemv / mp_dot path.
 * Every function cites the reference lines (under /root/reference/src) it follows.  Written from the
 * algorithm's definition for a run-time moduli count; it is validated bit-for-bit against the
 * reference itself (oracle/_ref) by tests/test_oracle_vs_ref.py and against committed golden vectors.
 *
 * Build: gcc -O2 -std=c11 -frounding-math -ffp-contract=off -fopenmp (see Makefile).  The directed
 * rounding of the DEVICE flavour relies on fesetround, hence -frounding-math and the volatile temps.
 */
#define _GNU_SOURCE
#include "mpres_oracle.h"

#include <fenv.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <omp.h>


struct orc_ctx {
    int N, log2M, flavor, mp_h, mp_j, ref_factor;
    double accuracy;
    orc_er_t unit_low, unit_upp, inv_low, inv_upp;
    int *moduli, *part_inverse, *pow2, *m_pow2, *mi_pow2, *pow2_inv, *mrc_inv;
    double *recip_rd, *recip_ru;
};

static void *dup_mem(const void *p, size_t n) { void *q = malloc(n); memcpy(q, p, n); return q; }

orc_ctx *orc_create(int N, int log2M, int flavor, int mp_h, int mp_j, int ref_factor, double accuracy,
                    const orc_er_t *unit_low, const orc_er_t *unit_upp, const orc_er_t *inv_low,
                    const orc_er_t *inv_upp, const int *moduli, const int *part_inverse, const int *pow2,
                    const int *m_pow2, const int *mi_pow2, const int *pow2_inv, const int *mrc_inv,
                    const double *recip_rd, const double *recip_ru) {
    if (N < 1 || N > ORC_MAX_N) return NULL;
    orc_ctx *c = calloc(1, sizeof(*c));
    c->N = N; c->log2M = log2M; c->flavor = flavor; c->mp_h = mp_h; c->mp_j = mp_j;
    c->ref_factor = ref_factor; c->accuracy = accuracy;
    c->unit_low = *unit_low; c->unit_upp = *unit_upp; c->inv_low = *inv_low; c->inv_upp = *inv_upp;
    c->moduli = dup_mem(moduli, sizeof(int) * N);
    c->part_inverse = dup_mem(part_inverse, sizeof(int) * N);
    c->pow2 = dup_mem(pow2, sizeof(int) * N * (log2M + 1));
    c->m_pow2 = dup_mem(m_pow2, sizeof(int) * 30);
    c->mi_pow2 = dup_mem(mi_pow2, sizeof(int) * 30 * N);
    c->pow2_inv = dup_mem(pow2_inv, sizeof(int) * 30 * N);
    c->mrc_inv = dup_mem(mrc_inv, sizeof(int) * N * N);
    c->recip_rd = dup_mem(recip_rd, sizeof(double) * N);
    c->recip_ru = dup_mem(recip_ru, sizeof(double) * N);
    return c;
}

void orc_destroy(orc_ctx *c) {
    if (!c) return;
    free(c->moduli); free(c->part_inverse); free(c->pow2); free(c->m_pow2); free(c->mi_pow2);
    free(c->pow2_inv); free(c->mrc_inv); free(c->recip_rd); free(c->recip_ru); free(c);
}

int orc_record_size(const orc_ctx *c) { return 4 * c->N + 40; }

/* record field access (types.cuh:69-74) */
#define DIG(p) ((int *) (p))
#define SIGN(c, p) (((int *) (p))[(c)->N])
#define EXPO(c, p) (((int *) (p))[(c)->N + 1])
#define EVAL(c, p) ((orc_er_t *) ((char *) (p) + 4 * (c)->N + 8))
#define REC(c, base, i) ((void *) ((char *) (base) + (size_t) (i) * (4 * (c)->N + 40)))
#define CREC(c, base, i) ((const void *) ((const char *) (base) + (size_t) (i) * (4 * (c)->N + 40)))

/* ---------------- directed double arithmetic ------------------------------------------------------
 * HOST: dinterval.cuh:55-165 (c -/+ (phi*|c| + eta), no FMA since EMPLOY_STD_FMA is false);
 * DEVICE: the IEEE-754 directed intrinsics __dadd_rd ... __ddiv_ru. */
static const double PHI1 = 0x1.0000000000001p-53;  /* bitwise.cuh: DBL_EPS*(1+2*DBL_EPS) */
static const double ETA = 0x1p-1074;

static inline double widen(double c, int up) {
    volatile double e = PHI1 * fabs(c);
    e = e + ETA;
    return up ? c + e : c - e;
}

static inline double d_op(const orc_ctx *c, int op, double a, double b, int up) {
    volatile double va = a, vb = b, r;
    if (c->flavor == ORC_HOST) {
        switch (op) { case 0: r = va + vb; break; case 1: r = va - vb; break; case 2: r = va * vb; break; default: r = va / vb; }
        return widen(r, up);
    }
    fesetround(up ? FE_UPWARD : FE_DOWNWARD);
    switch (op) { case 0: r = va + vb; break; case 1: r = va - vb; break; case 2: r = va * vb; break; default: r = va / vb; }
    fesetround(FE_TONEAREST);
    return r;
}
#define D_ADD(c, a, b, up) d_op(c, 0, a, b, up)
#define D_SUB(c, a, b, up) d_op(c, 1, a, b, up)
#define D_MUL(c, a, b, up) d_op(c, 2, a, b, up)
#define D_DIV(c, a, b, up) d_op(c, 3, a, b, up)

/* pairwise.cuh:141-475 (host) / 598-931 (device): psumP<LEN> = psumP/2<LEN>(x) (+) (LEN > P/2 ?
 * psumP/2<LEN-P/2>(x+P/2) : 0); the top-level P is chosen from LEN (psum_rd/ru<LENGTH>). */
static double psum_rec(const orc_ctx *c, const double *x, int len, int P, int up) {
    if (P == 2) return D_ADD(c, x[0], len > 1 ? x[1] : 0.0, up);
    double s1 = psum_rec(c, x, len, P / 2, up);
    double s2 = len > P / 2 ? psum_rec(c, x + P / 2, len - P / 2, P / 2, up) : 0.0;
    return D_ADD(c, s1, s2, up);
}
static double psum(const orc_ctx *c, const double *x, int len, int up) {
    if (len <= 0) return 0;
    if (len == 1) return x[0];
    if (len == 2) return D_ADD(c, x[0], x[1], up);
    if (len == 3) return D_ADD(c, D_ADD(c, x[0], x[1], up), x[2], up);
    int P = 8;
    while (len >= P) P *= 2;   /* 4..7 -> psum8, 8..15 -> psum16, ... */
    return psum_rec(c, x, len, P, up);
}

/* ---------------- extended-range floats (extrange.cuh) -------------------------------------------- */
typedef union { double d; uint64_t i; } du_t;

static void er_adjust(orc_er_t *x) { /* extrange.cuh:134-143 / 412-421 */
    du_t u; u.d = x->frac;
    x->exp = (x->exp + (long) ((u.i & ~((uint64_t) 1 << 63)) >> 52) - 1023) * (x->frac != 0);
    u.i = (u.i & ((((uint64_t) 1 << 52) - 1) | ((uint64_t) 1 << 63))) | ((uint64_t) 1023 << 52);
    x->frac = u.d * (x->frac != 0);
}
static void er_set_d(orc_er_t *r, double x) { /* extrange.cuh:46-59 */
    du_t u; u.d = x;
    if (x != 0) {
        r->exp = (long) ((u.i & ~((uint64_t) 1 << 63)) >> 52) - 1023;
        u.i = (u.i & ((((uint64_t) 1 << 52) - 1) | ((uint64_t) 1 << 63))) | ((uint64_t) 1023 << 52);
        r->frac = u.d;
    } else { r->exp = 0; r->frac = 0; }
}
static double fast_scalbn(double x, long n) { /* bitwise.cuh:92-97 */
    du_t u; u.d = x;
    u.i += (uint64_t) n << 52;
    return u.d * (x != 0 && n >= -1023);
}
static double scale(const orc_ctx *c, double x, long n) {
    /* host er_add uses fast_scalbn (extrange.cuh:181-185), device uses scalbn (:458-461) */
    return c->flavor == ORC_HOST ? fast_scalbn(x, n) : scalbn(x, (int) n);
}
static orc_er_t er_add_dir(const orc_ctx *c, orc_er_t x, orc_er_t y, int up) { /* extrange.cuh:173-212 / 451-490 */
    x.exp *= (x.frac != 0);
    y.exp *= (y.frac != 0);
    long dexp = (x.exp - y.exp) * (x.frac != 0) * (y.frac != 0);
    orc_er_t r;
    if (dexp > 0) { r.exp = x.exp; r.frac = D_ADD(c, x.frac, scale(c, y.frac, -dexp), up); }
    else if (dexp < 0) { r.exp = y.exp; r.frac = D_ADD(c, y.frac, scale(c, x.frac, dexp), up); }
    else { r.exp = (x.exp == 0) ? y.exp : x.exp; r.frac = D_ADD(c, x.frac, y.frac, up); }
    er_adjust(&r);
    return r;
}
static orc_er_t er_md_dir(const orc_ctx *c, orc_er_t x, orc_er_t y, orc_er_t z, int up) { /* extrange.cuh:349-366 / 627-644 */
    orc_er_t r;
    r.exp = x.exp + y.exp - z.exp;
    r.frac = D_DIV(c, D_MUL(c, x.frac, y.frac, up), z.frac, up);
    er_adjust(&r);
    return r;
}

/* ---------------- modular arithmetic (modular.cuh) ------------------------------------------------ */
static inline int mod_mul_exact(int x, int y, int m) { return (int) (((long) x * (long) y) % (long) m); } /* :81-85, :150-154 */
static inline int mod_mul_recip(int x, int y, int m) { /* :90-95, d = 1.0/m as in rns_mul :218-222 */
    long r = (long) x * (long) y;
    double q = (double) r * (1.0 / m);
    int i = (int) q;
    return (int) (r - (long) i * (long) m);
}
static inline int mod_psub(int x, int y, int m) { return (int) (((long) x - (long) y + (long) m) % (long) m); } /* :70-74 */
/* rns_mul: host = reciprocal trick (:218-222), device = exact % (:250-256) */
static inline int rns_mul1(const orc_ctx *c, int x, int y, int m) {
    return c->flavor == ORC_HOST ? mod_mul_recip(x, y, m) : mod_mul_exact(x, y, m);
}
static int pow2_mod(const orc_ctx *c, long j, int i) {
    /* RNS_POW2[j][i] (rns.cuh:352-358); beyond the table (reference reads out of bounds)
     * we return the mathematically intended 2^j mod m_i */
    if (j >= 0 && j <= c->log2M) return c->pow2[j * c->N + i];
    long r = 1, b = 2, m = c->moduli[i];
    for (long e = j; e > 0; e >>= 1) { if (e & 1) r = r * b % m; b = b * b % m; }
    return (int) r;
}

/* ---------------- mixed-radix conversion (rns.cuh:495-556 / 570-630) ------------------------------- */
static void mrc(const orc_ctx *c, int *mr, const int *x) {
    for (int i = 0; i < c->N; i++) {
        mr[i] = x[i];
        for (int j = 0; j < i; j++) {
            if (mr[i] < mr[j]) mr[i] = (int) ((long) c->moduli[i] - (long) mr[j] + (long) mr[i]);
            else mr[i] = mr[i] - mr[j];
            mr[i] = mod_mul_exact(mr[i], c->mrc_inv[j * c->N + i], c->moduli[i]);
        }
    }
}
int orc_mrc_compare(const orc_ctx *c, const int *x, const int *y) {
    int mx[ORC_MAX_N], my[ORC_MAX_N];
    mrc(c, mx, x); mrc(c, my, y);
    for (int i = c->N - 1; i >= 0; i--) {
        if (mx[i] > my[i]) return 1;
        if (my[i] > mx[i]) return -1;
    }
    return 0;
}
static int mrd(const orc_ctx *c, const int *x) { int mr[ORC_MAX_N]; mrc(c, mr, x); return mr[c->N - 1]; }

/* ---------------- interval evaluation (rns.cuh:645-781 host, 797-933 device) ----------------------- */
static void eval_impl(const orc_ctx *c, orc_er_t *low, orc_er_t *upp, const int *x, int fast) {
    const int N = c->N, host = c->flavor == ORC_HOST;
    int s[ORC_MAX_N];
    double fl[ORC_MAX_N], fu[ORC_MAX_N];
    if (host) { /* :653-657 zero test on the digits */
        int z = 1;
        for (int i = 0; i < N; i++) if (x[i]) { z = 0; break; }
        if (z) { low->frac = 0; low->exp = 0; upp->frac = 0; upp->exp = 0; return; }
    }
    for (int i = 0; i < N; i++) {
        s[i] = rns_mul1(c, x[i], c->part_inverse[i], c->moduli[i]);
        fl[i] = D_MUL(c, (double) s[i], c->recip_rd[i], 0);
        fu[i] = D_MUL(c, (double) s[i], c->recip_ru[i], 1);
    }
    double suml = psum(c, fl, N, 0), sumu = psum(c, fu, N, 1);
    if (!host && suml == 0 && sumu == 0) { /* :813-817 */
        low->frac = 0; low->exp = 0; upp->frac = 0; upp->exp = 0; return;
    }
    unsigned int whl = (unsigned int) suml, whu = (unsigned int) sumu;
    suml = suml - whl;   /* exact in either flavour */
    sumu = sumu - whu;
    if (!fast) {
        er_set_d(low, suml); er_set_d(upp, sumu);
        int mr = -1;
        if (whl != whu) mr = mrd(c, x);
        if (mr > 0) { *upp = c->inv_upp; return; }
        if (mr == 0) *low = c->unit_low;
        if (sumu >= c->accuracy) return;
    } else if (sumu >= c->accuracy) {
        er_set_d(low, suml); er_set_d(upp, sumu);
        return;
    }
    int K = 0;
    while (sumu < c->accuracy) { /* :695-706 / :845-856 */
        double kd = -(ceil(log2(sumu)) + 1);
        int k = (int) (kd > c->ref_factor ? kd : (double) c->ref_factor);
        for (int i = 0; i < N; i++) {
            s[i] = rns_mul1(c, s[i], pow2_mod(c, k, i), c->moduli[i]);
            fu[i] = host ? D_DIV(c, (double) s[i], (double) c->moduli[i], 1) : D_MUL(c, (double) s[i], c->recip_ru[i], 1);
        }
        sumu = psum(c, fu, N, 1);
        sumu -= (unsigned int) sumu;
        K += k;
    }
    for (int i = 0; i < N; i++)
        fl[i] = host ? D_DIV(c, (double) s[i], (double) c->moduli[i], 0) : D_MUL(c, (double) s[i], c->recip_rd[i], 0);
    suml = psum(c, fl, N, 0);
    suml -= (unsigned int) suml;
    er_set_d(low, suml); er_set_d(upp, sumu);
    low->exp -= K; upp->exp -= K;
}
void orc_eval_compute(const orc_ctx *c, orc_er_t *low, orc_er_t *upp, const int *d) { eval_impl(c, low, upp, d, 0); }
void orc_eval_compute_fast(const orc_ctx *c, orc_er_t *low, orc_er_t *upp, const int *d) { eval_impl(c, low, upp, d, 1); }

/* ---------------- power-of-two scaling (rns.cuh:946-1048 host, 1061-1160 device) ------------------- */
static int rank_full(const orc_ctx *c, const int *x, const int *s) { /* :946-975 / :1061-1083 */
    double fl[ORC_MAX_N], fu[ORC_MAX_N];
    for (int i = 0; i < c->N; i++) {
        fl[i] = D_MUL(c, (double) s[i], c->recip_rd[i], 0);
        fu[i] = D_MUL(c, (double) s[i], c->recip_ru[i], 1);
    }
    unsigned int whl = (unsigned int) psum(c, fl, c->N, 0), whu = (unsigned int) psum(c, fu, c->N, 1);
    if (whl == whu) return (int) whl;
    return mrd(c, x) == 0 ? (int) whu : (int) whl;
}
static int rank_fast(const orc_ctx *c, const int *s) { /* :983-992 / :1091-1099 */
    double fu[ORC_MAX_N];
    for (int i = 0; i < c->N; i++) fu[i] = D_MUL(c, (double) s[i], c->recip_ru[i], 1);
    return (int) psum(c, fu, c->N, 1);
}
static void scaling_step(const orc_ctx *c, int *y, int k, unsigned int j, int pow2j, const int *x, const int *cc) { /* :997-1014 / :1104-1124 */
    long residue = 0;
    for (int i = 0; i < c->N; i++) residue += (long) mod_mul_exact(c->mi_pow2[(j - 1) * c->N + i], cc[i], pow2j);
    long temp = (long) k * (long) c->m_pow2[j - 1];
    residue = (residue - temp) % pow2j;
    if (residue < 0) residue += pow2j;
    for (int i = 0; i < c->N; i++) {
        int multiple = (int) (residue % c->moduli[i]);
        multiple = mod_psub(x[i], multiple, c->moduli[i]);
        y[i] = rns_mul1(c, multiple, c->pow2_inv[(j - 1) * c->N + i], c->moduli[i]);
    }
}
void orc_scale2pow(const orc_ctx *c, int *result, const int *x, unsigned int D) { /* :1022-1048 / :1132-1160 */
    const int N = c->N;
    int cur[ORC_MAX_N], cc[ORC_MAX_N], nxt[ORC_MAX_N];
    memcpy(cur, x, sizeof(int) * N);
    int t = D / 30, k;
    if (t > 0) {
        for (int i = 0; i < N; i++) cc[i] = rns_mul1(c, cur[i], c->part_inverse[i], c->moduli[i]);
        k = rank_full(c, cur, cc);
        scaling_step(c, nxt, k, 30, 1 << 30, cur, cc);
        memcpy(cur, nxt, sizeof(int) * N);
        t -= 1;
    }
    while (t > 0) {
        for (int i = 0; i < N; i++) cc[i] = rns_mul1(c, cur[i], c->part_inverse[i], c->moduli[i]);
        k = rank_fast(c, cc);
        scaling_step(c, nxt, k, 30, 1 << 30, cur, cc);
        memcpy(cur, nxt, sizeof(int) * N);
        t -= 1;
    }
    unsigned int d = D % 30;
    if (d > 0) {
        for (int i = 0; i < N; i++) cc[i] = rns_mul1(c, cur[i], c->part_inverse[i], c->moduli[i]);
        if (c->flavor == ORC_HOST) k = d < D ? rank_fast(c, cc) : rank_full(c, cur, cc);        /* :1045 */
        else k = d < D ? rank_full(c, cur, cc) : rank_fast(c, cc);                               /* :1157 */
        scaling_step(c, nxt, k, d, 1 << d, cur, cc);
        memcpy(cur, nxt, sizeof(int) * N);
    }
    memcpy(result, cur, sizeof(int) * N);
}

/* ---------------- scalar multiple-precision arithmetic -------------------------------------------- */
void orc_mp_round(const orc_ctx *c, void *x, int n) { /* arith_utils.cuh:138-146 / 184-193 */
    if (n > 0) {
        EXPO(c, x) += n;
        orc_scale2pow(c, DIG(x), DIG(x), (unsigned) n);
        orc_eval_compute_fast(c, &EVAL(c, x)[0], &EVAL(c, x)[1], DIG(x));
    }
}
static void round_if_needed(const orc_ctx *c, void *r) { /* mul.cuh:39-41, add.cuh:111-113 */
    orc_er_t *e = EVAL(c, r);
    if (e[1].frac != 0 && e[1].exp >= c->mp_h) orc_mp_round(c, r, (int) (e[1].exp - c->mp_h + 1));
}

static void mul_noround(const orc_ctx *c, void *r, const void *x, const void *y) { /* mul.cuh:31-38 / 53-61 */
    char tmp[4 * ORC_MAX_N + 40];
    const orc_er_t *ex = EVAL(c, x), *ey = EVAL(c, y);
    EXPO(c, tmp) = EXPO(c, x) + EXPO(c, y);
    SIGN(c, tmp) = SIGN(c, x) ^ SIGN(c, y);
    EVAL(c, tmp)[0] = er_md_dir(c, ex[0], ey[0], c->unit_upp, 0);
    EVAL(c, tmp)[1] = er_md_dir(c, ex[1], ey[1], c->unit_low, 1);
    for (int i = 0; i < c->N; i++) DIG(tmp)[i] = mod_mul_exact(DIG(x)[i], DIG(y)[i], c->moduli[i]);
    memcpy(r, tmp, orc_record_size(c));
}
void orc_mp_mul(const orc_ctx *c, void *r, const void *x, const void *y) {
    mul_noround(c, r, x, y);
    round_if_needed(c, r);
}

static int sign_estimate(const orc_ctx *c, const int *dx, const int *dy, int sx, int sy, int gamma, int theta, int nzx, int nzy) { /* arith_utils.cuh:152-162 / 199-207 */
    int lx[ORC_MAX_N], ly[ORC_MAX_N];
    for (int i = 0; i < c->N; i++) {
        lx[i] = nzx * mod_mul_exact(dx[i], pow2_mod(c, gamma, i), c->moduli[i]);
        ly[i] = nzy * mod_mul_exact(dy[i], pow2_mod(c, theta, i), c->moduli[i]);
    }
    int cmp = orc_mrc_compare(c, lx, ly);
    return (cmp < 0 ? sy : sx) * (cmp != 0);
}

/* style 0: scalar mp_add (add.cuh:31-114 host / 126-200 device).
 * style 1: the matrix/vector add kernels of the v1 BLAS (mpmatrix.cuh:47-171): no MRC for an
 *          ambiguous sign (q4) and the reciprocal mod_axby in the digits kernel (q11). */
static void add_noround(const orc_ctx *c, void *r, const void *xin, const void *yin, int style) {
    const int N = c->N;
    char xb[4 * ORC_MAX_N + 40], yb[4 * ORC_MAX_N + 40], tmp[4 * ORC_MAX_N + 40];
    memcpy(xb, xin, orc_record_size(c)); memcpy(yb, yin, orc_record_size(c));
    orc_er_t *evx = EVAL(c, xb), *evy = EVAL(c, yb), *evr = EVAL(c, tmp);
    int ex = EXPO(c, xb), ey = EXPO(c, yb), sx = SIGN(c, xb), sy = SIGN(c, yb);
    int dexp = ex - ey;
    int gamma = dexp * (dexp > 0);
    int theta = -dexp * (dexp < 0);
    int nzx = ((evy[1].frac == 0) || (theta + evy[1].exp) < c->mp_j);
    int nzy = ((evx[1].frac == 0) || (gamma + evx[1].exp) < c->mp_j);
    gamma = gamma * nzy;
    theta = theta * nzx;
    ex = (ex - gamma) * nzx;
    ey = (ey - theta) * nzy;
    sx *= nzx; sy *= nzy;
    int fx = (1 - 2 * sx) * nzx, fy = (1 - 2 * sy) * nzy;
    evx[0].exp += gamma; evx[1].exp += gamma; evy[0].exp += theta; evy[1].exp += theta;
    evx[0].frac *= fx; evx[1].frac *= fx; evy[0].frac *= fy; evy[1].frac *= fy;
    evr[0] = er_add_dir(c, evx[sx], evy[sy], 0);
    evr[1] = er_add_dir(c, evx[1 - sx], evy[1 - sy], 1);
    int sr;
    if (style == 1) {
        sr = evr[0].frac < 0 && evr[1].frac < 0;   /* mpmatrix.cuh:110-111 */
    } else if (c->flavor == ORC_HOST) {
        if (evr[0].frac * evr[1].frac >= 0) sr = (evr[0].frac < 0);
        else {
            sr = sign_estimate(c, DIG(xb), DIG(yb), sx, sy, gamma, theta, nzx, nzy);
            evr[sr].frac = c->unit_low.frac * (1 - 2 * sr);
            evr[sr].exp = c->unit_low.exp;
        }
    } else {
        sr = evr[0].frac < 0;
        if (sr != (evr[1].frac < 0)) {
            sr = sign_estimate(c, DIG(xb), DIG(yb), sx, sy, gamma, theta, nzx, nzy);
            evr[sr].frac = c->unit_low.frac * (1 - 2 * sr);
            evr[sr].exp = c->unit_low.exp;
        }
    }
    SIGN(c, tmp) = sr;
    EXPO(c, tmp) = (ex == 0) ? ey : ex;
    for (int i = 0; i < N; i++) {
        int m = c->moduli[i];
        long a = (long) DIG(xb)[i] * fx, b = (long) DIG(yb)[i] * fy;
        long acc = a * (long) pow2_mod(c, gamma, i) + b * (long) pow2_mod(c, theta, i);
        int res;
        if (style == 1) { /* cuda::mod_axby(..., m, 1.0/m) modular.cuh:163-168, then mpmatrix.cuh:160-164 */
            double q = (double) acc * (1.0 / m);
            int qi = (int) q;
            res = (int) (acc - (long) qi * (long) m);
            if (sr == 1) res = (int) (((long) m - (long) res) % (long) m);
            res = res < 0 ? res + m : res;
        } else {
            res = (int) (acc % (long) m);
            if (res < 0) res += m;
            if (sr == 1) res = (res != 0) * (m - res);   /* host (m - d) % m gives the same value */
        }
        DIG(tmp)[i] = res;
    }
    if (sr == 1) {
        orc_er_t t = evr[0];
        evr[0].frac = -evr[1].frac; evr[0].exp = evr[1].exp;
        evr[1].frac = -1 * t.frac; evr[1].exp = t.exp;
    }
    memcpy(r, tmp, orc_record_size(c));
}
void orc_mp_add(const orc_ctx *c, void *r, const void *x, const void *y) {
    add_noround(c, r, x, y, 0);
    round_if_needed(c, r);
}

/* mp_set_mpfr (assign.cuh:86-127) for a significand given in binary limbs */
void orc_mp_set(const orc_ctx *c, void *r, int sign, const uint32_t *limbs, int nlimbs, int exp) {
    uint32_t w[160];
    if (nlimbs > 160) nlimbs = 160;
    memcpy(w, limbs, 4 * nlimbs);
    int top = nlimbs;
    while (top > 0 && w[top - 1] == 0) top--;
    if (top == 0) { /* zero: mpfr_get_str gives "0...0" -> digits 0; sign 0 */
        memset(r, 0, orc_record_size(c));
        /* the reference's exponent for an exact zero comes out of mpfr_get_str (exp 0 - length);
         * callers of the oracle never rely on it */
        return;
    }
    int tz = 0;
    while (((w[tz / 32] >> (tz % 32)) & 1u) == 0) tz++;
    /* shift right by tz */
    int ws = tz / 32, bs = tz % 32;
    for (int i = 0; i < top; i++) {
        uint64_t lo = (i + ws < top) ? w[i + ws] : 0, hi = (i + ws + 1 < top) ? w[i + ws + 1] : 0;
        w[i] = bs ? (uint32_t) ((lo >> bs) | (hi << (32 - bs))) : (uint32_t) lo;
    }
    for (int i = 0; i < c->N; i++) {
        uint64_t m = (uint64_t) c->moduli[i], acc = 0;
        for (int l = top - 1; l >= 0; l--) acc = ((acc << 32) | w[l]) % m;
        DIG(r)[i] = (int) acc;
    }
    SIGN(c, r) = sign ? 1 : 0;
    EXPO(c, r) = exp + tz;
    orc_eval_compute(c, &EVAL(c, r)[0], &EVAL(c, r)[1], DIG(r));
}

static uint64_t splitmix64(uint64_t *st) {
    uint64_t z = (*st += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
void orc_random_fill(const orc_ctx *c, void *recs, long n, int bits, uint64_t seed) {
    int nl = (bits + 31) / 32;
    if (nl > 150) nl = 150;
    #pragma omp parallel for schedule(static)
    for (long i = 0; i < n; i++) {
        uint64_t st = seed * 0x100000001b3ull + (uint64_t) i * 0x9e3779b97f4a7c15ull + 1;
        double z = (double) (splitmix64(&st) >> 11) * 0x1p-53, u = (double) (splitmix64(&st) >> 11) * 0x1p-52 - 1.0;
        double v = z * fabs(u);
        if (v < 0x1p-200) v = 0x1p-200;
        int e;
        double mant = frexp(v, &e);
        uint64_t top = (uint64_t) (mant * 0x1p53);
        uint32_t w[160];
        for (int j = 0; j < nl; j++) w[j] = (uint32_t) splitmix64(&st);
        int sh = bits - 53;
        if (sh >= 0) {
            for (int j = 0; j < nl; j++) {
                int lo_bit = 32 * j;
                uint32_t keep = lo_bit + 32 <= sh ? 0xffffffffu : (lo_bit < sh ? ((1u << (sh - lo_bit)) - 1u) : 0u);
                int s = sh - lo_bit;
                uint32_t hi = s >= 32 ? 0u : (s >= 0 ? (uint32_t) (top << s) : (-s < 64 ? (uint32_t) (top >> (-s)) : 0u));
                w[j] = hi | (w[j] & keep);
            }
        } else {
            uint64_t t = top >> (-sh);
            for (int j = 0; j < nl; j++) w[j] = j < 2 ? (uint32_t) (t >> (32 * j)) : 0u;
        }
        orc_mp_set(c, REC(c, recs, i), u < 0, w, nl, e - bits);
    }
}

/* ---------------- vector and BLAS-level restatements ---------------------------------------------- */
static void set_zero(const orc_ctx *c, void *r) { memset(r, 0, orc_record_size(c)); } /* MP_ZERO, arith_utils.cuh:64-72 */

void orc_mul_vec(const orc_ctx *c, void *r, const void *x, const void *y, long n) {
    #pragma omp parallel for schedule(static)
    for (long i = 0; i < n; i++) orc_mp_mul(c, REC(c, r, i), CREC(c, x, i), CREC(c, y, i));
}
void orc_add_vec(const orc_ctx *c, void *r, const void *x, const void *y, long n) {
    #pragma omp parallel for schedule(static)
    for (long i = 0; i < n; i++) orc_mp_add(c, REC(c, r, i), CREC(c, x, i), CREC(c, y, i));
}

void orc_dot_seq(const orc_ctx *c, void *r, const void *x, const void *y, long n) { /* SURVEY 3.5 */
    char acc[4 * ORC_MAX_N + 40], t[4 * ORC_MAX_N + 40];
    set_zero(c, acc);
    for (long i = 0; i < n; i++) {
        orc_mp_mul(c, t, CREC(c, x, i), CREC(c, y, i));
        orc_mp_add(c, acc, acc, t);
    }
    memcpy(r, acc, orc_record_size(c));
}

int orc_dot_omp(const orc_ctx *c, void *r, const void *x, const void *y, long n) {
    int nt = omp_get_max_threads();
    int rs = orc_record_size(c);
    char *part = calloc(nt, rs);
    #pragma omp parallel num_threads(nt)
    {
        int t = omp_get_thread_num();
        long lo = n * t / nt, hi = n * (t + 1) / nt;
        orc_dot_seq(c, part + (size_t) t * rs, CREC(c, x, lo), CREC(c, y, lo), hi - lo);
    }
    char acc[4 * ORC_MAX_N + 40];
    set_zero(c, acc);
    for (int t = 0; t < nt; t++) orc_mp_add(c, acc, acc, part + (size_t) t * rs);
    memcpy(r, acc, rs);
    free(part);
    return nt;
}

static unsigned next_pow2(unsigned x) { unsigned p = 1; while (p < x) p <<= 1; return p; } /* common.cuh:92-101 */

/* one launch of mp_array_reduce_sum_kernel1/2 (mpreduct.cuh:38-112): out[b] for b in [0, grid) */
static void reduce_pass(const orc_ctx *c, char *out, const char *in, long n, int grid, int block) {
    int rs = orc_record_size(c);
    #pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < grid; b++) {
        char *sd = calloc(block, rs);
        for (int t = 0; t < block; t++) {
            char *acc = sd + (size_t) t * rs;
            for (long i = (long) b * block + t; i < n; i += (long) grid * block) orc_mp_add(c, acc, acc, in + (size_t) i * rs);
        }
        for (unsigned i = next_pow2(block) >> 1; i >= 1; i >>= 1)
            for (unsigned t = 0; t < i && t + i < (unsigned) block; t++)
                orc_mp_add(c, sd + (size_t) t * rs, sd + (size_t) t * rs, sd + (size_t) (t + i) * rs);
        memcpy(out + (size_t) b * rs, sd, rs);
        free(sd);
    }
}

void orc_dot_v1(const orc_ctx *c, void *r, const void *x, const void *y, long n, int grid, int block) { /* dot.cuh:84-107 */
    int rs = orc_record_size(c);
    char *buf = malloc((size_t) n * rs), *part = malloc((size_t) grid * rs);
    orc_mul_vec(c, buf, x, y, n);          /* esi + digits + round kernels == mp_mul incl. rounding */
    reduce_pass(c, part, buf, n, grid, block);
    reduce_pass(c, r, part, grid, 1, block);
    free(buf); free(part);
}

int orc_gemm_rows(const orc_ctx *c, int row0, int row1, int m, int n, int k, const void *alpha, const void *A,
                  int lda, const void *B, int ldb, const void *beta, void *C, int ldc, void *AB_out) {
    int nt = omp_get_max_threads();
    int style = c->flavor == ORC_DEVICE ? 1 : 0;
    #pragma omp parallel for collapse(2) schedule(static)
    for (int j = 0; j < n; j++) {
        for (int i = row0; i < row1; i++) {
            char sum[4 * ORC_MAX_N + 40], mul[4 * ORC_MAX_N + 40], t2[4 * ORC_MAX_N + 40];
            set_zero(c, sum);
            for (int l = 0; l < k; l++) { /* gemm.cuh:46-49 */
                orc_mp_mul(c, mul, CREC(c, A, (size_t) lda * l + i), CREC(c, B, (size_t) ldb * j + l));
                orc_mp_add(c, sum, sum, mul);
            }
            if (AB_out) memcpy(REC(c, AB_out, (size_t) m * j + i), sum, orc_record_size(c));
            orc_mp_mul(c, sum, sum, alpha);                                   /* gemm.cuh:142-148 */
            void *cij = REC(c, C, (size_t) ldc * j + i);
            orc_mp_mul(c, t2, cij, beta);                                     /* gemm.cuh:151-157 */
            add_noround(c, cij, t2, sum, style);                              /* gemm.cuh:160-163 */
            round_if_needed(c, cij);                                          /* gemm.cuh:166 */
        }
    }
    return nt;
}

int orc_gemv(const orc_ctx *c, int trans, int m, int n, const void *alpha, const void *A, int lda,
             const void *x, const void *beta, void *y, int block) {
    int nt = omp_get_max_threads();
    int rs = orc_record_size(c);
    int lenx = trans == 111 ? n : m, leny = trans == 111 ? m : n;
    char *ax = malloc((size_t) lenx * rs);
    for (int j = 0; j < lenx; j++) orc_mp_mul(c, ax + (size_t) j * rs, CREC(c, x, j), alpha);  /* gemv.cuh:175-181 */
    #pragma omp parallel for schedule(static)
    for (int o = 0; o < leny; o++) {
        char *sd = calloc(block, rs), t[4 * ORC_MAX_N + 40];
        void *yo = REC(c, y, o);
        orc_mp_mul(c, yo, yo, beta);                                                          /* :184-190 */
        for (int tt = 0; tt < block; tt++) {                                                  /* gemv.cuh:44-76 / 88-120 */
            char *acc = sd + (size_t) tt * rs;
            for (int q = tt; q < lenx; q += block) {
                const void *a = trans == 111 ? CREC(c, A, (size_t) lda * q + o) : CREC(c, A, (size_t) lda * o + q);
                orc_mp_mul(c, t, a, ax + (size_t) q * rs);                                    /* :199-205 */
                orc_mp_add(c, acc, acc, t);
            }
        }
        for (unsigned i = next_pow2(block) >> 1; i >= 1; i >>= 1)
            for (unsigned tt = 0; tt < i && tt + i < (unsigned) block; tt++)
                orc_mp_add(c, sd + (size_t) tt * rs, sd + (size_t) tt * rs, sd + (size_t) (tt + i) * rs);
        orc_mp_add(c, yo, yo, sd);                                                            /* add.cuh:225-250 */
        free(sd);
    }
    free(ax);
    return nt;
}
