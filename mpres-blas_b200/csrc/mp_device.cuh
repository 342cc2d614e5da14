// mp_device.cuh -- residue-parallel device arithmetic on multiple-precision numbers.
//
// A multiple-precision number lives in a GROUP of G lanes (G = 8, 16 or 32, a power of two inside
// one warp), each lane owning R consecutive residues (element index e = lane_in_group * R + r).
// Sign, exponent and the interval evaluation are replicated in every lane of the group, so every
// control decision is group-uniform; the O(N) sums of the reference's scalar code become butterfly
// shuffles.  The reference keeps a whole number in ONE thread (src/blas/gemm.cuh:39-58), which
// spills for N >= 32.
//
// Arithmetic definitions follow the reference's cuda:: functions so that results are bit-identical,
// including the interval evaluations:
//   er_*            src/extrange.cuh:412-677          mp_mul   src/arith/mul.cuh:53-111
//   eval            src/rns.cuh:797-933               mp_add   src/arith/add.cuh:126-200
//   scale2pow       src/rns.cuh:1061-1160             mp_round src/arith/arith_utils.cuh:172-207
//   mrc / mrd       src/rns.cuh:570-630
// The pairwise FP64 sums of src/pairwise.cuh:855-931 are balanced binary trees over the zero-padded
// array; a xor-butterfly over lanes evaluates the same tree (IEEE addition is commutative and
// x + 0 == x in every rounding mode), so the directed-rounding results are the same bits.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace mpres {

constexpr int kMaxN = 128;
constexpr int kThresh = 30;  // RNS_P2_SCALING_THRESHOLD

struct SmallDev;

struct Er {
    double frac;
    long long exp;
};

// Device-resident constants of one moduli set (one copy per device of a context).
struct DevConsts {
    int N, log2M, mp_h, mp_j, ref_factor, precision;   // precision = MP_PRECISION (src/arith/arith_utils.cuh:44-61)
    double accuracy;
    Er unit_low, unit_upp, inv_low, inv_upp;
    int moduli[kMaxN];
    int part_inverse[kMaxN];
    unsigned long long barrett[kMaxN];  // floor((2^64-1)/m)
    double recip_rd[kMaxN], recip_ru[kMaxN];
    int m_pow2[kThresh];
    int pad1[2];
    int mi_pow2[kThresh][kMaxN];
    int pow2_inv[kThresh][kMaxN];
    const int *pow2;      // [log2M+1][N]   2^j mod m_i
    const int *inv_pow2;  // [log2M+1][N]   2^-j mod m_i
    const int *mrc_inv;   // [N][N]         m_i^-1 mod m_j (j > i)
    const int *prefix_mod;  // [N+1][N]     (m_0 ... m_{i-1}) mod m_q
    int prefix_log2[kMaxN + 1];   // floor(log2(m_0 ... m_{n-1}))
    const int *ext_w;     // [N][N]       CRT base extension from the first c moduli: (M'/m_i)^-1 mod m_i
    const int *ext_t;     // [N][N][N]    [c][q][i] = (M'/m_i) mod m_q
    int ext_lazy, pad2;
    const int *wpow2;     // [log2M+1][N] w_i * 2^j mod m_i
    const int *spow2;     // [2 (log2M+1) + 1][N]  +2^s, -2^s interleaved; last row zeros
    const struct SmallDev *small;   // tables of the small-modulus stage 2 (kernels_small.cuh); nullptr if unusable
};

// SoA view of mp_array_t / mp_collection_t (src/types.cuh:85-104).  `len` is the ALLOCATED length:
// upper bounds start at eval[len] (src/arith/mul.cuh:100).
struct SoA {
    int *digits;
    int *sign;
    int *exp;
    Er *eval;
    const int *len_ptr;  // device scalar of mp_array_t, or nullptr
    long long len_val;   // explicit length for mp_collection_t
    __device__ __forceinline__ long long len() const { return len_ptr ? (long long) *len_ptr : len_val; }
};

// ---- modular arithmetic -------------------------------------------------------------------------

// (a * b) mod m for 0 <= a, b < 2^31, canonical result in [0, m); mu = floor((2^64-1)/m).
// Same value as the reference's exact 64-bit % (src/modular.cuh:150-154) without the ~70-instruction
// software remainder.
__device__ __forceinline__ int mulmod(int a, int b, int m, unsigned long long mu) {
    unsigned long long p = (unsigned long long) (unsigned) a * (unsigned) b;
    unsigned long long q = __umul64hi(p, mu);
    unsigned r = (unsigned) (p - q * (unsigned long long) (unsigned) m);
    return (int) (r >= (unsigned) m ? r - (unsigned) m : r);
}
__device__ __forceinline__ int reduce64(unsigned long long p, int m, unsigned long long mu) {
    unsigned long long q = __umul64hi(p, mu);
    unsigned long long r = p - q * (unsigned long long) (unsigned) m;  // < 2m
    return (int) (r >= (unsigned long long) (unsigned) m ? r - (unsigned) m : r);
}
// v mod m for v < 2^(2 kb) when m has bit length kb (24 <= kb <= 27) and mu2 = floor(2^(kb + 30) / m):
// q' = floor((v >> (kb - 2)) mu2 / 2^32) misses floor(v / m) by at most one (the truncated bits cost < 1/2, the
// floor of mu2 < 1/8), so a single conditional subtraction yields the canonical residue.  Six instructions
// (funnel shift, IMAD.HI, IMAD, IADD, VIMNMX) against ~16 for the generic 64-bit step.
__device__ __forceinline__ unsigned barrett_k(unsigned long long v, unsigned m, unsigned mu2, int kb) {
    const unsigned ph = (unsigned) (v >> (kb - 2));
    const unsigned r = (unsigned) v - __umulhi(ph, mu2) * m;   // in [0, 2 m)
    return min(r, r - m);
}
__device__ __forceinline__ int submod(int a, int b, int m) {  // a, b in [0, m)
    int t = a - b;
    return t < 0 ? t + m : t;
}
// 2^j mod m for j outside the table (the reference reads out of bounds there)
static __device__ __noinline__ int pow2_slow(long long j, int m, unsigned long long mu) {
    int r = 1, b = 2 % m;
    while (j > 0) {
        if (j & 1) r = mulmod(r, b, m, mu);
        b = mulmod(b, b, m, mu);
        j >>= 1;
    }
    return r;
}

// ---- per-lane constants --------------------------------------------------------------------------

template <int R>
struct Lane {
    int m[R];
    unsigned long long mu[R];
    int w[R];
    double rrd[R], rru[R];
    int idx[R];  // residue index, clamped to N-1 for padding lanes
    bool act[R];
};

template <int G, int R>
__device__ __forceinline__ void lane_init(const DevConsts &C, Lane<R> &L) {
    const int gl = threadIdx.x & (G - 1);
#pragma unroll
    for (int r = 0; r < R; ++r) {
        int e = gl * R + r;
        L.act[r] = e < C.N;
        int ee = L.act[r] ? e : 0;
        L.idx[r] = ee;
        L.m[r] = L.act[r] ? C.moduli[ee] : 1;
        L.mu[r] = L.act[r] ? C.barrett[ee] : ~0ull;
        L.w[r] = L.act[r] ? C.part_inverse[ee] : 0;
        L.rrd[r] = L.act[r] ? C.recip_rd[ee] : 0.0;
        L.rru[r] = L.act[r] ? C.recip_ru[ee] : 0.0;
    }
}

template <int G>
__device__ __forceinline__ unsigned group_mask() {
    if constexpr (G == 32) {
        return 0xffffffffu;
    } else {
        const unsigned lane = threadIdx.x & 31u;
        return ((1u << G) - 1u) << (lane & ~(unsigned) (G - 1));
    }
}

// ---- group reductions ----------------------------------------------------------------------------

template <int G, int R, bool UP>
__device__ __forceinline__ double gsum_dir(const double (&x)[R]) {
    double t[R];
#pragma unroll
    for (int r = 0; r < R; ++r) t[r] = x[r];
#pragma unroll
    for (int w = 1; w < R; w <<= 1)
#pragma unroll
        for (int r = 0; r + w < R; r += 2 * w) t[r] = UP ? __dadd_ru(t[r], t[r + w]) : __dadd_rd(t[r], t[r + w]);
    double v = t[0];
    const unsigned mask = group_mask<G>();
#pragma unroll
    for (int o = 1; o < G; o <<= 1) {
        double u = __shfl_xor_sync(mask, v, o, G);
        v = UP ? __dadd_ru(v, u) : __dadd_rd(v, u);
    }
    return v;
}

template <int G>
__device__ __forceinline__ long long gsum_ll(long long v) {
    const unsigned mask = group_mask<G>();
#pragma unroll
    for (int o = 1; o < G; o <<= 1) v += __shfl_xor_sync(mask, v, o, G);
    return v;
}
template <int G>
__device__ __forceinline__ int gor(int v) {
    const unsigned mask = group_mask<G>();
#pragma unroll
    for (int o = 1; o < G; o <<= 1) v |= __shfl_xor_sync(mask, v, o, G);
    return v;
}

// ---- extended-range floats -----------------------------------------------------------------------

__device__ __forceinline__ void er_adjust(Er &x) {
    unsigned long long u = (unsigned long long) __double_as_longlong(x.frac);
    const bool nz = x.frac != 0;
    x.exp = nz ? x.exp + (long long) ((u & 0x7fffffffffffffffull) >> 52) - 1023 : 0;
    u = (u & 0x800fffffffffffffull) | 0x3ff0000000000000ull;
    // a zero keeps the sign bit, as (+-1.f * 0) does in the reference (extrange.cuh:420)
    x.frac = nz ? __longlong_as_double((long long) u) : __longlong_as_double((long long) (u & 0x8000000000000000ull));
}
__device__ __forceinline__ Er er_from_double(double x) {
    Er r;
    if (x != 0) {
        unsigned long long u = (unsigned long long) __double_as_longlong(x);
        r.exp = (long long) ((u & 0x7fffffffffffffffull) >> 52) - 1023;
        u = (u & 0x800fffffffffffffull) | 0x3ff0000000000000ull;
        r.frac = __longlong_as_double((long long) u);
    } else {
        r.exp = 0;
        r.frac = 0;
    }
    return r;
}
template <bool UP>
__device__ __forceinline__ Er er_add_dir(Er x, Er y) {
    if (x.frac == 0) x.exp = 0;
    if (y.frac == 0) y.exp = 0;
    const long long dexp = (x.frac != 0 && y.frac != 0) ? x.exp - y.exp : 0;
    Er r;
    if (dexp > 0) {
        r.exp = x.exp;
        double t = scalbn(y.frac, (int) -dexp);
        r.frac = UP ? __dadd_ru(x.frac, t) : __dadd_rd(x.frac, t);
    } else if (dexp < 0) {
        r.exp = y.exp;
        double t = scalbn(x.frac, (int) dexp);
        r.frac = UP ? __dadd_ru(y.frac, t) : __dadd_rd(y.frac, t);
    } else {
        r.exp = (x.exp == 0) ? y.exp : x.exp;
        r.frac = UP ? __dadd_ru(x.frac, y.frac) : __dadd_rd(x.frac, y.frac);
    }
    er_adjust(r);
    return r;
}
template <bool UP>
__device__ __forceinline__ Er er_md_dir(const Er &x, const Er &y, const Er &z) {
    Er r;
    r.exp = x.exp + y.exp - z.exp;
    r.frac = UP ? __ddiv_ru(__dmul_ru(x.frac, y.frac), z.frac) : __ddiv_rd(__dmul_rd(x.frac, y.frac), z.frac);
    er_adjust(r);
    return r;
}

// ---- a multiple-precision number spread over a group ---------------------------------------------

template <int R>
struct Num {
    int d[R];
    int sign;
    int exp;
    Er lo, up;
};

template <int R>
__device__ __forceinline__ void num_zero(Num<R> &x) {  // MP_ZERO, arith_utils.cuh:64-72
#pragma unroll
    for (int r = 0; r < R; ++r) x.d[r] = 0;
    x.sign = 0;
    x.exp = 0;
    x.lo.frac = 0; x.lo.exp = 0;
    x.up.frac = 0; x.up.exp = 0;
}

template <int G, int R>
__device__ __forceinline__ void load_digits(const DevConsts &C, const Lane<R> &L, const int *digits, long long idx, int (&d)[R]) {
    const int *p = digits + idx * C.N;
#pragma unroll
    for (int r = 0; r < R; ++r) d[r] = L.act[r] ? __ldg(p + L.idx[r]) : 0;
}
template <int G, int R>
__device__ __forceinline__ void store_digits(const DevConsts &C, const Lane<R> &L, int *digits, long long idx, const int (&d)[R]) {
    int *p = digits + idx * C.N;
#pragma unroll
    for (int r = 0; r < R; ++r) if (L.act[r]) p[L.idx[r]] = d[r];
}
template <int G, int R>
__device__ __forceinline__ void load_num(const DevConsts &C, const Lane<R> &L, const SoA &a, long long idx, Num<R> &x) {
    load_digits<G, R>(C, L, a.digits, idx, x.d);
    x.sign = a.sign[idx];
    x.exp = a.exp[idx];
    x.lo = a.eval[idx];
    x.up = a.eval[idx + a.len()];
}
template <int G, int R>
__device__ __forceinline__ void store_num(const DevConsts &C, const Lane<R> &L, const SoA &a, long long idx, const Num<R> &x) {
    store_digits<G, R>(C, L, a.digits, idx, x.d);
    if ((threadIdx.x & (G - 1)) == 0) {
        a.sign[idx] = x.sign;
        a.exp[idx] = x.exp;
        a.eval[idx] = x.lo;
        a.eval[idx + a.len()] = x.up;
    }
}
// AoS mp_float_t record (src/types.cuh:69-74): int digits[N]; int sign; int exp; er_float_t eval[2]
template <int G, int R>
__device__ __forceinline__ void load_rec(const DevConsts &C, const Lane<R> &L, const char *base, long long idx, Num<R> &x) {
    const char *p = base + idx * (4ll * C.N + 40);
    const int *pi = (const int *) p;
#pragma unroll
    for (int r = 0; r < R; ++r) x.d[r] = L.act[r] ? pi[L.idx[r]] : 0;
    x.sign = pi[C.N];
    x.exp = pi[C.N + 1];
    const Er *pe = (const Er *) (p + 4 * C.N + 8);
    x.lo = pe[0];
    x.up = pe[1];
}
template <int G, int R>
__device__ __forceinline__ void store_rec(const DevConsts &C, const Lane<R> &L, char *base, long long idx, const Num<R> &x) {
    char *p = base + idx * (4ll * C.N + 40);
    int *pi = (int *) p;
#pragma unroll
    for (int r = 0; r < R; ++r) if (L.act[r]) pi[L.idx[r]] = x.d[r];
    if ((threadIdx.x & (G - 1)) == 0) {
        pi[C.N] = x.sign;
        pi[C.N + 1] = x.exp;
        Er *pe = (Er *) (p + 4 * C.N + 8);
        pe[0] = x.lo;
        pe[1] = x.up;
    }
}

// ---- mixed-radix conversion: most significant digit and comparison -------------------------------

// Mixed-radix digits of x, lane-parallel: after step j every element e > j holds
// (x_e - a_0 - a_1 m_0 - ...)/(m_0..m_j) mod m_e.  (src/rns.cuh:570-582; the reference's negative
// intermediate when a_j >= m_e, a 1e-7 event, is replaced by the canonical residue.)
template <int G, int R>
__device__ __forceinline__ void mrc(const DevConsts &C, const Lane<R> &L, const int (&x)[R], int (&mr)[R]) {
    const unsigned mask = group_mask<G>();
    const int gl = threadIdx.x & (G - 1);
#pragma unroll
    for (int r = 0; r < R; ++r) mr[r] = x[r];
    for (int j = 0; j < C.N - 1; ++j) {
        int mj = 0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            int v = __shfl_sync(mask, mr[r], j / R, G);
            if (r == j % R) mj = v;
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            int e = gl * R + r;
            if (L.act[r] && e > j) {
                int m = L.m[r];
                int mjr = mj >= m ? mj - m : mj;  // moduli of one set differ by far less than a factor 2
                if (mjr >= m) mjr %= m;
                int t = submod(mr[r], mjr, m);
                mr[r] = mulmod(t, __ldg(C.mrc_inv + (long long) j * C.N + e), m, L.mu[r]);
            }
        }
    }
}
template <int G, int R>
__device__ __forceinline__ int mrd(const DevConsts &C, const Lane<R> &L, const int (&x)[R]) {
    int mr[R];
    mrc<G, R>(C, L, x, mr);
    const unsigned mask = group_mask<G>();
    int out = 0;
    const int last = C.N - 1;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        int v = __shfl_sync(mask, mr[r], last / R, G);
        if (r == last % R) out = v;
    }
    return out;
}
// 1 if X > Y, -1 if X < Y, 0 if equal (src/rns.cuh:608-630)
template <int G, int R>
__device__ __forceinline__ int mrc_compare(const DevConsts &C, const Lane<R> &L, const int (&x)[R], const int (&y)[R]) {
    int mx[R], my[R];
    mrc<G, R>(C, L, x, mx);
    mrc<G, R>(C, L, y, my);
    // lexicographic from the most significant digit: encode per-element comparison, pick the highest
    // differing element
    const int gl = threadIdx.x & (G - 1);
    long long key = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        int e = gl * R + r;
        if (L.act[r] && mx[r] != my[r]) {
            long long k = ((long long) (e + 1) << 2) | (mx[r] > my[r] ? 1 : 2);
            key = k > key ? k : key;
        }
    }
    const unsigned mask = group_mask<G>();
#pragma unroll
    for (int o = 1; o < G; o <<= 1) {
        long long u = __shfl_xor_sync(mask, key, o, G);
        key = u > key ? u : key;
    }
    if (key == 0) return 0;
    return (key & 3) == 1 ? 1 : -1;
}

// ---- interval evaluation --------------------------------------------------------------------------

__device__ __forceinline__ int pow2_at(const DevConsts &C, long long j, int e, int m, unsigned long long mu) {
    if (j >= 0 && j <= C.log2M) return __ldg(C.pow2 + j * C.N + e);
    return pow2_slow(j, m, mu);
}

// FAST = rns_eval_compute_fast (src/rns.cuh:878-933), otherwise rns_eval_compute (:797-868)
template <int G, int R, bool FAST>
__device__ __forceinline__ void eval_compute(const DevConsts &C, const Lane<R> &L, const int (&x)[R], Er &low, Er &upp) {
    int s[R];
    double fl[R], fu[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        s[r] = mulmod(x[r], L.w[r], L.m[r], L.mu[r]);
        fl[r] = __dmul_rd((double) s[r], L.rrd[r]);
        fu[r] = __dmul_ru((double) s[r], L.rru[r]);
    }
    double suml = gsum_dir<G, R, false>(fl);
    double sumu = gsum_dir<G, R, true>(fu);
    if (suml == 0 && sumu == 0) {
        low.frac = 0; low.exp = 0; upp.frac = 0; upp.exp = 0;
        return;
    }
    const unsigned whl = (unsigned) suml, whu = (unsigned) sumu;
    suml = __dsub_rd(suml, (double) whl);
    sumu = __dsub_ru(sumu, (double) whu);
    if (!FAST) {
        low = er_from_double(suml);
        upp = er_from_double(sumu);
        int mr = -1;
        if (whl != whu) mr = mrd<G, R>(C, L, x);
        if (mr > 0) { upp = C.inv_upp; return; }
        if (mr == 0) low = C.unit_low;
        if (sumu >= C.accuracy) return;
    } else if (sumu >= C.accuracy) {
        low = er_from_double(suml);
        upp = er_from_double(sumu);
        return;
    }
    int K = 0;
    while (sumu < C.accuracy) {
        double kd = -(ceil(log2(sumu)) + 1);
        int k = (int) (kd > (double) C.ref_factor ? kd : (double) C.ref_factor);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            s[r] = mulmod(s[r], L.act[r] ? pow2_at(C, k, L.idx[r], L.m[r], L.mu[r]) : 0, L.m[r], L.mu[r]);
            fu[r] = __dmul_ru((double) s[r], L.rru[r]);
        }
        sumu = gsum_dir<G, R, true>(fu);
        sumu = __dsub_ru(sumu, (double) (unsigned) sumu);
        K += k;
    }
#pragma unroll
    for (int r = 0; r < R; ++r) fl[r] = __dmul_rd((double) s[r], L.rrd[r]);
    suml = gsum_dir<G, R, false>(fl);
    suml = __dsub_rd(suml, (double) (unsigned) suml);
    low = er_from_double(suml);
    upp = er_from_double(sumu);
    low.exp -= K;
    upp.exp -= K;
}

// ---- power-of-two scaling (floor(X / 2^D)) ---------------------------------------------------------

template <int G, int R>
__device__ __forceinline__ int rank_full(const DevConsts &C, const Lane<R> &L, const int (&x)[R], const int (&s)[R]) {
    double fl[R], fu[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        fl[r] = __dmul_rd((double) s[r], L.rrd[r]);
        fu[r] = __dmul_ru((double) s[r], L.rru[r]);
    }
    const unsigned whl = (unsigned) gsum_dir<G, R, false>(fl), whu = (unsigned) gsum_dir<G, R, true>(fu);
    if (whl == whu) return (int) whl;
    return mrd<G, R>(C, L, x) == 0 ? (int) whu : (int) whl;
}
template <int G, int R>
__device__ __forceinline__ int rank_fast(const Lane<R> &L, const int (&s)[R]) {
    double fu[R];
#pragma unroll
    for (int r = 0; r < R; ++r) fu[r] = __dmul_ru((double) s[r], L.rru[r]);
    return (int) gsum_dir<G, R, true>(fu);
}
// one step: y = (x - (X mod 2^j)) * (2^j)^-1   (src/rns.cuh:1104-1124)
template <int G, int R>
__device__ __forceinline__ void scaling_step(const DevConsts &C, const Lane<R> &L, int (&x)[R], int k, int j, const int (&c)[R]) {
    const long long pow2j = 1ll << j;
    long long part = 0;
#pragma unroll
    for (int r = 0; r < R; ++r)
        if (L.act[r]) part += ((long long) C.mi_pow2[j - 1][L.idx[r]] * (long long) c[r]) & (pow2j - 1);
    long long residue = gsum_ll<G>(part);
    // (a mod 2^j) made non-negative == two's-complement masking
    residue = (residue - (long long) k * (long long) C.m_pow2[j - 1]) & (pow2j - 1);
#pragma unroll
    for (int r = 0; r < R; ++r) {
        int mult = reduce64((unsigned long long) residue, L.m[r], L.mu[r]);
        mult = submod(x[r], mult, L.m[r]);
        x[r] = L.act[r] ? mulmod(mult, C.pow2_inv[j - 1][L.idx[r]], L.m[r], L.mu[r]) : 0;
    }
}
// src/rns.cuh:1132-1160 (device rank rule: the cheap rank only when the single step is the remainder step)
template <int G, int R>
__device__ __forceinline__ void scale2pow(const DevConsts &C, const Lane<R> &L, int (&x)[R], unsigned D) {
    int t = (int) (D / kThresh);
    int c[R];
    bool first = true;
    while (t > 0) {
#pragma unroll
        for (int r = 0; r < R; ++r) c[r] = mulmod(x[r], L.w[r], L.m[r], L.mu[r]);
        int k = first ? rank_full<G, R>(C, L, x, c) : rank_fast<G, R>(L, c);
        scaling_step<G, R>(C, L, x, k, kThresh, c);
        first = false;
        --t;
    }
    const unsigned d = D % kThresh;
    if (d > 0) {
#pragma unroll
        for (int r = 0; r < R; ++r) c[r] = mulmod(x[r], L.w[r], L.m[r], L.mu[r]);
        int k = d < D ? rank_full<G, R>(C, L, x, c) : rank_fast<G, R>(L, c);
        scaling_step<G, R>(C, L, x, k, (int) d, c);
    }
}

// ---- multiple-precision operations -----------------------------------------------------------------

template <int G, int R>
__device__ __forceinline__ void mp_round(const DevConsts &C, const Lane<R> &L, Num<R> &x, int n) {
    if (n > 0) {
        x.exp += n;
        scale2pow<G, R>(C, L, x.d, (unsigned) n);
        eval_compute<G, R, true>(C, L, x.d, x.lo, x.up);
    }
}
template <int G, int R>
__device__ __forceinline__ void round_if_needed(const DevConsts &C, const Lane<R> &L, Num<R> &x) {
    if (x.up.frac != 0 && x.up.exp >= C.mp_h) mp_round<G, R>(C, L, x, (int) (x.up.exp - C.mp_h + 1));
}

template <int G, int R, bool ROUND>
__device__ __forceinline__ void mp_mul(const DevConsts &C, const Lane<R> &L, Num<R> &r, const Num<R> &x, const Num<R> &y) {
    Num<R> t;
    t.exp = x.exp + y.exp;
    t.sign = x.sign ^ y.sign;
    t.lo = er_md_dir<false>(x.lo, y.lo, C.unit_upp);
    t.up = er_md_dir<true>(x.up, y.up, C.unit_low);
#pragma unroll
    for (int q = 0; q < R; ++q) t.d[q] = mulmod(x.d[q], y.d[q], L.m[q], L.mu[q]);
    r = t;
    if (ROUND) round_if_needed<G, R>(C, L, r);
}

// Exponent/sign/interval part of cuda::mp_add (src/arith/add.cuh:126-170), shared by the residue-parallel
// mp_add below and by the entry-per-thread normalisation kernel: alignment shifts, zeroing flags,
// masked signs and the interval of the signed sum (before the result sign is applied).
struct AddEsi {
    int gamma, theta, nzx, nzy, sx, sy, ex, ey;
    Er lo, up;
};
__device__ __forceinline__ AddEsi add_esi(const DevConsts &C, Er xl, Er xu, Er yl, Er yu, int ex, int ey, int sx, int sy) {
    AddEsi p;
    const int dexp = ex - ey;
    int gamma = dexp > 0 ? dexp : 0;
    int theta = dexp < 0 ? -dexp : 0;
    const int nzx = (yu.frac == 0) || ((long long) theta + yu.exp) < C.mp_j;
    const int nzy = (xu.frac == 0) || ((long long) gamma + xu.exp) < C.mp_j;
    gamma *= nzy;
    theta *= nzx;
    ex = (ex - gamma) * nzx;
    ey = (ey - theta) * nzy;
    sx *= nzx;
    sy *= nzy;
    const int fx = (1 - 2 * sx) * nzx, fy = (1 - 2 * sy) * nzy;
    xl.exp += gamma; xu.exp += gamma; yl.exp += theta; yu.exp += theta;
    xl.frac *= fx; xu.frac *= fx; yl.frac *= fy; yu.frac *= fy;
    p.lo = er_add_dir<false>(sx ? xu : xl, sy ? yu : yl);
    p.up = er_add_dir<true>(sx ? xl : xu, sy ? yl : yu);
    p.gamma = gamma; p.theta = theta; p.nzx = nzx; p.nzy = nzy; p.sx = sx; p.sy = sy; p.ex = ex; p.ey = ey;
    return p;
}

// STYLE 0: scalar cuda::mp_add (sign resolved by mixed-radix comparison when the interval straddles
// zero, add.cuh:165-170).  ROUND adds the trailing rounding of add.cuh:197-199.
template <int G, int R, bool ROUND>
__device__ __forceinline__ void mp_add(const DevConsts &C, const Lane<R> &L, Num<R> &res, const Num<R> &xin, const Num<R> &yin) {
    const AddEsi p = add_esi(C, xin.lo, xin.up, yin.lo, yin.up, xin.exp, yin.exp, xin.sign, yin.sign);
    const int gamma = p.gamma, theta = p.theta, nzx = p.nzx, nzy = p.nzy, sx = p.sx, sy = p.sy, ex = p.ex, ey = p.ey;
    Num<R> t;
    t.lo = p.lo;
    t.up = p.up;
    // shifted operands (needed for the digits and, rarely, for the sign)
    int ax[R], ay[R];
#pragma unroll
    for (int q = 0; q < R; ++q) {
        int pg = (L.act[q] && nzx) ? pow2_at(C, gamma, L.idx[q], L.m[q], L.mu[q]) : 0;
        int pt = (L.act[q] && nzy) ? pow2_at(C, theta, L.idx[q], L.m[q], L.mu[q]) : 0;
        ax[q] = mulmod(xin.d[q], pg, L.m[q], L.mu[q]);
        ay[q] = mulmod(yin.d[q], pt, L.m[q], L.mu[q]);
    }
    int sign = t.lo.frac < 0;
    if (sign != (t.up.frac < 0)) {
        int cmp = mrc_compare<G, R>(C, L, ax, ay);
        sign = (cmp < 0 ? sy : sx) * (cmp != 0);
        if (sign) { t.up.frac = -C.unit_low.frac; t.up.exp = C.unit_low.exp; }
        else { t.lo.frac = C.unit_low.frac; t.lo.exp = C.unit_low.exp; }
    }
    t.sign = sign;
    t.exp = (ex == 0) ? ey : ex;
#pragma unroll
    for (int q = 0; q < R; ++q) {
        // (fx * ax + fy * ay) mod m, canonical; negated when the result is negative
        int a = sx ? (ax[q] ? L.m[q] - ax[q] : 0) : ax[q];
        int b = sy ? (ay[q] ? L.m[q] - ay[q] : 0) : ay[q];
        int v = a + b - L.m[q];
        v = v < 0 ? v + L.m[q] : v;
        t.d[q] = sign ? (v ? L.m[q] - v : 0) : v;
        if (!L.act[q]) t.d[q] = 0;
    }
    if (sign) {
        Er tmp = t.lo;
        t.lo.frac = -t.up.frac; t.lo.exp = t.up.exp;
        t.up.frac = -tmp.frac; t.up.exp = tmp.exp;
    }
    res = t;
    if (ROUND) round_if_needed<G, R>(C, L, res);
}

}  // namespace mpres
