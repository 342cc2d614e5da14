// kernels_bin.cuh -- stage 3 of the fast mp_gemm path for sums that do NOT fit the number format: full-precision (p-bit) inputs.
//
// The reference rounds every product and every partial sum of its k-loop to the working precision (src/arith/mul.cuh:108-110,
// src/arith/add.cuh:197-199).  With p-bit significands the exact sums S(i,j) of stages 1-2 have about 2p + (exponent spread) +
// log2 k bits -- more than log2(M) - 2, so the window guard of kernels_norm.cuh fails for every entry.  The sums are still known
// EXACTLY through their residues modulo the one-byte base (|S| < M'/4, k_choose_base), so they are rebuilt in binary here and
// rounded ONCE:
//      xi_i = x_i (M'/p_i)^-1 mod p_i,   R = nearest integer of sum xi_i / p_i,   S = sum xi_i (M'/p_i) - R M'    (multiword, exact)
//      T = |S| rounded to nearest at MP_PRECISION bits (ties away from zero),  exponent = base + dropped bits
//      digits = T mod m_q (words of T against 2^(32 w) mod m_q),  interval evaluation of T / M from the leading 63 bits of T with
//      directed roundings (tighter than the reference's 1e-7 relative width, src/params.h:47)
// followed by the reference's epilogue C = round(round(beta C) + round(alpha T)) (src/blas/gemm.cuh:142-166) on the residue-parallel
// mp_mul / mp_add of mp_device.cuh.  One rounding of the exact sum instead of up to 2k roundings: the error is at most that of the
// reference's own model (tests/blas/accuracy/test_dot_accuracy.cu:41-72, u = 4 / sqrt(M)), the results are not its bits.
//
// One block = 128 consecutive rows of one column of C: phase 1 one THREAD per entry (binary reconstruction, rounding, digits into a
// shared-memory tile), phase 2 one LANE GROUP per entry (epilogue).
#pragma once

namespace mpres {

constexpr int kBinT = 128;            // entries per block

struct BinSmem {
    uint8_t *X8;          // [56][kBinT] one-byte residues of the tile, then the xi_i in place
    uint4 *mi4;           // [kSmallMax][kBinW / 4] words of M'/p_i
    unsigned *negmp;      // [kBinW]
    uint4 *c4;            // [64] (p, floor(2^32 / p), (M'/p)^-1 mod p, bits of 1 / p)
    unsigned *xw;         // [kBinW + 2][kBinT] words of |S| (two zero words on top)
    int *dig;             // [kBinT][N + 1] digits of the rounded sums
    int *sgn, *ex;        // [kBinT]
    Er *lo, *up;          // [kBinT]
};
__host__ __device__ inline size_t bin_smem_bytes(int N) {
    return (size_t) 56 * kBinT + (size_t) kSmallMax * kBinW * 4 + kBinW * 4 + 64 * 16 + (size_t) (kBinW + 2) * kBinT * 4 + (size_t) kBinT * (N + 1) * 4 +
           2 * kBinT * 4 + 2 * kBinT * 16 + 64;
}
__device__ __forceinline__ BinSmem bin_carve(uint8_t *base, int N) {
    BinSmem b;
    b.lo = (Er *) base; b.up = b.lo + kBinT;
    b.mi4 = (uint4 *) (b.up + kBinT);
    b.c4 = b.mi4 + kSmallMax * (kBinW / 4);
    b.negmp = (unsigned *) (b.c4 + 64);
    b.xw = b.negmp + kBinW;
    b.dig = (int *) (b.xw + (kBinW + 2) * kBinT);
    b.sgn = b.dig + kBinT * (N + 1);
    b.ex = b.sgn + kBinT;
    b.X8 = (uint8_t *) (b.ex + kBinT);
    return b;
}

// alpha / beta epilogue of one entry (src/blas/gemm.cuh:142-166), residue-parallel
template <int G, int R>
__device__ __forceinline__ void gemm_epilogue_entry(const DevConsts &C, const Lane<R> &L, const Num<R> &s, const Num<R> &al, const Num<R> &be, const SoA &Cm, long long ic) {
    Num<R> c, t1, t2;
    load_num<G, R>(C, L, Cm, ic, c);
    mp_mul<G, R, true>(C, L, t1, s, al);
    mp_mul<G, R, true>(C, L, t2, c, be);
    mp_add<G, R, true>(C, L, c, t2, t1);
    store_num<G, R>(C, L, Cm, ic, c);
}

// Interval evaluation of T / M from the binary T (words t[0 .. NW), bit length Lt > 0): the leading 63 bits as doubles rounded down /
// up, times 1 / M rounded down / up (DevConsts::unit_low, unit_upp).
template <int NW>
__device__ __forceinline__ void bin_eval(const DevConsts &C, const unsigned (&t)[NW], int Lt, Er &lo, Er &up) {
    const int tw = (Lt - 1) >> 5;                   // top word
    unsigned a = 0, b = 0, c = 0, below = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        a = w == tw ? t[w] : a;
        b = w == tw - 1 ? t[w] : b;
        c = w == tw - 2 ? t[w] : c;
        below |= w < tw - 2 ? t[w] : 0u;
    }
    // 96 bits (a, b, c); the leading one sits at bit ((Lt - 1) & 31) of a: shift left so that it lands on bit 95, keep the top 63
    const int lz = 31 - ((Lt - 1) & 31);
    unsigned long long hi = ((unsigned long long) a << 32) | b;
    unsigned long long mid = (unsigned long long) c << 32;
    if (lz) { hi = (hi << lz) | (mid >> (64 - lz)); mid <<= lz; }
    const unsigned long long top = hi >> 1;                                     // 63 bits, leading one at bit 62
    const bool sticky = (hi & 1ull) || mid != 0 || below != 0;
    const int e2 = Lt - 63;                                                       // T = (top + fraction) 2^e2 (e2 may be negative: then exact)
    const double dl = __ull2double_rd(top), du = __ull2double_ru(top + (sticky ? 1ull : 0ull));
    lo = er_from_double(dl); up = er_from_double(du);
    lo.exp += e2; up.exp += e2;
    // unit_low / unit_upp are the reference's 1 / M of the 53-bit TRUNCATED M (src/rns.cuh:336): up to 2^-52 above the true 1 / M, so both
    // bounds are moved out by 2^-50 (relative) to enclose T / M
    lo.frac = __dmul_rd(__dmul_rd(lo.frac, C.unit_low.frac), 1.0 - 8.8817841970012523e-16); lo.exp += C.unit_low.exp;
    up.frac = __dmul_ru(__dmul_ru(up.frac, C.unit_upp.frac), 1.0 + 8.8817841970012523e-16); up.exp += C.unit_upp.exp;
    er_adjust(lo); er_adjust(up);
}

template <int G, int R>
__global__ void __launch_bounds__(kBinT) k_bin_norm(const DevConsts *Cp, int m, int n, const uint8_t *S8, long long m_ps, long long n_ps, const int *sel,
                                                    const OuterInfo *ia, const OuterInfo *ib, SoA alpha, SoA beta, SoA Cm, int ldc, const int *gate) {
    extern __shared__ __align__(16) uint8_t bin_smem[];
    const int P = sel[0];
    if (P <= 0) return;
    if (gate && *gate != 0) return;                // the binary epilogue (k_bin_norm2) handled the call
    const DevConsts &C = *Cp;
    const SmallDev &SD = *C.small;
    const int N = C.N;
    const BinSmem B = bin_carve(bin_smem, N);
    Lane<R> L;
    lane_init<G, R>(C, L);
    // tables of this call's base
    for (int v = threadIdx.x; v < P * (kBinW / 4); v += kBinT) B.mi4[v] = __ldg((const uint4 *) (SD.bin_mi + (size_t) P * kSmallMax * kBinW) + v);
    if (threadIdx.x < kBinW) B.negmp[threadIdx.x] = SD.bin_negmp[P * kBinW + threadIdx.x];
    if (threadIdx.x < 64) {
        const int j = threadIdx.x;
        B.c4[j] = make_uint4((unsigned) SD.p[j], SD.mu[j], (unsigned) SD.inv[P * 64 + j], __float_as_uint(SD.rcp[j]));
    }
    B.xw[kBinW * kBinT + threadIdx.x] = 0u;
    B.xw[(kBinW + 1) * kBinT + threadIdx.x] = 0u;
    Num<R> al, be;
    load_num<G, R>(C, L, alpha, 0, al);
    load_num<G, R>(C, L, beta, 0, be);
    const int prec = C.precision;
    const int tiles = (m + kBinT - 1) / kBinT;
    const long long total = (long long) tiles * n;
    const long long plane = n_ps * m_ps;
    for (long long tl = blockIdx.x; tl < total; tl += gridDim.x) {
        const int col = (int) (tl / tiles);
        const int row0 = (int) (tl - (long long) col * tiles) * kBinT;
        __syncthreads();                                          // the previous tile is consumed (and the tables are in place)
        {
            const uint8_t *src = S8 + (long long) col * m_ps + row0;
            for (int v = threadIdx.x; v < P * (kBinT / 16); v += kBinT) {
                const int j = v >> 3, part = v & 7;
                cp_async16(B.X8 + j * kBinT + part * 16, src + (long long) j * plane + part * 16);
            }
            cp_async_commit();
            cp_async_wait<0>();
        }
        __syncthreads();
        // ---- phase 1: one thread per entry ----
        {
            const int row = row0 + threadIdx.x;
            int sign = 0, ex = 0, nzero = 1;
            Er lo, up;
            lo.frac = 0; lo.exp = 0; up.frac = 0; up.exp = 0;
            int *mydig = B.dig + threadIdx.x * (N + 1);
            OuterInfo ra = {}, cb = {};
            bool live = row < m;
            if (live) { ra = ia[row]; cb = ib[col]; live = ra.win >= 0 && cb.win >= 0; }
            if (live) {
                uint8_t *xs = B.X8 + threadIdx.x;
                float sum = 0.f;
                for (int j = 0; j < P; ++j) {
                    const uint4 cj = B.c4[j];
                    const unsigned tj = (unsigned) xs[j * kBinT] * cj.z;
                    const unsigned rj = tj - __umulhi(tj, cj.y) * cj.x;
                    const unsigned xi = min(rj, rj - cj.x);
                    sum = fmaf((float) xi, __uint_as_float(cj.w), sum);
                    xs[j * kBinT] = (uint8_t) xi;
                }
                const unsigned Rk = (unsigned) __float2int_rn(sum);
                unsigned x[kBinW];
                unsigned long long carry = 0;
#pragma unroll
                for (int w4 = 0; w4 < kBinW / 4; ++w4) {
                    unsigned long long c0 = (unsigned long long) Rk * B.negmp[4 * w4], c1 = (unsigned long long) Rk * B.negmp[4 * w4 + 1],
                                       c2 = (unsigned long long) Rk * B.negmp[4 * w4 + 2], c3 = (unsigned long long) Rk * B.negmp[4 * w4 + 3];
                    for (int j = 0; j < P; ++j) {
                        const unsigned long long xi = xs[j * kBinT];
                        const uint4 mm = B.mi4[j * (kBinW / 4) + w4];
                        c0 += xi * mm.x; c1 += xi * mm.y; c2 += xi * mm.z; c3 += xi * mm.w;
                    }
                    c0 += carry; x[4 * w4] = (unsigned) c0;
                    c1 += c0 >> 32; x[4 * w4 + 1] = (unsigned) c1;
                    c2 += c1 >> 32; x[4 * w4 + 2] = (unsigned) c2;
                    c3 += c2 >> 32; x[4 * w4 + 3] = (unsigned) c3;
                    carry = c3 >> 32;
                }
                // two's complement of the sum modulo 2^(32 kBinW): sign and magnitude
                sign = (int) (x[kBinW - 1] >> 31);
                if (sign) {
                    unsigned long long cy = 1;
#pragma unroll
                    for (int w = 0; w < kBinW; ++w) { cy += (unsigned long long) (~x[w]); x[w] = (unsigned) cy; cy >>= 32; }
                }
                int Lb = 0;
#pragma unroll
                for (int w = 0; w < kBinW; ++w) if (x[w]) Lb = 32 * w + 32 - __clz(x[w]);
                if (Lb > 0) {
                    nzero = 0;
                    int drop = Lb - prec;
                    if (drop > 0) {
                        // round to nearest: add half of the last dropped place, then cut
                        const int hw = (drop - 1) >> 5;
                        const unsigned hb = 1u << ((drop - 1) & 31);
                        unsigned long long cy = 0;
#pragma unroll
                        for (int w = 0; w < kBinW; ++w) { cy += (unsigned long long) x[w] + (w == hw ? hb : 0u); x[w] = (unsigned) cy; cy >>= 32; }
                    } else {
                        drop = 0;
                    }
#pragma unroll
                    for (int w = 0; w < kBinW; ++w) B.xw[w * kBinT + threadIdx.x] = x[w];
                    // T = x >> drop (at most prec + 1 bits)
                    const int ws = drop >> 5, bs = drop & 31;
                    unsigned t[kBinW];
                    int Lt = 0;
#pragma unroll
                    for (int w = 0; w < kBinW; ++w) {
                        unsigned v = 0;
                        if (w + ws < kBinW) {
                            const unsigned lo32 = B.xw[(w + ws) * kBinT + threadIdx.x], hi32 = B.xw[(w + ws + 1) * kBinT + threadIdx.x];
                            v = __funnelshift_r(lo32, hi32, bs);
                        }
                        t[w] = v;
                        if (v) Lt = 32 * w + 32 - __clz(v);
                    }
                    ex = ra.emin + cb.emin + drop;
                    bin_eval<kBinW>(C, t, Lt, lo, up);
                    // digits: sum_w t_w (2^(32 w) mod m_q); the words above the precision are zero
                    const int nwords = (Lt + 31) >> 5;
                    for (int q = 0; q < N; ++q) {
                        const int mq = C.moduli[q];
                        const unsigned long long muq = C.barrett[q];
                        unsigned long long acc = 0;
#pragma unroll
                        for (int w = 0; w < kBinW; ++w) {
                            if (w < nwords) {
                                acc += (unsigned long long) t[w] * (unsigned) __ldg(C.pow2 + (long long) (32 * w) * N + q);
                                if ((w & 7) == 7) acc = (unsigned long long) (unsigned) reduce64(acc, mq, muq);
                            }
                        }
                        mydig[q] = reduce64(acc, mq, muq);
                    }
                }
            }
            if (nzero) {
                sign = 0; ex = 0;
                for (int q = 0; q < N; ++q) mydig[q] = 0;
            }
            B.sgn[threadIdx.x] = sign; B.ex[threadIdx.x] = ex; B.lo[threadIdx.x] = lo; B.up[threadIdx.x] = up;
        }
        __syncthreads();
        // ---- phase 2: one lane group per entry: C = round(round(beta C) + round(alpha T)) ----
        for (int e = threadIdx.x / G; e < kBinT; e += kBinT / G) {
            const int row = row0 + e;
            if (row >= m) break;
            Num<R> s;
#pragma unroll
            for (int r = 0; r < R; ++r) s.d[r] = L.act[r] ? B.dig[e * (N + 1) + L.idx[r]] : 0;
            s.sign = B.sgn[e]; s.exp = B.ex[e]; s.lo = B.lo[e]; s.up = B.up[e];
            gemm_epilogue_entry<G, R>(C, L, s, al, be, Cm, row + (long long) col * ldc);
        }
    }
}

}  // namespace mpres

namespace mpres {

// ======================================================================================================================================
// Binary epilogue.  The residue-parallel mp_mul / mp_add of k_bin_norm spend ~2000 instructions per lane on every rounding
// (rns_scale2pow walks 30 bits per step, src/rns.cuh:1132-1160): three roundings per entry made the epilogue 85 % of the kernel.  With
// the sum already in binary the whole epilogue is cheaper in binary too, one THREAD per entry:
//      T  = rn(S)                          t1 = rn(alpha T)            t2 = rn(beta C)            C = rn(t1 + t2)
// rn = rounding to nearest at MP_PRECISION bits (ties away from zero) of the EXACT value; C reaches binary by the Chinese remainder
// theorem over all N moduli, alpha and beta once per call (k_scalar_binary).  The digits of the result are the words of its significand
// against 2^(32 w) mod m_q, its interval evaluation comes from its leading 63 bits.  Every step is exact integer arithmetic, so the kernel is
// pinned bit for bit by a Python-integer model (tests/test_gpu_fullprec.py); against the reference it is its error model with a smaller
// constant (correct rounding where rns_scale2pow truncates).
namespace mw {

template <int N>
__device__ __forceinline__ int bitlen(const unsigned (&x)[N]) {
    int L = 0;
#pragma unroll
    for (int w = 0; w < N; ++w) if (x[w]) L = 32 * w + 32 - __clz(x[w]);
    return L;
}
template <int N>
__device__ __forceinline__ void shr(unsigned (&x)[N], int sh) {        // logical, 0 <= sh < 32 N
    const int ws = sh >> 5, bs = sh & 31;
#pragma unroll
    for (int step = 1; step < N; step <<= 1)
        if (ws & step) {
#pragma unroll
            for (int w = 0; w < N; ++w) x[w] = (w + step < N) ? x[w + step] : 0u;
        }
#pragma unroll
    for (int w = 0; w < N; ++w) x[w] = __funnelshift_r(x[w], w + 1 < N ? x[w + 1] : 0u, bs);
}
template <int N>
__device__ __forceinline__ void shl(unsigned (&x)[N], int sh) {        // 0 <= sh < 32 N; bits shifted beyond the top are lost
    const int ws = sh >> 5, bs = sh & 31;
#pragma unroll
    for (int step = 1; step < N; step <<= 1)
        if (ws & step) {
#pragma unroll
            for (int w = N - 1; w >= 0; --w) x[w] = (w - step >= 0) ? x[w - step] : 0u;
        }
#pragma unroll
    for (int w = N - 1; w >= 0; --w) x[w] = __funnelshift_l(w > 0 ? x[w - 1] : 0u, x[w], bs);
}
// magnitude x of bit length L (< 32 N) rounded to at most p bits, nearest, ties away from zero; returns the bits dropped
template <int N>
__device__ __forceinline__ int round_to(unsigned (&x)[N], int L, int p) {
    const int drop = L - p;
    if (drop <= 0) return 0;
    const int hw = (drop - 1) >> 5;
    const unsigned hb = 1u << ((drop - 1) & 31);
    unsigned long long cy = 0;
#pragma unroll
    for (int w = 0; w < N; ++w) { cy += (unsigned long long) x[w] + (w == hw ? hb : 0u); x[w] = (unsigned) cy; cy >>= 32; }
    shr<N>(x, drop);
    return drop;
}
template <int NA, int NB>
__device__ __forceinline__ void mul(const unsigned (&a)[NA], const unsigned (&b)[NB], unsigned (&out)[NA + NB]) {
#pragma unroll
    for (int w = 0; w < NA + NB; ++w) out[w] = 0u;
#pragma unroll
    for (int i = 0; i < NA; ++i) {
        unsigned long long cy = 0;
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            const unsigned long long t = (unsigned long long) a[i] * b[j] + out[i + j] + cy;
            out[i + j] = (unsigned) t;
            cy = t >> 32;
        }
        out[i + NB] = (unsigned) cy;
    }
}
template <int N>
__device__ __forceinline__ int cmp(const unsigned (&a)[N], const unsigned (&b)[N]) {
    int r = 0;
#pragma unroll
    for (int w = 0; w < N; ++w) r = a[w] != b[w] ? (a[w] > b[w] ? 1 : -1) : r;
    return r;
}

}  // namespace mw

// alpha and beta of a call in binary: words [kScalW] | bit length | exponent | sign, twice; flag[0] = 1 when both fit PW words
constexpr int kScalW = 64;
struct ScalBin { unsigned w[kScalW]; int len, exp, sign, pad; };

// CRT over all N moduli: the significand of one number, one thread.  xw: out, nwf words (zero-extended).  Tables: fmi [N][nwf] words of M / m_i,
// fneg = 2^(32 nwf) - M, fmw = M (nwf = words of M + 1).
__device__ __forceinline__ void crt_full(const DevConsts &C, const int *dg, int dstride, const unsigned *fmi, const unsigned *fneg, const unsigned *fmw, int nwf,
                                         unsigned *xw) {
    const int N = C.N;
    double sum = 0.0;
    for (int q = 0; q < N; ++q) sum += (double) mulmod(dg[q * dstride], C.part_inverse[q], C.moduli[q], C.barrett[q]) / (double) C.moduli[q];
    long long Rk = (long long) floor(sum);
    for (int pass = 0; pass < 3; ++pass) {
        unsigned long long clo = 0, chi = 0;
        for (int w = 0; w < nwf; ++w) {
            unsigned long long lo = clo, hi = chi;
            for (int q = 0; q < N; ++q) {
                const unsigned long long xi = (unsigned long long) (unsigned) mulmod(dg[q * dstride], C.part_inverse[q], C.moduli[q], C.barrett[q]);
                const unsigned long long p = xi * fmi[q * nwf + w];
                lo += p; hi += lo < p ? 1ull : 0ull;
            }
            const unsigned long long p = (unsigned long long) Rk * fneg[w];
            lo += p; hi += lo < p ? 1ull : 0ull;
            xw[w] = (unsigned) lo;
            clo = (lo >> 32) | (hi << 32); chi = hi >> 32;
        }
        bool ge = true;
        for (int w = nwf - 1; w >= 0; --w) { if (xw[w] != fmw[w]) { ge = xw[w] > fmw[w]; break; } }
        if (!ge) break;
        if ((int) xw[nwf - 1] < 0) --Rk; else ++Rk;
    }
}

__global__ void k_scalar_binary(const DevConsts *Cp, SoA alpha, SoA beta, const unsigned *fmi, const unsigned *fneg, const unsigned *fmw, int nwf, int pw,
                                ScalBin *out, int *flag) {
    if (threadIdx.x > 0 || blockIdx.x > 0) return;
    const DevConsts &C = *Cp;
    unsigned x[kMaxN + 2];
    int ok = 1;
    for (int which = 0; which < 2; ++which) {
        const SoA &s = which == 0 ? alpha : beta;
        ScalBin &o = out[which];
        for (int w = 0; w < kScalW; ++w) o.w[w] = 0u;
        int L = 0;
        if (s.eval[s.len()].frac != 0) {
            crt_full(C, s.digits, 1, fmi, fneg, fmw, nwf, x);
            for (int w = 0; w < nwf; ++w) if (x[w]) L = 32 * w + 32 - __clz(x[w]);
            for (int w = 0; w < nwf && w < kScalW; ++w) o.w[w] = x[w];
        }
        o.len = L; o.exp = s.exp[0]; o.sign = s.sign[0]; o.pad = 0;
        if (L > 32 * pw) ok = 0;
    }
    *flag = ok;
}

struct BinTabs {                 // full-base CRT tables (device) and the per-call scalars
    const unsigned *fmi, *fneg, *fmw;
    int nwf;
    const ScalBin *scal;
    const int *ok;
};

// NQ moduli; PW = words of a value of MP_PRECISION + 1 bits; NWF = words of M + 1; BW = words of the exact sum: kBinW (one set of planes)
// or kBinBig (significands cut into slices: S = sum_d S_d 2^(width d) over the 2 slices - 1 plane sets [d][P] of stage 2)
template <int NQ, int PW, int NWF, int BW>
__global__ void __launch_bounds__(kBinT) k_bin_norm2(const DevConsts *Cp, BinTabs T, int m, int n, const uint8_t *S8, long long m_ps, long long n_ps, const int *sel,
                                                     const OuterInfo *ia, const OuterInfo *ib, SoA Cm, int ldc) {
    extern __shared__ __align__(16) uint8_t bin_smem[];
    const int P = sel[0];
    if (P <= 0 || *T.ok == 0) return;
    const DevConsts &C = *Cp;
    const SmallDev &SD = *C.small;
    constexpr int SW = kBinW;                       // words of one sum rebuilt from the one-byte base
    constexpr int AW = 2 * PW + 2;                  // words of the exact t1 + t2
    const int slices = BW > SW ? max(1, sel[kSelSlices]) : 1, ND = 2 * slices - 1;
    const int width = BW > SW ? sel[kSelWidth] : 0;
    // shared: X8 [56][128] | mi4 [54][3] | negmp [12] | c4 [64] | fmi [NQ][NWF] | fneg, fmw [NWF] | p32 [PW][NQ] | al, be [PW] | cd [128][NQ + 1]
    uint8_t *X8 = bin_smem;
    uint4 *mi4 = (uint4 *) (X8 + 56 * kBinT);
    uint4 *c4 = mi4 + kSmallMax * (kBinW / 4);
    unsigned *negmp = (unsigned *) (c4 + 64);
    unsigned *fmi = negmp + kBinW;
    unsigned *fneg = fmi + NQ * NWF, *fmw = fneg + NWF;
    unsigned *p32 = fmw + NWF;
    unsigned *s_al = p32 + PW * NQ, *s_be = s_al + PW;
    int *cd = (int *) (s_be + PW);
    __shared__ int s_meta[8];                       // alpha: len, exp, sign; beta: len, exp, sign
    for (int v = threadIdx.x; v < P * (kBinW / 4); v += kBinT) mi4[v] = __ldg((const uint4 *) (SD.bin_mi + (size_t) P * kSmallMax * kBinW) + v);
    if (threadIdx.x < kBinW) negmp[threadIdx.x] = SD.bin_negmp[P * kBinW + threadIdx.x];
    if (threadIdx.x < 64) {
        const int j = threadIdx.x;
        c4[j] = make_uint4((unsigned) SD.p[j], SD.mu[j], (unsigned) SD.inv[P * 64 + j], __float_as_uint(SD.rcp[j]));
    }
    for (int v = threadIdx.x; v < NQ * NWF; v += kBinT) fmi[v] = T.fmi[v];
    for (int v = threadIdx.x; v < NWF; v += kBinT) { fneg[v] = T.fneg[v]; fmw[v] = T.fmw[v]; }
    for (int v = threadIdx.x; v < PW * NQ; v += kBinT) { const int w = v / NQ, q = v - w * NQ; p32[v] = (unsigned) C.pow2[(long long) (32 * w) * NQ + q]; }
    for (int v = threadIdx.x; v < PW; v += kBinT) { s_al[v] = T.scal[0].w[v]; s_be[v] = T.scal[1].w[v]; }
    if (threadIdx.x == 0) {
        s_meta[0] = T.scal[0].len; s_meta[1] = T.scal[0].exp; s_meta[2] = T.scal[0].sign;
        s_meta[3] = T.scal[1].len; s_meta[4] = T.scal[1].exp; s_meta[5] = T.scal[1].sign;
    }
    const int prec = C.precision;
    const int tiles = (m + kBinT - 1) / kBinT;
    const long long total = (long long) tiles * n;
    const long long plane = n_ps * m_ps;
    constexpr int CP = NQ + 1;
    for (long long tl = blockIdx.x; tl < total; tl += gridDim.x) {
        const int col = (int) (tl / tiles);
        const int row0 = (int) (tl - (long long) col * tiles) * kBinT;
        const int rows_live = min(kBinT, m - row0);
        __syncthreads();
        {
            const uint8_t *src = S8 + (long long) col * m_ps + row0;
            for (int v = threadIdx.x; v < P * (kBinT / 16); v += kBinT) {
                const int j = v >> 3, part = v & 7;
                cp_async16(X8 + j * kBinT + part * 16, src + (long long) j * plane + part * 16);
            }
            // the digits of the block's C entries: one contiguous run, copied into padded rows
            const int *csrc = Cm.digits + (row0 + (long long) col * ldc) * NQ;
            for (int v = threadIdx.x; v < rows_live * NQ; v += kBinT) {
                const unsigned sa = (unsigned) __cvta_generic_to_shared(cd + (v / NQ) * CP + v % NQ);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(sa), "l"(csrc + v));
            }
            cp_async_commit();
            cp_async_wait<0>();
        }
        __syncthreads();
        const int row = row0 + threadIdx.x;
        bool ok = false;
        // ---- the exact sum S of this thread's entry (two's complement, then sign and magnitude) ----
        unsigned Sx[BW];
        int sSx = 0;
#pragma unroll
        for (int w = 0; w < BW; ++w) Sx[w] = 0u;
        {
            bool line_ok = false;
            if (row < m) { const OuterInfo ra0 = ia[row], cb0 = ib[col]; line_ok = ra0.win >= 0 && cb0.win >= 0 && s_meta[0] > 0; }
            for (int d = 0; d < ND; ++d) {
                if (d > 0) {                                  // the next slice sum's planes
                    __syncthreads();
                    const uint8_t *src = S8 + (long long) col * m_ps + row0;
                    for (int v = threadIdx.x; v < P * (kBinT / 16); v += kBinT) {
                        const int j = v >> 3, part = v & 7;
                        cp_async16(X8 + j * kBinT + part * 16, src + (long long) (d * P + j) * plane + part * 16);
                    }
                    cp_async_commit();
                    cp_async_wait<0>();
                    __syncthreads();
                }
                if (line_ok) {
                    uint8_t *xs = X8 + threadIdx.x;
                    // xi_i = x_i (M'/p_i)^-1 mod p_i, the rank R = nearest integer of sum xi_i / p_i, and the column sums of xi_i (M'/p_i) - R M' word by word:
                    // one pass over the moduli, twelve accumulators (each below 2^46), carries at the end
                    unsigned x[SW];
                    {
                        float sum = 0.f;
                        unsigned long long cw[SW];
#pragma unroll
                        for (int w = 0; w < SW; ++w) cw[w] = 0ull;
                        for (int j = 0; j < P; ++j) {
                            const uint4 cj = c4[j];
                            const unsigned tj = (unsigned) xs[j * kBinT] * cj.z;
                            const unsigned rj = tj - __umulhi(tj, cj.y) * cj.x;
                            const unsigned long long xi = min(rj, rj - cj.x);
                            sum = fmaf((float) (unsigned) xi, __uint_as_float(cj.w), sum);
#pragma unroll
                            for (int w4 = 0; w4 < SW / 4; ++w4) {
                                const uint4 mm = mi4[j * (kBinW / 4) + w4];
                                cw[4 * w4] += xi * mm.x; cw[4 * w4 + 1] += xi * mm.y; cw[4 * w4 + 2] += xi * mm.z; cw[4 * w4 + 3] += xi * mm.w;
                            }
                        }
                        const unsigned long long Rk = (unsigned) __float2int_rn(sum);
                        unsigned long long carry = 0;
#pragma unroll
                        for (int w = 0; w < SW; ++w) { cw[w] += carry + Rk * negmp[w]; x[w] = (unsigned) cw[w]; carry = cw[w] >> 32; }
                    }
                    if (BW == SW) {
#pragma unroll
                        for (int w = 0; w < SW; ++w) Sx[w] = x[w];
                    } else {
                        // S += S_d 2^(width d): sign-extended, shifted and added modulo 2^(32 BW)
                        unsigned tmp[BW];
                        const unsigned ext = (x[SW - 1] >> 31) ? 0xffffffffu : 0u;
#pragma unroll
                        for (int w = 0; w < BW; ++w) tmp[w] = w < SW ? x[w] : ext;
                        mw::shl<BW>(tmp, d * width);
                        unsigned long long cy = 0;
#pragma unroll
                        for (int w = 0; w < BW; ++w) { cy += (unsigned long long) Sx[w] + tmp[w]; Sx[w] = (unsigned) cy; cy >>= 32; }
                    }
                }
            }
            sSx = (int) (Sx[BW - 1] >> 31);
            if (sSx) {
                unsigned long long cy = 1;
#pragma unroll
                for (int w = 0; w < BW; ++w) { cy += (unsigned long long) (~Sx[w]); Sx[w] = (unsigned) cy; cy >>= 32; }
            }
        }
        if (row < m) {
            const long long ic = row + (long long) col * ldc;
            int *mycd = cd + threadIdx.x * CP;
            // ---- t1 = rn(alpha rn(S)) ----
            unsigned t1[PW];
            int e1 = 0, s1 = 0, L1 = 0;
#pragma unroll
            for (int w = 0; w < PW; ++w) t1[w] = 0u;
            const OuterInfo ra = ia[row], cb = ib[col];
            if (ra.win >= 0 && cb.win >= 0 && s_meta[0] > 0) {
                // the exact sum was put together before the row test (block-wide staging of the slice planes): S, sS
                unsigned (&x)[BW] = Sx;
                const int sS = sSx;
                const int Ls = mw::bitlen<BW>(x);
                if (Ls > 0) {
                    const int dS = mw::round_to<BW>(x, Ls, prec);
                    unsigned tt[PW], al[PW], pr[2 * PW];
#pragma unroll
                    for (int w = 0; w < PW; ++w) { tt[w] = w < BW ? x[w] : 0u; al[w] = s_al[w]; }
                    mw::mul<PW, PW>(tt, al, pr);
                    const int Lp = mw::bitlen<2 * PW>(pr);
                    const int d1 = mw::round_to<2 * PW>(pr, Lp, prec);
#pragma unroll
                    for (int w = 0; w < PW; ++w) t1[w] = pr[w];
                    L1 = mw::bitlen<PW>(t1);
                    e1 = ra.emin + cb.emin + dS + s_meta[1] + d1;
                    s1 = sS ^ s_meta[2];
                }
            }
            // ---- t2 = rn(beta C) ----
            unsigned t2[PW];
            int e2 = 0, s2 = 0, L2 = 0;
#pragma unroll
            for (int w = 0; w < PW; ++w) t2[w] = 0u;
            if (s_meta[3] > 0 && Cm.eval[ic + Cm.len()].frac != 0) {
                unsigned cx[NWF];
                {
                    // CRT over all moduli (crt_full with the tables in shared memory and the digits in this thread's padded row)
                    unsigned xi[NQ];
                    double sum = 0.0;
#pragma unroll
                    for (int q = 0; q < NQ; ++q) {
                        xi[q] = (unsigned) mulmod(mycd[q], C.part_inverse[q], C.moduli[q], C.barrett[q]);
                        sum += (double) xi[q] / (double) C.moduli[q];
                    }
                    long long Rk = (long long) floor(sum);
                    for (int pass = 0; pass < 3; ++pass) {
                        unsigned long long clo = 0, chi = 0;
#pragma unroll
                        for (int w = 0; w < NWF; ++w) {
                            unsigned long long lo = clo, hi = chi;
#pragma unroll
                            for (int q = 0; q < NQ; ++q) {
                                const unsigned long long p = (unsigned long long) xi[q] * fmi[q * NWF + w];
                                lo += p; hi += lo < p ? 1ull : 0ull;
                            }
                            const unsigned long long p = (unsigned long long) Rk * fneg[w];
                            lo += p; hi += lo < p ? 1ull : 0ull;
                            cx[w] = (unsigned) lo;
                            clo = (lo >> 32) | (hi << 32); chi = hi >> 32;
                        }
                        int ge = 0;
#pragma unroll
                        for (int w = 0; w < NWF; ++w) ge = cx[w] != fmw[w] ? (cx[w] > fmw[w] ? 1 : -1) : ge;
                        if (ge < 0) break;
                        if ((int) cx[NWF - 1] < 0) --Rk; else ++Rk;
                    }
                }
                unsigned be[PW], pr[NWF + PW];
#pragma unroll
                for (int w = 0; w < PW; ++w) be[w] = s_be[w];
                mw::mul<NWF, PW>(cx, be, pr);
                const int Lp = mw::bitlen<NWF + PW>(pr);
                if (Lp > 0) {
                    const int d2 = mw::round_to<NWF + PW>(pr, Lp, prec);
#pragma unroll
                    for (int w = 0; w < PW; ++w) t2[w] = pr[w];
                    L2 = mw::bitlen<PW>(t2);
                    e2 = Cm.exp[ic] + s_meta[4] + d2;
                    s2 = (Cm.sign[ic] ^ s_meta[5]) & 1;
                }
            }
            // ---- C = rn(t1 + t2) ----
            unsigned r[PW];
            int er = 0, sr = 0, Lr = 0;
            if (L1 == 0 || (L2 > 0 && (e2 + L2) - (e1 + L1) > prec + 2)) {
#pragma unroll
                for (int w = 0; w < PW; ++w) r[w] = t2[w];
                er = e2; sr = s2; Lr = L2;
            } else if (L2 == 0 || (e1 + L1) - (e2 + L2) > prec + 2) {
#pragma unroll
                for (int w = 0; w < PW; ++w) r[w] = t1[w];
                er = e1; sr = s1; Lr = L1;
            } else {
                unsigned a[AW], b[AW];
#pragma unroll
                for (int w = 0; w < AW; ++w) { a[w] = w < PW ? t1[w] : 0u; b[w] = w < PW ? t2[w] : 0u; }
                const int emin = min(e1, e2);
                mw::shl<AW>(a, e1 - emin);
                mw::shl<AW>(b, e2 - emin);
                if (s1 == s2) {
                    unsigned long long cy = 0;
#pragma unroll
                    for (int w = 0; w < AW; ++w) { cy += (unsigned long long) a[w] + b[w]; a[w] = (unsigned) cy; cy >>= 32; }
                    sr = s1;
                } else {
                    const int cm = mw::cmp<AW>(a, b);
                    long long bw = 0;
#pragma unroll
                    for (int w = 0; w < AW; ++w) {
                        const long long df = cm >= 0 ? (long long) a[w] - b[w] - bw : (long long) b[w] - a[w] - bw;
                        a[w] = (unsigned) df;
                        bw = df < 0 ? 1 : 0;
                    }
                    sr = cm >= 0 ? s1 : s2;
                }
                const int La = mw::bitlen<AW>(a);
                const int dr = mw::round_to<AW>(a, La, prec);
#pragma unroll
                for (int w = 0; w < PW; ++w) r[w] = a[w];
                Lr = mw::bitlen<PW>(r);
                er = emin + dr;
            }
            Er lo, up;
            lo.frac = 0; lo.exp = 0; up.frac = 0; up.exp = 0;
            if (Lr == 0) {
                er = 0; sr = 0;
#pragma unroll 8
                for (int q = 0; q < NQ; ++q) mycd[q] = 0;
            } else {
                bin_eval<PW>(C, r, Lr, lo, up);
#pragma unroll 4
                for (int q = 0; q < NQ; ++q) {
                    unsigned long long acc = 0;
#pragma unroll
                    for (int w = 0; w < PW; ++w) {
                        acc += (unsigned long long) r[w] * p32[w * NQ + q];
                        if ((w & 7) == 7) acc = (unsigned long long) (unsigned) reduce64(acc, C.moduli[q], C.barrett[q]);
                    }
                    mycd[q] = reduce64(acc, C.moduli[q], C.barrett[q]);
                }
            }
            Cm.sign[ic] = sr;
            Cm.exp[ic] = er;
            Cm.eval[ic] = lo;
            Cm.eval[ic + Cm.len()] = up;
            ok = true;
        }
        (void) ok;
        __syncthreads();
        int4 *cd4 = (int4 *) (Cm.digits + (row0 + (long long) col * ldc) * NQ);
        for (int v = threadIdx.x; v < rows_live * (NQ / 4); v += kBinT) {
            const int ent = (4 * v) / NQ;
            const int *src = cd + ent * CP + (4 * v) % NQ;
            cd4[v] = make_int4(src[0], src[1], src[2], src[3]);
        }
    }
}
// every entry of a segment into the todo list (*gate == 0: neither binary epilogue could take the call)
__global__ void k_fill_todo(long long *todo, int *todo_count, int m, int nc, const int *gate) {
    if (*gate != 0) return;
    const long long total = (long long) m * nc;
    for (long long e = (long long) blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long) gridDim.x * blockDim.x) todo[e] = e;
    if (blockIdx.x == 0 && threadIdx.x == 0) *todo_count = (int) total;
}

// ---- the same epilogue for the wider formats (N >= 16) -----------------------------------------------------------------------------------
// k_bin_norm2 keeps every multiword value in registers; from N = 16 on (28-word significands of C, 42-word products, 32-word slice-sum
// accumulators) that no longer fits: the compiler spilled 8.5 KB per thread and the kernel took 127 of 193 ms at 4096^3 / 424 bit.  Here
// the long values live in a per-thread column of shared memory (word w of thread t at [w][t]: conflict-free), walked by short run-time
// loops; only the MP_PRECISION-sized values (T, t1, t2, the result) and the 2 PW-word product alpha T stay in registers.
constexpr int kBin3RA = 2 * kBinBig - 20;        // words of scratch region A (slice-sum accumulator 33, then the product beta C: NWF + PW <= 44, then a)
constexpr int kBin3RB = 32;                     // region B: C in binary [NWF <= 28] (the xi_q overwrite the entry's staged digits); then b
template <int NQ, int PW, int NWF>
__global__ void __launch_bounds__(kBinT, 3) k_bin_norm3(const DevConsts *Cp, BinTabs T, int m, int n, const uint8_t *S8, long long m_ps, long long n_ps, const int *sel,
                                                        const OuterInfo *ia, const OuterInfo *ib, SoA Cm, int ldc) {
    extern __shared__ __align__(16) uint8_t bin_smem[];
    static_assert(NWF + PW <= kBin3RA && 2 * PW + 2 <= kBin3RA && 2 * PW + 2 <= kBin3RB && NWF <= kBin3RB && kBinBig + 1 <= kBin3RA, "scratch regions");
    const int P = sel[0];
    if (P <= 0 || *T.ok == 0) return;
    const DevConsts &C = *Cp;
    const SmallDev &SD = *C.small;
    constexpr int SW = kBinW, BW = kBinBig, AW = 2 * PW + 2;
    const int slices = max(1, sel[kSelSlices]), ND = 2 * slices - 1;
    const int width = slices > 1 ? sel[kSelWidth] : 0;
    uint8_t *X8 = bin_smem;
    uint4 *mi4 = (uint4 *) (X8 + 56 * kBinT);
    uint4 *c4 = mi4 + kSmallMax * (kBinW / 4);
    unsigned *negmp = (unsigned *) (c4 + 64);
    unsigned *fmi = negmp + kBinW;
    unsigned *fneg = fmi + NQ * NWF, *fmw = fneg + NWF;
    unsigned *p32 = fmw + NWF;
    unsigned *s_al = p32 + PW * NQ, *s_be = s_al + PW;
    int *cd = (int *) (s_be + PW);
    unsigned *RA = (unsigned *) (cd + kBinT * (NQ + 1)), *RB = RA + kBin3RA * kBinT;
    __shared__ int s_meta[8];
    __shared__ double s_rcpm[NQ];
    for (int v = threadIdx.x; v < P * (kBinW / 4); v += kBinT) mi4[v] = __ldg((const uint4 *) (SD.bin_mi + (size_t) P * kSmallMax * kBinW) + v);
    if (threadIdx.x < kBinW) negmp[threadIdx.x] = SD.bin_negmp[P * kBinW + threadIdx.x];
    if (threadIdx.x < 64) {
        const int j = threadIdx.x;
        c4[j] = make_uint4((unsigned) SD.p[j], SD.mu[j], (unsigned) SD.inv[P * 64 + j], __float_as_uint(SD.rcp[j]));
    }
    for (int v = threadIdx.x; v < NQ * NWF; v += kBinT) fmi[v] = T.fmi[v];
    for (int v = threadIdx.x; v < NWF; v += kBinT) { fneg[v] = T.fneg[v]; fmw[v] = T.fmw[v]; }
    for (int v = threadIdx.x; v < PW * NQ; v += kBinT) { const int w = v / NQ, q = v - w * NQ; p32[v] = (unsigned) C.pow2[(long long) (32 * w) * NQ + q]; }
    for (int v = threadIdx.x; v < PW; v += kBinT) { s_al[v] = T.scal[0].w[v]; s_be[v] = T.scal[1].w[v]; }
    for (int v = threadIdx.x; v < NQ; v += kBinT) s_rcpm[v] = 1.0 / (double) C.moduli[v];
    if (threadIdx.x == 0) {
        s_meta[0] = T.scal[0].len; s_meta[1] = T.scal[0].exp; s_meta[2] = T.scal[0].sign;
        s_meta[3] = T.scal[1].len; s_meta[4] = T.scal[1].exp; s_meta[5] = T.scal[1].sign;
    }
    const int prec = C.precision;
    const int tiles = (m + kBinT - 1) / kBinT;
    const long long total = (long long) tiles * n;
    const long long plane = n_ps * m_ps;
    constexpr int CP = NQ + 1;
    unsigned *ra = RA + threadIdx.x, *rb = RB + threadIdx.x;            // word w at ra[w * kBinT]
    // helpers on a scratch column
    auto bitlen_s = [](const unsigned *col, int nw) { int L = 0; for (int w = nw - 1; w >= 0; --w) { const unsigned v = col[w * kBinT]; if (v) { L = 32 * w + 32 - __clz(v); break; } } return L; };
    // round the magnitude in col (nw words, bit length L < 32 nw) to prec bits, nearest, ties away; the kept bits to out[PW]; returns the bits dropped
    auto round_s = [&](unsigned *col, int nw, int L, unsigned (&out)[PW]) {
        int drop = L - prec;
        if (drop > 0) {
            unsigned long long cy = 1ull << ((drop - 1) & 31);
            for (int w = (drop - 1) >> 5; w < nw && cy; ++w) { cy += col[w * kBinT]; col[w * kBinT] = (unsigned) cy; cy >>= 32; }
        } else {
            drop = 0;
        }
        const int ws = drop >> 5, bs = drop & 31;
#pragma unroll
        for (int w = 0; w < PW; ++w) {
            const unsigned lo = w + ws < nw ? col[(w + ws) * kBinT] : 0u, hi = w + ws + 1 < nw ? col[(w + ws + 1) * kBinT] : 0u;
            out[w] = __funnelshift_r(lo, hi, bs);
        }
        return drop;
    };
    // col (nw words, zeroed by the caller) = v << sh, v of PW words
    auto place_s = [](unsigned *col, int nw, const unsigned (&v)[PW], int sh) {
        const int ws = sh >> 5, bs = sh & 31;
#pragma unroll
        for (int j = 0; j <= PW; ++j) {
            const unsigned cur = j < PW ? v[j] : 0u, prev = j > 0 ? v[j - 1] : 0u;
            if (ws + j < nw) col[(ws + j) * kBinT] = __funnelshift_l(prev, cur, bs);
        }
    };
    for (long long tl = blockIdx.x; tl < total; tl += gridDim.x) {
        const int col = (int) (tl / tiles);
        const int row0 = (int) (tl - (long long) col * tiles) * kBinT;
        const int rows_live = min(kBinT, m - row0);
        __syncthreads();
        {
            const uint8_t *src = S8 + (long long) col * m_ps + row0;
            for (int v = threadIdx.x; v < P * (kBinT / 16); v += kBinT) {
                const int j = v >> 3, part = v & 7;
                cp_async16(X8 + j * kBinT + part * 16, src + (long long) j * plane + part * 16);
            }
            const int *csrc = Cm.digits + (row0 + (long long) col * ldc) * NQ;
            for (int v = threadIdx.x; v < rows_live * NQ; v += kBinT) {
                const unsigned sa = (unsigned) __cvta_generic_to_shared(cd + (v / NQ) * CP + v % NQ);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(sa), "l"(csrc + v));
            }
            cp_async_commit();
            cp_async_wait<0>();
        }
        __syncthreads();
        const int row = row0 + threadIdx.x;
        OuterInfo ra0 = {}, cb0 = {};
        bool line_ok = false;
        if (row < m) { ra0 = ia[row]; cb0 = ib[col]; line_ok = ra0.win >= 0 && cb0.win >= 0 && s_meta[0] > 0; }
        // ---- the exact sum: S = sum_d S_d 2^(width d) in region A (two's complement, BW words) ----
        for (int w = 0; w <= BW; ++w) ra[w * kBinT] = 0u;
        for (int d = 0; d < ND; ++d) {
            if (d > 0) {
                __syncthreads();
                const uint8_t *src = S8 + (long long) col * m_ps + row0;
                for (int v = threadIdx.x; v < P * (kBinT / 16); v += kBinT) {
                    const int j = v >> 3, part = v & 7;
                    cp_async16(X8 + j * kBinT + part * 16, src + (long long) (d * P + j) * plane + part * 16);
                }
                cp_async_commit();
                cp_async_wait<0>();
                __syncthreads();
            }
            if (line_ok) {
                uint8_t *xs = X8 + threadIdx.x;
                // xi_i = x_i (M'/p_i)^-1 mod p_i, the rank R = nearest integer of sum xi_i / p_i, and the column sums of xi_i (M'/p_i) - R M' word by word:
                // one pass over the moduli, twelve accumulators (each below 2^46), carries at the end
                unsigned x[SW];
                {
                    float sum = 0.f;
                    unsigned long long cw[SW];
#pragma unroll
                    for (int w = 0; w < SW; ++w) cw[w] = 0ull;
                    for (int j = 0; j < P; ++j) {
                        const uint4 cj = c4[j];
                        const unsigned tj = (unsigned) xs[j * kBinT] * cj.z;
                        const unsigned rj = tj - __umulhi(tj, cj.y) * cj.x;
                        const unsigned long long xi = min(rj, rj - cj.x);
                        sum = fmaf((float) (unsigned) xi, __uint_as_float(cj.w), sum);
#pragma unroll
                        for (int w4 = 0; w4 < SW / 4; ++w4) {
                            const uint4 mm = mi4[j * (kBinW / 4) + w4];
                            cw[4 * w4] += xi * mm.x; cw[4 * w4 + 1] += xi * mm.y; cw[4 * w4 + 2] += xi * mm.z; cw[4 * w4 + 3] += xi * mm.w;
                        }
                    }
                    const unsigned long long Rk = (unsigned) __float2int_rn(sum);
                    unsigned long long carry = 0;
#pragma unroll
                    for (int w = 0; w < SW; ++w) { cw[w] += carry + Rk * negmp[w]; x[w] = (unsigned) cw[w]; carry = cw[w] >> 32; }
                }
                // S += sign-extended x << (width d), modulo 2^(32 BW)
                const int sh = d * width, ws = sh >> 5, bs = sh & 31;
                const unsigned ext = (x[SW - 1] >> 31) ? 0xffffffffu : 0u;
                unsigned long long cy = 0;
#pragma unroll
                for (int j = 0; j <= SW; ++j) {
                    const unsigned cur = j < SW ? x[j] : ext, prev = j > 0 ? x[j - 1] : 0u;
                    if (ws + j < BW) { cy += (unsigned long long) ra[(ws + j) * kBinT] + __funnelshift_l(prev, cur, bs); ra[(ws + j) * kBinT] = (unsigned) cy; cy >>= 32; }
                }
                for (int w = ws + SW + 1; w < BW; ++w) { cy += (unsigned long long) ra[w * kBinT] + ext; ra[w * kBinT] = (unsigned) cy; cy >>= 32; }
            }
        }
        if (row < m) {
            const long long ic = row + (long long) col * ldc;
            int *mycd = cd + threadIdx.x * CP;
            // ---- t1 = rn(alpha rn(S)) ----
            unsigned t1[PW];
            int e1 = 0, s1 = 0, L1 = 0;
#pragma unroll
            for (int w = 0; w < PW; ++w) t1[w] = 0u;
            if (line_ok) {
                const int sS = (int) (ra[(BW - 1) * kBinT] >> 31);
                if (sS) {
                    unsigned long long cy = 1;
                    for (int w = 0; w < BW; ++w) { cy += (unsigned long long) (~ra[w * kBinT]); ra[w * kBinT] = (unsigned) cy; cy >>= 32; }
                }
                const int Ls = bitlen_s(ra, BW);
                if (Ls > 0) {
                    unsigned tt[PW], al[PW], pr[2 * PW];
                    const int dS = round_s(ra, BW + 1, Ls, tt);
#pragma unroll
                    for (int w = 0; w < PW; ++w) al[w] = s_al[w];
                    mw::mul<PW, PW>(tt, al, pr);
                    const int Lp = mw::bitlen<2 * PW>(pr);
                    const int d1 = mw::round_to<2 * PW>(pr, Lp, prec);
#pragma unroll
                    for (int w = 0; w < PW; ++w) t1[w] = pr[w];
                    L1 = mw::bitlen<PW>(t1);
                    e1 = ra0.emin + cb0.emin + dS + s_meta[1] + d1;
                    s1 = sS ^ s_meta[2];
                }
            }
            // ---- t2 = rn(beta C): C in binary (region B: xi, then the words), the product in region A ----
            unsigned t2[PW];
            int e2 = 0, s2 = 0, L2 = 0;
#pragma unroll
            for (int w = 0; w < PW; ++w) t2[w] = 0u;
            if (s_meta[3] > 0 && Cm.eval[ic + Cm.len()].frac != 0) {
                double sum = 0.0;
#pragma unroll 4
                for (int q = 0; q < NQ; ++q) {
                    const unsigned xi = (unsigned) mulmod(mycd[q], C.part_inverse[q], C.moduli[q], C.barrett[q]);
                    mycd[q] = (int) xi;                                  // (the digits are not needed again)
                    sum += (double) xi * s_rcpm[q];
                }
                long long Rk = (long long) floor(sum);
                unsigned *cx = rb;
                for (int pass = 0; pass < 3; ++pass) {
                    unsigned long long clo = 0, chi = 0;
                    for (int w = 0; w < NWF; ++w) {
                        unsigned long long lo = clo, hi = chi;
#pragma unroll 4
                        for (int q = 0; q < NQ; ++q) {
                            const unsigned long long p = (unsigned long long) (unsigned) mycd[q] * fmi[q * NWF + w];
                            lo += p; hi += lo < p ? 1ull : 0ull;
                        }
                        const unsigned long long p = (unsigned long long) Rk * fneg[w];
                        lo += p; hi += lo < p ? 1ull : 0ull;
                        cx[w * kBinT] = (unsigned) lo;
                        clo = (lo >> 32) | (hi << 32); chi = hi >> 32;
                    }
                    int ge = 1;
                    for (int w = NWF - 1; w >= 0; --w) { const unsigned v = cx[w * kBinT]; if (v != fmw[w]) { ge = v > fmw[w] ? 1 : -1; break; } }
                    if (ge < 0) break;
                    if ((int) cx[(NWF - 1) * kBinT] < 0) --Rk; else ++Rk;
                }
                const int ncw = (bitlen_s(cx, NWF) + 31) >> 5;
                unsigned be[PW];
#pragma unroll
                for (int w = 0; w < PW; ++w) be[w] = s_be[w];
                for (int w = 0; w < NWF + PW; ++w) ra[w * kBinT] = 0u;
                for (int i = 0; i < ncw; ++i) {
                    const unsigned long long ci = cx[i * kBinT];
                    unsigned long long cy = 0;
#pragma unroll
                    for (int j = 0; j < PW; ++j) {
                        const unsigned long long t = ci * be[j] + ra[(i + j) * kBinT] + cy;
                        ra[(i + j) * kBinT] = (unsigned) t;
                        cy = t >> 32;
                    }
                    ra[(i + PW) * kBinT] = (unsigned) cy;
                }
                const int Lp = bitlen_s(ra, NWF + PW);
                if (Lp > 0) {
                    const int d2 = round_s(ra, NWF + PW, Lp, t2);
                    L2 = mw::bitlen<PW>(t2);
                    e2 = Cm.exp[ic] + s_meta[4] + d2;
                    s2 = (Cm.sign[ic] ^ s_meta[5]) & 1;
                }
            }
            // ---- C = rn(t1 + t2) ----
            unsigned r[PW];
            int er = 0, sr = 0, Lr = 0;
            if (L1 == 0 || (L2 > 0 && (e2 + L2) - (e1 + L1) > prec + 2)) {
#pragma unroll
                for (int w = 0; w < PW; ++w) r[w] = t2[w];
                er = e2; sr = s2; Lr = L2;
            } else if (L2 == 0 || (e1 + L1) - (e2 + L2) > prec + 2) {
#pragma unroll
                for (int w = 0; w < PW; ++w) r[w] = t1[w];
                er = e1; sr = s1; Lr = L1;
            } else {
                const int emin = min(e1, e2);
                for (int w = 0; w < AW; ++w) { ra[w * kBinT] = 0u; rb[w * kBinT] = 0u; }
                place_s(ra, AW, t1, e1 - emin);
                place_s(rb, AW, t2, e2 - emin);
                if (s1 == s2) {
                    unsigned long long cy = 0;
                    for (int w = 0; w < AW; ++w) { cy += (unsigned long long) ra[w * kBinT] + rb[w * kBinT]; ra[w * kBinT] = (unsigned) cy; cy >>= 32; }
                    sr = s1;
                } else {
                    int cm = 0;
                    for (int w = AW - 1; w >= 0; --w) { const unsigned a = ra[w * kBinT], b = rb[w * kBinT]; if (a != b) { cm = a > b ? 1 : -1; break; } }
                    long long bw = 0;
                    for (int w = 0; w < AW; ++w) {
                        const long long a = ra[w * kBinT], b = rb[w * kBinT];
                        const long long df = cm >= 0 ? a - b - bw : b - a - bw;
                        ra[w * kBinT] = (unsigned) df;
                        bw = df < 0 ? 1 : 0;
                    }
                    sr = cm >= 0 ? s1 : s2;
                }
                const int La = bitlen_s(ra, AW);
                const int dr = round_s(ra, AW, La, r);
                Lr = mw::bitlen<PW>(r);
                er = emin + dr;
            }
            Er lo, up;
            lo.frac = 0; lo.exp = 0; up.frac = 0; up.exp = 0;
            if (Lr == 0) {
                er = 0; sr = 0;
#pragma unroll 8
                for (int q = 0; q < NQ; ++q) mycd[q] = 0;
            } else {
                bin_eval<PW>(C, r, Lr, lo, up);
#pragma unroll 2
                for (int q = 0; q < NQ; ++q) {
                    unsigned long long acc = 0;
#pragma unroll
                    for (int w = 0; w < PW; ++w) {
                        acc += (unsigned long long) r[w] * p32[w * NQ + q];
                        if ((w & 7) == 7) acc = (unsigned long long) (unsigned) reduce64(acc, C.moduli[q], C.barrett[q]);
                    }
                    mycd[q] = reduce64(acc, C.moduli[q], C.barrett[q]);
                }
            }
            Cm.sign[ic] = sr;
            Cm.exp[ic] = er;
            Cm.eval[ic] = lo;
            Cm.eval[ic + Cm.len()] = up;
        }
        __syncthreads();
        int4 *cd4 = (int4 *) (Cm.digits + (row0 + (long long) col * ldc) * NQ);
        for (int v = threadIdx.x; v < rows_live * (NQ / 4); v += kBinT) {
            const int ent = (4 * v) / NQ;
            const int *src = cd + ent * CP + (4 * v) % NQ;
            cd4[v] = make_int4(src[0], src[1], src[2], src[3]);
        }
    }
}
template <int NQ, int PW, int NWF>
__host__ __device__ inline size_t bin3_smem_bytes() {
    return (size_t) 56 * kBinT + (size_t) kSmallMax * kBinW * 4 + 64 * 16 + kBinW * 4 + ((size_t) NQ * NWF + 2 * NWF + (size_t) PW * NQ + 2 * PW) * 4 +
           (size_t) kBinT * (NQ + 1) * 4 + (size_t) (kBin3RA + kBin3RB) * kBinT * 4 + 64;
}

template <int NQ, int PW, int NWF>
__host__ __device__ inline size_t bin2_smem_bytes() {
    return (size_t) 56 * kBinT + (size_t) kSmallMax * kBinW * 4 + 64 * 16 + kBinW * 4 + ((size_t) NQ * NWF + 2 * NWF + (size_t) PW * NQ + 2 * PW) * 4 +
           (size_t) kBinT * (NQ + 1) * 4 + 64;
}

}  // namespace mpres

// ---- host side: stage 3 of a call (or of one column segment of it) whose exact sums are rounded in binary ---------------------------------
// first: the scalars have to be converted (once per call).  The binary epilogue runs when the format has an instantiation and alpha, beta
// fit MP_PRECISION + 1 bits; otherwise the flag stays 0 and the residue-parallel epilogue (k_bin_norm) takes the segment.
template <int NQ, int PW, int NWF>
static inline void bin2_launch(mpres_ctx *c, bool sliced, const mpres::BinTabs &T, unsigned gx, int m, int nc, const uint8_t *S8s, long long m_ps, long long n_ps,
                               const int *sel, const mpres::OuterInfo *IA, const mpres::OuterInfo *IBs, mpres::SoA Cg, int ldc, cudaStream_t st) {
    const size_t sm = mpres::bin2_smem_bytes<NQ, PW, NWF>();
    if (!(c->attr_bin2 >> (NQ / 8) & 1ull)) {
        cudaFuncSetAttribute(mpres::k_bin_norm2<NQ, PW, NWF, mpres::kBinW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm);
        cudaFuncSetAttribute(mpres::k_bin_norm2<NQ, PW, NWF, mpres::kBinBig>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm);
        c->attr_bin2 |= 1ull << (NQ / 8);
    }
    if (sliced) mpres::k_bin_norm2<NQ, PW, NWF, mpres::kBinBig><<<gx, mpres::kBinT, sm, st>>>(c->dconsts, T, m, nc, S8s, m_ps, n_ps, sel, IA, IBs, Cg, ldc);
    else mpres::k_bin_norm2<NQ, PW, NWF, mpres::kBinW><<<gx, mpres::kBinT, sm, st>>>(c->dconsts, T, m, nc, S8s, m_ps, n_ps, sel, IA, IBs, Cg, ldc);
}
// the shared-memory scratch variant (formats of 16 moduli and more)
template <int NQ, int PW, int NWF>
static inline void bin3_launch(mpres_ctx *c, const mpres::BinTabs &T, int m, int nc, const uint8_t *S8s, long long m_ps, long long n_ps,
                               const int *sel, const mpres::OuterInfo *IA, const mpres::OuterInfo *IBs, mpres::SoA Cg, int ldc, cudaStream_t st) {
    const size_t sm = mpres::bin3_smem_bytes<NQ, PW, NWF>();
    if (!(c->attr_bin2 >> (NQ / 8) & 1ull)) {
        cudaFuncSetAttribute(mpres::k_bin_norm3<NQ, PW, NWF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm);
        c->attr_bin2 |= 1ull << (NQ / 8);
    }
    const unsigned gx = (unsigned) std::min<long long>((long long) ((m + mpres::kBinT - 1) / mpres::kBinT) * nc, (long long) c->sm_count * 3);
    mpres::k_bin_norm3<NQ, PW, NWF><<<gx, mpres::kBinT, sm, st>>>(c->dconsts, T, m, nc, S8s, m_ps, n_ps, sel, IA, IBs, Cg, ldc);
}

// the formats the binary epilogue is instantiated for (0: none)
static inline int bin2_variant(const mpres_ctx *c) {
    const int N = c->hc.N, pw = (c->hc.mp_precision + 1 + 31) / 32, nwf = c->sc.full_nw;
    if (const char *env = getenv("MPRES_BIN_EPILOGUE")) { if (atoi(env) == 0) return 0; }
    switch (N) {
        case 8: return (pw == 4 && nwf == 8) ? 8 : 0;
        case 16: return (pw == 7 && nwf == 15) ? 16 : 0;
        case 24: return (pw == 10 && nwf == 21) ? 24 : 0;
        case 32: return (pw == 14 && nwf == 28) ? 32 : 0;
        default: return 0;
    }
}

// slices > 1: the sums come as slice sums (stage 2 on the pieces of the significands); only the binary epilogue puts them together, so a call
// whose scalars do not fit it is recomputed in reference order (todo list of the segment).
inline int bin_norm_segment(mpres_ctx *c, bool first, int m, int nc, const uint8_t *S8s, long long m_ps, long long n_ps, const int *sel, const mpres::OuterInfo *IA,
                            const mpres::OuterInfo *IBs, mpres::SoA alpha, mpres::SoA beta, mpres::SoA Cg, int ldc, cudaStream_t st, int *launches,
                            int slices = 1, long long *todo = nullptr, int *todo_count = nullptr) {
    using namespace mpres;
    const int N = c->hc.N;
    const SmallConsts &sc = c->sc;
    void *p;
    int rc;
    if ((rc = ws_reserve(c, 21, 2 * sizeof(ScalBin) + 64, &p))) return rc;
    ScalBin *scal = (ScalBin *) p;
    int *ok = (int *) (scal + 2);
    const int pw = (c->hc.mp_precision + 1 + 31) / 32;
    const int variant = bin2_variant(c);
    if (slices > 1 && !variant) return -60;
    const unsigned gx = (unsigned) std::min<long long>((long long) ((m + kBinT - 1) / kBinT) * nc, (long long) c->sm_count * 4);
    if (first) {
        if (variant) {
            k_scalar_binary<<<1, 32, 0, st>>>(c->dconsts, alpha, beta, (const unsigned *) c->d_small[10], (const unsigned *) c->d_small[11],
                                              (const unsigned *) c->d_small[12], sc.full_nw, pw, scal, ok);
            ++*launches;
        } else {
            CUDA_TRY(cudaMemsetAsync(ok, 0, sizeof(int), st));
        }
    }
    if (variant) {
        BinTabs T;
        T.fmi = (const unsigned *) c->d_small[10]; T.fneg = (const unsigned *) c->d_small[11]; T.fmw = (const unsigned *) c->d_small[12];
        T.nwf = sc.full_nw; T.scal = scal; T.ok = ok;
        const bool sl = slices > 1;
        switch (variant) {
            case 8: bin2_launch<8, 4, 8>(c, sl, T, gx, m, nc, S8s, m_ps, n_ps, sel, IA, IBs, Cg, ldc, st); break;
            case 16: bin3_launch<16, 7, 15>(c, T, m, nc, S8s, m_ps, n_ps, sel, IA, IBs, Cg, ldc, st); break;
            case 24: bin3_launch<24, 10, 21>(c, T, m, nc, S8s, m_ps, n_ps, sel, IA, IBs, Cg, ldc, st); break;
            default: bin3_launch<32, 14, 28>(c, T, m, nc, S8s, m_ps, n_ps, sel, IA, IBs, Cg, ldc, st); break;
        }
        ++*launches;
    }
    if (slices > 1) {
        k_fill_todo<<<c->sm_count, 256, 0, st>>>(todo, todo_count, m, nc, ok);
        ++*launches;
        return 0;                                    // (the caller runs k_gemm_todo on the list)
    }
    const size_t smb = bin_smem_bytes(N);
    MPRES_DISPATCH(N, {
        if (!c->attr_bin) { CUDA_TRY(cudaFuncSetAttribute(k_bin_norm<G, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smb)); c->attr_bin = true; }
        k_bin_norm<G, R><<<gx, kBinT, smb, st>>>(c->dconsts, m, nc, S8s, m_ps, n_ps, sel, IA, IBs, alpha, beta, Cg, ldc, ok);
    });
    ++*launches;
    return 0;
}
