// mpres_ops.cu -- operations either side of the GEMM / GEMV / DOT path, a translation unit of their own:
//   mpres_norm (infinity norm; the one norm forwards to mpres_asum)      src/blas/norm.cuh:43, src/mpreduct.cuh:199-291
//   mpres_spmv_csr2st / mpres_spmv_ell2st over mp_collection_t           src/sparse/mpmtx/spmv_mpmtx_csr2st.cuh:106, spmv_mpmtx_ell2st.cuh:119
//   mpres_array_set_d / mpres_array_get_d                                 src/arith/assign.cuh:54-81, 154-180
// Everything numeric runs in the kernels below on the residue-parallel arithmetic of mp_device.cuh; there is no CPU path.
#include "../../include/mpres_b200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <mutex>
#include <vector>

#include "ctx.hpp"
#include "kernels_ref_order.cuh"

#define NEED_DEVICE(c) do { if ((c) && (c)->device < 0) return -100; } while (0)

namespace mpres {

__device__ __forceinline__ long long ops_inc_index(long long i, long long n, int inc) {   // BLAS convention for negative strides
    return inc >= 0 ? i * inc : (n - 1 - i) * (long long) (-inc);
}

// ---- comparison of magnitudes (src/arith/cmpabs.cuh:72-109, src/rns.cuh:1210-1225) -----------------------------------------------
// unsigned comparison of two non-negative extended-range numbers (src/extrange.cuh: er_ucmp)
__device__ __forceinline__ int er_ucmp_dev(const Er &x, const Er &y) {
    if (x.frac == 0 || y.frac == 0) return x.frac == 0 ? (y.frac == 0 ? 0 : -1) : 1;
    if (x.exp != y.exp) return x.exp > y.exp ? 1 : -1;
    return x.frac > y.frac ? 1 : (x.frac < y.frac ? -1 : 0);
}
// 1 if |x| > |y|, -1 if |x| < |y|, 0 if equal
template <int G, int R>
__device__ __forceinline__ int cmp_abs(const DevConsts &C, const Lane<R> &L, const Num<R> &x, const Num<R> &y) {
    const int dexp = x.exp - y.exp;
    int gamma = dexp > 0 ? dexp : 0, theta = dexp < 0 ? -dexp : 0;
    const int nzx = (y.up.frac == 0) || ((long long) theta + y.up.exp) < C.mp_j;
    const int nzy = (x.up.frac == 0) || ((long long) gamma + x.up.exp) < C.mp_j;
    gamma *= nzy; theta *= nzx;
    Er xl = x.lo, xu = x.up, yl = y.lo, yu = y.up;
    xl.exp += gamma; xu.exp += gamma; yl.exp += theta; yu.exp += theta;
    xl.frac *= nzx; xu.frac *= nzx; yl.frac *= nzy; yu.frac *= nzy;
    if (er_ucmp_dev(xl, yu) > 0) return 1;
    if (er_ucmp_dev(yl, xu) > 0) return -1;
    int ax[R], ay[R], diff = 0;
#pragma unroll
    for (int q = 0; q < R; ++q) {
        const int pg = (L.act[q] && nzx) ? pow2_at(C, gamma, L.idx[q], L.m[q], L.mu[q]) : 0;
        const int pt = (L.act[q] && nzy) ? pow2_at(C, theta, L.idx[q], L.m[q], L.mu[q]) : 0;
        ax[q] = mulmod(x.d[q], pg, L.m[q], L.mu[q]);
        ay[q] = mulmod(y.d[q], pt, L.m[q], L.mu[q]);
        diff |= ax[q] ^ ay[q];
    }
    if (gor<G>(diff) == 0) return 0;
    return mrc_compare<G, R>(C, L, ax, ay);
}

// Largest magnitude of x[0 .. n): every lane group walks a strided share and keeps its best element; the groups' candidates go to
// `parts` (AoS records), a second launch over the candidates writes r[0] (sign cleared, src/mpreduct.cuh:247-250).  The interval
// fields of an element are looked at first: its digits are only fetched when it can beat the current best.
template <int G, int R>
__global__ void __launch_bounds__(256) k_maxabs(const DevConsts *Cp, long long n, SoA x, int incx, const char *recs_in, char *parts, SoA r) {
    const DevConsts &C = *Cp;
    Lane<R> L;
    lane_init<G, R>(C, L);
    const long long grp = ((long long) blockIdx.x * blockDim.x + threadIdx.x) / G;
    const long long ngrp = (long long) gridDim.x * blockDim.x / G;
    Num<R> best;
    num_zero(best);
    const long long lenx = recs_in ? 0 : x.len();
    for (long long i = grp; i < n; i += ngrp) {
        Num<R> e;
        if (recs_in) {
            load_rec<G, R>(C, L, recs_in, i, e);
        } else {
            const long long ix = ops_inc_index(i, n, incx);
            e.up = x.eval[ix + lenx];
            if (e.up.frac == 0) continue;                       // an exact zero never beats anything
            e.lo = x.eval[ix];
            e.exp = x.exp[ix];
            // |e| <= up 2^exp M and |best| >= lo 2^exp M: skip without touching the digits when that already decides it
            if (best.up.frac != 0) {
                Er eu = e.up, bl = best.lo;
                eu.exp += e.exp; bl.exp += best.exp;
                if (bl.frac != 0 && er_ucmp_dev(bl, eu) > 0) continue;
            }
            e.sign = x.sign[ix];
            load_digits<G, R>(C, L, x.digits, ix, e.d);
        }
        if (cmp_abs<G, R>(C, L, e, best) == 1) best = e;
    }
    best.sign = 0;
    if (parts) store_rec<G, R>(C, L, parts, grp, best);
    else if (grp == 0) store_num<G, R>(C, L, r, 0, best);
}

// ---- y = A x, A sparse (CSR / ELLPACK) with multiple-precision entries in an mp_collection_t ------------------------------------
// The reference materialises the nnz products in a buffer (two kernels), rounds them (third) and sums every row with one THREAD
// (fourth).  Here one lane group owns a row and walks its entries in the reference's order: round(a x_j), then sum = round(sum +
// product) starting from MP_ZERO -- the same sequence of mp_mul / mp_add, so the same bits, in one pass and without the buffer.
template <int G, int R, bool ELL>
__global__ void __launch_bounds__(128) k_spmv_2st(const DevConsts *Cp, int m, int width, const int *ptr, const int *ja, SoA as, SoA x, SoA y) {
    const DevConsts &C = *Cp;
    Lane<R> L;
    lane_init<G, R>(C, L);
    const long long grp = ((long long) blockIdx.x * blockDim.x + threadIdx.x) / G;
    const long long ngrp = (long long) gridDim.x * blockDim.x / G;
    for (long long row = grp; row < m; row += ngrp) {
        Num<R> sum, a, b, p;
        num_zero(sum);
        const int beg = ELL ? 0 : ptr[row], end = ELL ? width : ptr[row + 1];
        for (int t = beg; t < end; ++t) {
            const long long idx = ELL ? (long long) t * m + row : t;       // ELLPACK: column-major m x maxnzr, padding marked by ja < 0
            const int col = ja[idx];
            if (ELL && col < 0) continue;
            load_num<G, R>(C, L, as, idx, a);
            load_num<G, R>(C, L, x, col, b);
            mp_mul<G, R, true>(C, L, p, a, b);
            mp_add<G, R, true>(C, L, sum, sum, p);
        }
        store_num<G, R>(C, L, y, row, sum);
    }
}

// ---- double -> multiple precision (src/arith/assign.cuh:54-81) ---------------------------------------------------------------------
template <int G, int R>
__global__ void __launch_bounds__(128) k_set_d(const DevConsts *Cp, long long n, const double *src, SoA dst, long long offset) {
    const DevConsts &C = *Cp;
    Lane<R> L;
    lane_init<G, R>(C, L);
    const long long grp = ((long long) blockIdx.x * blockDim.x + threadIdx.x) / G;
    const long long ngrp = (long long) gridDim.x * blockDim.x / G;
    for (long long i = grp; i < n; i += ngrp) {
        const unsigned long long u = (unsigned long long) __double_as_longlong(src[i]);
        const int bexp = (int) ((u >> 52) & 0x7ff);
        unsigned long long sig = u & 0xfffffffffffffull;
        if (bexp) sig |= 1ull << 52;
        int ex = bexp - 1023 - 52;
        if (sig) { const int tz = __ffsll((long long) sig) - 1; sig >>= tz; ex += tz; } else { ex = 0; }
        Num<R> v;
        v.sign = (int) (u >> 63);
        v.exp = ex;
#pragma unroll
        for (int q = 0; q < R; ++q) v.d[q] = L.act[q] ? (int) (sig % (unsigned long long) (unsigned) L.m[q]) : 0;
        eval_compute<G, R, false>(C, L, v.d, v.lo, v.up);
        round_if_needed<G, R>(C, L, v);
        store_num<G, R>(C, L, dst, offset + i, v);
    }
}

// ---- multiple precision -> double (src/arith/assign.cuh:154-180: the exact value rounded to nearest) ------------------------------
// One thread per element: the significand is rebuilt in binary by the Chinese remainder theorem over all N moduli,
//      X = sum_i xi_i (M / m_i) - R M,   xi_i = x_i w_i mod m_i,   R = floor(sum xi_i / m_i)
// (R from a double sum; off by one at most, seen as X outside [0, M) and put right), then the leading 53 bits are rounded to nearest even.
struct GetDTab { int nw; const unsigned *mi; const unsigned *negm; const unsigned *mw; };     // mi [N][nw], negm = 2^(32 nw) - M, mw = M
__global__ void __launch_bounds__(128) k_get_d(const DevConsts *Cp, GetDTab T, long long n, SoA src, long long offset, double *dst) {
    const DevConsts &C = *Cp;
    const int N = C.N, nw = T.nw;
    unsigned x[kMaxN + 2];
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x) {
        const long long idx = offset + i;
        const int *dg = src.digits + idx * N;
        double sum = 0.0;
        for (int w = 0; w < nw; ++w) x[w] = 0;
        // column sums with a running carry: xi_i < 2^31, words < 2^32, N <= 128 terms: below 2^70 -- two 64-bit halves
        unsigned long long carry_lo = 0, carry_hi = 0;
        unsigned xi_cache = 0;
        (void) xi_cache;
        // xi_i are recomputed per word to keep the register footprint small (N mulmods per word; the conversion is not a hot path)
        for (int q = 0; q < N; ++q) sum += (double) mulmod(dg[q], C.part_inverse[q], C.moduli[q], C.barrett[q]) / (double) C.moduli[q];
        long long Rk = (long long) floor(sum);
        for (int pass = 0; pass < 3; ++pass) {
            carry_lo = 0; carry_hi = 0;
            for (int w = 0; w < nw; ++w) {
                unsigned long long lo = carry_lo, hi = carry_hi;
                for (int q = 0; q < N; ++q) {
                    const unsigned long long xi = (unsigned long long) (unsigned) mulmod(dg[q], C.part_inverse[q], C.moduli[q], C.barrett[q]);
                    const unsigned long long p = xi * T.mi[(size_t) q * nw + w];
                    lo += p; hi += lo < p ? 1ull : 0ull;
                }
                const unsigned long long p = (unsigned long long) Rk * T.negm[w];
                lo += p; hi += lo < p ? 1ull : 0ull;
                x[w] = (unsigned) lo;
                carry_lo = (lo >> 32) | (hi << 32); carry_hi = hi >> 32;
            }
            // the value modulo 2^(32 nw) is X + (R_true - Rk) M: in [0, M) when the rank was right
            bool ge = true;                                   // x >= M ?
            for (int w = nw - 1; w >= 0; --w) { if (x[w] != T.mw[w]) { ge = x[w] > T.mw[w]; break; } }
            if (!ge) break;
            if ((int) x[nw - 1] < 0) --Rk; else ++Rk;         // "negative": the rank was one too large; >= M: one too small
        }
        int Lb = 0;
        for (int w = 0; w < nw; ++w) if (x[w]) Lb = 32 * w + 32 - __clz(x[w]);
        double v = 0.0;
        if (Lb > 0) {
            // leading 64 bits (or all of them) and a sticky bit for the rest
            unsigned long long top = 0;
            bool sticky = false;
            const int sh = Lb - 64;                           // bits dropped below the window
            if (sh <= 0) {
                top = (unsigned long long) x[0] | (nw > 1 ? (unsigned long long) x[1] << 32 : 0ull);
            } else {
                const int ws = sh >> 5, bs = sh & 31;
                unsigned w0 = x[ws], w1 = ws + 1 < nw ? x[ws + 1] : 0u, w2 = ws + 2 < nw ? x[ws + 2] : 0u;
                const unsigned long long lo64 = (unsigned long long) w0 | ((unsigned long long) w1 << 32);
                top = bs ? (lo64 >> bs) | ((unsigned long long) w2 << (64 - bs)) : lo64;
                for (int w = 0; w < ws; ++w) sticky |= x[w] != 0;
                if (bs) sticky |= (w0 & ((1u << bs) - 1u)) != 0;
            }
            int e2 = sh > 0 ? sh : 0;
            // round `top` (up to 64 bits) to 53 bits, nearest even, with the sticky bit
            const int tl = 64 - __clzll((long long) top);
            if (tl > 53) {
                const int d = tl - 53;
                const unsigned long long rem = top & ((1ull << d) - 1ull), half = 1ull << (d - 1);
                unsigned long long mnt = top >> d;
                if (rem > half || (rem == half && (sticky || (mnt & 1ull)))) ++mnt;
                top = mnt; e2 += d;
            }
            v = scalbn((double) top, e2 + src.exp[idx]);
            if (src.sign[idx]) v = -v;
        }
        dst[i] = v;
    }
}

// ---- division (src/arith/div.cuh:33-66: the reference copies both numbers to the host and divides with MPFR at MP_PRECISION bits) ----------
// Here: one thread rebuilds both significands in binary (CRT over all moduli), divides by shift-and-subtract, rounds the quotient to nearest
// (ties to even, MPFR_RNDN) at `prec` bits and trims its trailing zeros as mp_set_mpfr does (assign.cuh:95-111); the lane group then forms
// the digits' interval evaluation.  r[0] = x[0] / y[0]; a zero divisor leaves r untouched and sets *status.
__device__ __forceinline__ void ops_crt(const DevConsts &C, const int *dg, const GetDTab &T, unsigned *x) {
    const int N = C.N, nw = T.nw;
    double sum = 0.0;
    for (int q = 0; q < N; ++q) sum += (double) mulmod(dg[q], C.part_inverse[q], C.moduli[q], C.barrett[q]) / (double) C.moduli[q];
    long long Rk = (long long) floor(sum);
    for (int pass = 0; pass < 3; ++pass) {
        unsigned long long clo = 0, chi = 0;
        for (int w = 0; w < nw; ++w) {
            unsigned long long lo = clo, hi = chi;
            for (int q = 0; q < N; ++q) {
                const unsigned long long xi = (unsigned long long) (unsigned) mulmod(dg[q], C.part_inverse[q], C.moduli[q], C.barrett[q]);
                const unsigned long long p = xi * T.mi[(size_t) q * nw + w];
                lo += p; hi += lo < p ? 1ull : 0ull;
            }
            const unsigned long long p = (unsigned long long) Rk * T.negm[w];
            lo += p; hi += lo < p ? 1ull : 0ull;
            x[w] = (unsigned) lo;
            clo = (lo >> 32) | (hi << 32); chi = hi >> 32;
        }
        bool ge = true;
        for (int w = nw - 1; w >= 0; --w) { if (x[w] != T.mw[w]) { ge = x[w] > T.mw[w]; break; } }
        if (!ge) break;
        if ((int) x[nw - 1] < 0) --Rk; else ++Rk;
    }
}
__device__ __forceinline__ int ops_bitlen(const unsigned *x, int nw) {
    int L = 0;
    for (int w = 0; w < nw; ++w) if (x[w]) L = 32 * w + 32 - __clz(x[w]);
    return L;
}
constexpr int kDivW = 2 * (kMaxN + 2) + 2;
template <int G, int R>
__global__ void k_div(const DevConsts *Cp, GetDTab T, SoA xs, SoA ys, SoA rs, int *status) {
    const DevConsts &C = *Cp;
    Lane<R> L;
    lane_init<G, R>(C, L);
    __shared__ unsigned q[kMaxN + 4];
    __shared__ int s_exp, s_sign, s_nw, s_ok;
    if (threadIdx.x == 0) {
        s_ok = 0;
        const bool xz = xs.eval[xs.len()].frac == 0, yz = ys.eval[ys.len()].frac == 0;
        if (yz) { if (status) *status = 1; }
        else if (xz) { s_ok = 2; }
        else {
            unsigned X[kMaxN + 2], Y[kMaxN + 2], Nn[kDivW], Rm[kMaxN + 3];
            const int nw = T.nw, prec = C.precision;
            ops_crt(C, xs.digits, T, X);
            ops_crt(C, ys.digits, T, Y);
            const int Lx = ops_bitlen(X, nw), Ly = ops_bitlen(Y, nw);
            const int s = prec + 2 + Ly - Lx;                       // the quotient of X 2^s / Y has prec + 2 or prec + 3 bits
            // numerator X 2^max(s, 0), divisor Y 2^max(-s, 0) (Ly + |s| bits: at most 2 nw words)
            const int nn = 2 * nw + 2;
            for (int w = 0; w < nn; ++w) Nn[w] = 0u;
            const int sx = s > 0 ? s : 0;
            const int nxw = (Lx + 31) >> 5, nyw = (Ly + 31) >> 5;
            for (int w = 0; w < nxw; ++w) {
                const unsigned long long v = (unsigned long long) X[w] << (sx & 31);
                Nn[w + (sx >> 5)] |= (unsigned) v;
                Nn[w + (sx >> 5) + 1] |= (unsigned) (v >> 32);
            }
            unsigned D[kMaxN + 4];
            const int nd = nw + 2;                                    // s < 0 only when Lx > prec + 2 + Ly: |s| < Lx: fits
            for (int w = 0; w < nd; ++w) D[w] = 0u;
            const int sy = s < 0 ? -s : 0;
            {
                for (int w = 0; w < nyw; ++w) {
                    const unsigned long long v = (unsigned long long) Y[w] << (sy & 31);
                    D[w + (sy >> 5)] |= (unsigned) v;
                    D[w + (sy >> 5) + 1] |= (unsigned) (v >> 32);
                }
                const int Ln = ops_bitlen(Nn, nn);
                for (int w = 0; w < nd + 1; ++w) Rm[w] = 0u;
                for (int w = 0; w < nw + 2; ++w) q[w] = 0u;
                for (int b = Ln - 1; b >= 0; --b) {
                    // R = 2 R + bit b of the numerator
                    unsigned cy = (Nn[b >> 5] >> (b & 31)) & 1u;
                    for (int w = 0; w < nd + 1; ++w) { const unsigned nx = Rm[w] >> 31; Rm[w] = (Rm[w] << 1) | cy; cy = nx; }
                    bool ge = Rm[nd] != 0;
                    if (!ge) { ge = true; for (int w = nd - 1; w >= 0; --w) { if (Rm[w] != D[w]) { ge = Rm[w] > D[w]; break; } } }
                    if (ge) {
                        long long bw = 0;
                        for (int w = 0; w < nd; ++w) { const long long df = (long long) Rm[w] - D[w] - bw; Rm[w] = (unsigned) df; bw = df < 0 ? 1 : 0; }
                        Rm[nd] -= (unsigned) bw;
                        if (b < 32 * (nw + 2)) q[b >> 5] |= 1u << (b & 31);
                    }
                }
                bool sticky = false;
                for (int w = 0; w < nd + 1; ++w) sticky |= Rm[w] != 0;
                int Lq = ops_bitlen(q, nw + 2);
                int ex = xs.exp[0] - ys.exp[0] - s;
                const int d = Lq - prec;
                if (d > 0) {
                    const int hw = (d - 1) >> 5;
                    const unsigned hb = 1u << ((d - 1) & 31);
                    // remainder below the cut against one half
                    bool above = false, half = (q[hw] & hb) != 0;
                    for (int w = 0; w <= hw; ++w) {
                        const unsigned mask = w < hw ? 0xffffffffu : (hb - 1u);
                        if (q[w] & mask) above = true;
                    }
                    // shift right by d
                    const int ws = d >> 5, bs = d & 31;
                    for (int w = 0; w < nw + 2; ++w) {
                        const unsigned lo = w + ws < nw + 2 ? q[w + ws] : 0u, hi = w + ws + 1 < nw + 2 ? q[w + ws + 1] : 0u;
                        q[w] = bs ? (lo >> bs) | (hi << (32 - bs)) : lo;
                    }
                    if (half && (above || sticky || (q[0] & 1u))) {
                        unsigned long long cy2 = 1;
                        for (int w = 0; w < nw + 2 && cy2; ++w) { cy2 += q[w]; q[w] = (unsigned) cy2; cy2 >>= 32; }
                    }
                    ex += d;
                }
                // trailing zeros into the exponent
                Lq = ops_bitlen(q, nw + 2);
                int tz = 0;
                while (tz < Lq && !((q[tz >> 5] >> (tz & 31)) & 1u)) ++tz;
                if (tz) {
                    const int ws = tz >> 5, bs = tz & 31;
                    for (int w = 0; w < nw + 2; ++w) {
                        const unsigned lo = w + ws < nw + 2 ? q[w + ws] : 0u, hi = w + ws + 1 < nw + 2 ? q[w + ws + 1] : 0u;
                        q[w] = bs ? (lo >> bs) | (hi << (32 - bs)) : lo;
                    }
                    ex += tz;
                }
                s_exp = ex; s_sign = (xs.sign[0] ^ ys.sign[0]) & 1; s_nw = nw + 2; s_ok = 1;
            }
        }
    }
    __syncthreads();
    if (threadIdx.x >= G) return;
    Num<R> v;
    if (s_ok == 2) { num_zero(v); store_num<G, R>(C, L, rs, 0, v); return; }
    if (s_ok != 1) return;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        unsigned long long acc = 0;
        if (L.act[r]) {
            for (int w = 0; w < s_nw; ++w) {
                if (q[w] == 0u) continue;
                const int pw = 32 * w <= C.log2M ? __ldg(C.pow2 + (long long) (32 * w) * C.N + L.idx[r]) : pow2_slow(32 * w, L.m[r], L.mu[r]);
                acc += (unsigned long long) q[w] * (unsigned) pw;
                acc = (unsigned long long) (unsigned) reduce64(acc, L.m[r], L.mu[r]);
            }
        }
        v.d[r] = (int) acc;
    }
    v.sign = s_sign; v.exp = s_exp;
    eval_compute<G, R, false>(C, L, v.d, v.lo, v.up);
    store_num<G, R>(C, L, rs, 0, v);
}

// r = a x + y element-wise with the roundings of mp_mul / mp_add (cuda::mp_axpy of src/blas/v2/axpy_v2.cuh; negate: cuda::mp_maxpy, maxpy_v2.cuh:39-53)
template <int G, int R>
__global__ void k_axpy_out(const DevConsts *Cp, long long n, SoA a, bool negate, SoA x, SoA y, SoA r) {
    const DevConsts &C = *Cp;
    Lane<R> L;
    lane_init<G, R>(C, L);
    const long long ngrp = (long long) gridDim.x * blockDim.x / G;
    Num<R> s, xv, yv, t;
    load_num<G, R>(C, L, a, 0, s);
    if (negate) s.sign ^= 1;
    for (long long i = ((long long) blockIdx.x * blockDim.x + threadIdx.x) / G; i < n; i += ngrp) {
        load_num<G, R>(C, L, x, i, xv);
        load_num<G, R>(C, L, y, i, yv);
        mp_mul<G, R, true>(C, L, t, s, xv);
        mp_add<G, R, true>(C, L, t, t, yv);
        store_num<G, R>(C, L, r, i, t);
    }
}
// r = x - y (cuda::mp_diff, src/blas/v2/diff_v2.cuh) and r = x * y element-wise (cuda::mp_prod_d with the doubles converted once)
template <int G, int R, bool MUL>
__global__ void k_elementwise(const DevConsts *Cp, long long n, SoA x, SoA y, SoA r) {
    const DevConsts &C = *Cp;
    Lane<R> L;
    lane_init<G, R>(C, L);
    const long long ngrp = (long long) gridDim.x * blockDim.x / G;
    Num<R> xv, yv, t;
    for (long long i = ((long long) blockIdx.x * blockDim.x + threadIdx.x) / G; i < n; i += ngrp) {
        load_num<G, R>(C, L, x, i, xv);
        load_num<G, R>(C, L, y, i, yv);
        if (MUL) mp_mul<G, R, true>(C, L, t, xv, yv);
        else { yv.sign ^= 1; mp_add<G, R, true>(C, L, t, xv, yv); }
        store_num<G, R>(C, L, r, i, t);
    }
}
template <int G, int R>
__global__ void k_copy(const DevConsts *Cp, long long n, SoA x, SoA r) {
    const DevConsts &C = *Cp;
    Lane<R> L;
    lane_init<G, R>(C, L);
    const long long ngrp = (long long) gridDim.x * blockDim.x / G;
    Num<R> xv;
    for (long long i = ((long long) blockIdx.x * blockDim.x + threadIdx.x) / G; i < n; i += ngrp) {
        load_num<G, R>(C, L, x, i, xv);
        store_num<G, R>(C, L, r, i, xv);
    }
}

}  // namespace mpres

namespace {

struct OpsTables {            // per context, built at first use
    unsigned *d_mi = nullptr, *d_negm = nullptr, *d_mw = nullptr;
    int nw = 0;
};
std::mutex g_ops_mu;
std::vector<std::pair<mpres_ctx *, OpsTables>> g_ops;

int ops_tables(mpres_ctx *c, OpsTables *out) {
    std::lock_guard<std::mutex> lk(g_ops_mu);
    for (auto &e : g_ops) if (e.first == c) { *out = e.second; return 0; }
    const int N = c->hc.N;
    BigUInt M(1);
    for (int i = 0; i < N; ++i) M.mul_small((uint32_t) c->hc.moduli[i]);
    OpsTables t;
    t.nw = (int) M.limb.size() + 1;                      // one spare word: sum xi_i (M/m_i) < N M
    std::vector<unsigned> mi((size_t) N * t.nw, 0), negm(t.nw, 0), mw(t.nw, 0);
    for (int w = 0; w < (int) M.limb.size(); ++w) mw[w] = M.limb[w];
    unsigned long long borrow = 0;
    for (int w = 0; w < t.nw; ++w) {
        const unsigned long long sub = (unsigned long long) mw[w] + borrow;
        negm[w] = (unsigned) (0ull - sub);
        borrow = sub != 0 ? 1 : 0;
    }
    for (int i = 0; i < N; ++i) {
        BigUInt q = M;
        q.div_small((uint32_t) c->hc.moduli[i]);
        for (int w = 0; w < (int) q.limb.size(); ++w) mi[(size_t) i * t.nw + w] = q.limb[w];
    }
    CUDA_TRY(cudaMalloc(&t.d_mi, mi.size() * 4));
    CUDA_TRY(cudaMalloc(&t.d_negm, negm.size() * 4));
    CUDA_TRY(cudaMalloc(&t.d_mw, mw.size() * 4));
    CUDA_TRY(cudaMemcpy(t.d_mi, mi.data(), mi.size() * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(t.d_negm, negm.data(), negm.size() * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(t.d_mw, mw.data(), mw.size() * 4, cudaMemcpyHostToDevice));
    g_ops.emplace_back(c, t);
    *out = t;
    return 0;
}

}  // namespace

extern "C" {

// called by mpres_finalize (mpres_b200.cu)
void mpres_ops_release(mpres_ctx *c) {
    std::lock_guard<std::mutex> lk(g_ops_mu);
    for (size_t i = 0; i < g_ops.size(); ++i)
        if (g_ops[i].first == c) {
            cudaFree(g_ops[i].second.d_mi); cudaFree(g_ops[i].second.d_negm); cudaFree(g_ops[i].second.d_mw);
            g_ops.erase(g_ops.begin() + (long) i);
            return;
        }
}

// r[0] = the element of x[0 .. n) (stride incx > 0) of largest magnitude, sign cleared.  The caller holds the context's lock.
int mpres_internal_maxabs(mpres_ctx *c, long long n, const SoA *x, int incx, const SoA *r, cudaStream_t st) {
    const int N = c->hc.N;
    const size_t rs = 4 * (size_t) N + 40;
    int rc = 0;
    MPRES_DISPATCH(N, {
        const long long gpb = 256 / G;
        const long long blocks = std::max<long long>(1, std::min<long long>((n + gpb - 1) / gpb, (long long) c->sm_count * 4));
        const long long groups = blocks * gpb;
        void *parts;
        rc = ws_reserve(c, 18, (size_t) (groups + gpb) * rs, &parts);
        if (rc) return rc;
        char *p1 = (char *) parts, *p2 = p1 + (size_t) groups * rs;
        // every group keeps the best of its share; one block folds the groups' candidates; one group folds that block's
        k_maxabs<G, R><<<(unsigned) blocks, 256, 0, st>>>(c->dconsts, n, *x, incx, nullptr, p1, *r);
        k_maxabs<G, R><<<1, 256, 0, st>>>(c->dconsts, groups, *x, 1, p1, p2, *r);
        k_maxabs<G, R><<<1, G, 0, st>>>(c->dconsts, gpb, *x, 1, p2, nullptr, *r);
    });
    for (int i = 0; i < 3; ++i) LAUNCHED(c);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int mpres_norm(mpres_ctx *c, int norm, int n, const mpres_array_t *x, int incx, mpres_array_t *r, mpres_stream_t stream) {
    NEED_DEVICE(c);
    if (!c || !x || !r) return -1;
    if (norm != MPRES_ONE_NORM && norm != MPRES_INF_NORM) return -2;
    if (norm == MPRES_ONE_NORM) return mpres_asum(c, n, x, incx, r, stream);
    if (n <= 0 || incx <= 0) return 0;                       // src/blas/norm.cuh:52-55
    DeviceGuard g(c->device);
    std::lock_guard<std::mutex> lk(c->mu);
    cudaStream_t st = (cudaStream_t) stream;
    int rc = call_begin(c, st);
    if (rc) return rc;
    c->last_stream = st;
    const SoA xs = view(x), rv = view(r);
    if ((rc = mpres_internal_maxabs(c, (long long) n, &xs, incx, &rv, st))) return rc;
    return call_end(c, st);
}

static int spmv_impl(mpres_ctx *c, bool ell, int m, int n, int width, const int *ptr, const int *ja, const mpres_collection_t *as, size_t len_as,
                     const mpres_array_t *x, mpres_array_t *y, cudaStream_t st) {
    NEED_DEVICE(c);
    if (!c || !ja || !as || !x || !y || (!ell && !ptr)) return -1;
    if (m <= 0 || n <= 0) return 0;
    DeviceGuard g(c->device);
    std::lock_guard<std::mutex> lk(c->mu);
    int rc = call_begin(c, st);
    if (rc) return rc;
    c->last_stream = st;
    const int N = c->hc.N;
    MPRES_DISPATCH(N, {
        const long long gpb = 128 / G;
        const long long blocks = std::max<long long>(1, std::min<long long>(((long long) m + gpb - 1) / gpb, (long long) c->sm_count * 16));
        if (ell) k_spmv_2st<G, R, true><<<(unsigned) blocks, 128, 0, st>>>(c->dconsts, m, width, ptr, ja, view(as, len_as), view(x), view(y));
        else k_spmv_2st<G, R, false><<<(unsigned) blocks, 128, 0, st>>>(c->dconsts, m, width, ptr, ja, view(as, len_as), view(x), view(y));
    });
    LAUNCHED(c);
    CUDA_TRY(cudaGetLastError());
    return call_end(c, st);
}

int mpres_spmv_csr2st(mpres_ctx *c, int m, int n, int nnz, const int *irp, const int *ja, const mpres_collection_t *as, const mpres_array_t *x,
                      mpres_array_t *y, mpres_collection_t *buffer, mpres_stream_t stream) {
    (void) buffer;                                             // the reference's nnz-element scratch is not needed
    if (nnz < 0) return -2;
    return spmv_impl(c, false, m, n, 0, irp, ja, as, (size_t) nnz, x, y, (cudaStream_t) stream);
}

int mpres_spmv_ell2st(mpres_ctx *c, int m, int n, int maxnzr, const int *ja, const mpres_collection_t *as, const mpres_array_t *x, mpres_array_t *y,
                      mpres_collection_t *buffer, mpres_stream_t stream) {
    (void) buffer;
    if (maxnzr < 0) return -2;
    return spmv_impl(c, true, m, n, maxnzr, nullptr, ja, as, (size_t) m * (size_t) maxnzr, x, y, (cudaStream_t) stream);
}

int mpres_array_set_d(mpres_ctx *c, mpres_array_t *dst, size_t offset, const double *src, size_t n, mpres_stream_t stream) {
    NEED_DEVICE(c);
    if (!c || !dst || !src) return -1;
    if (n == 0) return 0;
    DeviceGuard g(c->device);
    cudaStream_t st = (cudaStream_t) stream;
    MPRES_DISPATCH(c->hc.N, {
        const long long gpb = 128 / G;
        const long long blocks = std::min<long long>(((long long) n + gpb - 1) / gpb, (long long) c->sm_count * 16);
        k_set_d<G, R><<<(unsigned) blocks, 128, 0, st>>>(c->dconsts, (long long) n, src, view(dst), (long long) offset);
    });
    LAUNCHED(c);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

static int ops_gettab(mpres_ctx *c, GetDTab *T) {
    OpsTables t;
    int rc = ops_tables(c, &t);
    if (rc) return rc;
    if (t.nw > kMaxN + 2) return -4;
    T->nw = t.nw; T->mi = t.d_mi; T->negm = t.d_negm; T->mw = t.d_mw;
    return 0;
}

int mpres_div(mpres_ctx *c, mpres_array_t *r, const mpres_array_t *x, const mpres_array_t *y, mpres_stream_t stream) {
    NEED_DEVICE(c);
    if (!c || !r || !x || !y) return -1;
    DeviceGuard g(c->device);
    GetDTab T;
    int rc = ops_gettab(c, &T);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t) stream;
    MPRES_DISPATCH(c->hc.N, { k_div<G, R><<<1, 32, 0, st>>>(c->dconsts, T, view(x), view(y), view(r), nullptr); });
    LAUNCHED(c);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// Conjugate gradients, optionally with a diagonal preconditioner (src/sparse/solver/cg_csr.cuh:53-110, pcg_csr.cuh:57-130), built from this
// library's operations: two-stage SpMV over the matrix converted once to multiple precision, mpres_dot, division, fused a x + y.
// A: CSR with double entries (irp, ja, as: device pointers, as in the reference's csr_t), b, x: mp_array_t of n elements (x: initial guess in,
// solution out), M: n doubles (device; the inverse diagonal) or NULL, tol: relative residual, resvec: maxit + 1 doubles on the host or NULL.
int mpres_cg_csr(mpres_ctx *c, int n, int nnz, const int *irp, const int *ja, const double *as, const mpres_array_t *b, double tol, int maxit, const double *M,
                 mpres_array_t *x, int *iters, double *resvec, mpres_stream_t stream) {
    NEED_DEVICE(c);
    if (!c || !irp || !ja || !as || !b || !x || n <= 0 || nnz < 0) return -1;
    DeviceGuard g(c->device);
    cudaStream_t st = (cudaStream_t) stream;
    const int N = c->hc.N;
    mpres_array_t r, p, q, z, sc, Mv;                     // sc: rho, rhop, alpha, beta, pq, nrm as six one-element views of one array
    mpres_collection_t Am;
    memset(&r, 0, sizeof(r)); memset(&p, 0, sizeof(p)); memset(&q, 0, sizeof(q)); memset(&z, 0, sizeof(z)); memset(&sc, 0, sizeof(sc)); memset(&Mv, 0, sizeof(Mv));
    memset(&Am, 0, sizeof(Am));
    double *d_nrm = nullptr;
    auto cleanup = [&]() {                                   // every exit: nothing of this call stays allocated (clearing a never-allocated array is a no-op)
        cudaStreamSynchronize(st);
        mpres_array_clear(c, &r); mpres_array_clear(c, &p); mpres_array_clear(c, &q); mpres_array_clear(c, &sc); mpres_collection_clear(c, &Am);
        mpres_array_clear(c, &z); mpres_array_clear(c, &Mv);
        cudaFree(d_nrm);
    };
    int rc;
    if ((rc = mpres_array_init(c, &r, n)) || (rc = mpres_array_init(c, &p, n)) || (rc = mpres_array_init(c, &q, n)) || (rc = mpres_array_init(c, &sc, 6)) ||
        (rc = mpres_collection_init(c, &Am, nnz ? nnz : 1)) || (M && ((rc = mpres_array_init(c, &z, n)) || (rc = mpres_array_init(c, &Mv, n))))) {
        cleanup();
        return rc;
    }
    {
        const cudaError_t e = cudaMalloc(&d_nrm, sizeof(double));
        if (e != cudaSuccess) { d_nrm = nullptr; cudaGetLastError(); cleanup(); return (int) e; }
    }
    auto scalar = [&](int i) {                               // element i of sc as a length-1 array (same allocated length: same offset of the upper bounds)
        mpres_array_t v = sc;
        v.digits += (size_t) i * N; v.sign += i; v.exp += i; v.eval += i;
        return v;
    };
    mpres_array_t rho = scalar(0), rhop = scalar(1), alpha = scalar(2), beta = scalar(3), pq = scalar(4), nrm = scalar(5);
    auto grid = [&](long long items, int G_) { return (unsigned) std::max<long long>(1, std::min<long long>((items * G_ + 127) / 128, (long long) c->sm_count * 16)); };
    // the matrix (and the preconditioner) in multiple precision, exactly
    {
        mpres_array_t av;                                    // the collection's arrays as an array view for the conversion kernel
        memset(&av, 0, sizeof(av));
        av.digits = Am.digits; av.sign = Am.sign; av.exp = Am.exp; av.eval = Am.eval;
        SoA dstv = view(&Am, (size_t) (nnz ? nnz : 1));
        if (nnz) { MPRES_DISPATCH(N, { k_set_d<G, R><<<grid(nnz, G), 128, 0, st>>>(c->dconsts, (long long) nnz, as, dstv, 0); }); LAUNCHED(c); }
        if (M && (rc = mpres_array_set_d(c, &Mv, 0, M, (size_t) n, stream))) { cleanup(); return rc; }
    }
    auto norm2 = [&](const mpres_array_t *v, double *out) -> int {   // sqrt(v . v) as a double (cuda::mp_norm2, src/blas/v2/norm2_v2.cuh:108-118)
        int e = mpres_dot(c, n, v, 1, v, 1, &nrm, nullptr, stream);
        if (e) return e;
        if ((e = mpres_array_get_d(c, d_nrm, &nrm, 0, 1, stream))) return e;
        double h = 0;
        CUDA_TRY(cudaMemcpyAsync(&h, d_nrm, sizeof(double), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        *out = sqrt(h);
        return 0;
    };
    auto fail = [&](int e) { cleanup(); return e; };
    // r = b - A x
    if ((rc = mpres_spmv_csr2st(c, n, n, nnz, irp, ja, &Am, x, &r, nullptr, stream))) return fail(rc);
    MPRES_DISPATCH(N, { k_elementwise<G, R, false><<<grid(n, G), 128, 0, st>>>(c->dconsts, (long long) n, view(b), view(&r), view(&r)); });
    LAUNCHED(c);
    double norm0 = 0, normk = 0;
    if ((rc = norm2(&r, &norm0))) return fail(rc);
    const double eps = norm0 * tol;
    normk = norm0;
    int k = 0;
    if (resvec) resvec[0] = norm0 > 0 ? 1.0 : 0.0;
    while (normk > eps && k < maxit) {
        const mpres_array_t *zz = &r;
        if (M) {
            MPRES_DISPATCH(N, { k_elementwise<G, R, true><<<grid(n, G), 128, 0, st>>>(c->dconsts, (long long) n, view(&r), view(&Mv), view(&z)); });
            LAUNCHED(c);
            zz = &z;
        }
        MPRES_DISPATCH(N, { k_copy<G, R><<<1, 32, 0, st>>>(c->dconsts, 1, view(&rho), view(&rhop)); });
        if ((rc = mpres_dot(c, n, &r, 1, zz, 1, &rho, nullptr, stream))) return fail(rc);
        if (k == 0) {
            MPRES_DISPATCH(N, { k_copy<G, R><<<grid(n, G), 128, 0, st>>>(c->dconsts, (long long) n, view(zz), view(&p)); });
        } else {
            if ((rc = mpres_div(c, &beta, &rho, &rhop, stream))) return fail(rc);
            MPRES_DISPATCH(N, { k_axpy_out<G, R><<<grid(n, G), 128, 0, st>>>(c->dconsts, (long long) n, view(&beta), false, view(&p), view(zz), view(&p)); });
        }
        LAUNCHED(c);
        if ((rc = mpres_spmv_csr2st(c, n, n, nnz, irp, ja, &Am, &p, &q, nullptr, stream))) return fail(rc);
        if ((rc = mpres_dot(c, n, &p, 1, &q, 1, &pq, nullptr, stream))) return fail(rc);
        if ((rc = mpres_div(c, &alpha, &rho, &pq, stream))) return fail(rc);
        MPRES_DISPATCH(N, {
            k_axpy_out<G, R><<<grid(n, G), 128, 0, st>>>(c->dconsts, (long long) n, view(&alpha), false, view(&p), view(x), view(x));
            k_axpy_out<G, R><<<grid(n, G), 128, 0, st>>>(c->dconsts, (long long) n, view(&alpha), true, view(&q), view(&r), view(&r));
        });
        LAUNCHED(c); LAUNCHED(c);
        if ((rc = norm2(&r, &normk))) return fail(rc);
        ++k;
        if (resvec) resvec[k] = norm0 > 0 ? normk / norm0 : 0.0;
    }
    if (iters) *iters = k;
    const cudaError_t last = cudaGetLastError();
    cleanup();
    return (int) last;
}

int mpres_array_get_d(mpres_ctx *c, double *dst, const mpres_array_t *src, size_t offset, size_t n, mpres_stream_t stream) {
    NEED_DEVICE(c);
    if (!c || !dst || !src) return -1;
    if (n == 0) return 0;
    DeviceGuard g(c->device);
    OpsTables t;
    int rc = ops_tables(c, &t);
    if (rc) return rc;
    if (t.nw > kMaxN + 2) return -4;
    cudaStream_t st = (cudaStream_t) stream;
    GetDTab T;
    T.nw = t.nw; T.mi = t.d_mi; T.negm = t.d_negm; T.mw = t.d_mw;
    const long long blocks = std::min<long long>(((long long) n + 127) / 128, (long long) c->sm_count * 8);
    k_get_d<<<(unsigned) blocks, 128, 0, st>>>(c->dconsts, T, (long long) n, view(src), (long long) offset, dst);
    LAUNCHED(c);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

}  // extern "C"
