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
