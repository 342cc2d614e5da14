// kernels_fast.cuh -- the exact-window fast path of mp_gemm: three stages, all sm_100a kernels.
//
//  Stage 1  (alignment)      k_outer_info + k_align_planes + k_minplus
//      Per row i of op(A) / column j of op(B): the minimum exponent of its non-zero entries and the
//      magnitude window (from the interval evaluations).  Every entry is pre-shifted to that common
//      exponent, X' = +-X * 2^(e - e_min) mod m_q, and written as four unsigned 8-bit limb planes per
//      modulus, K-major.  A (min,+) product of the shift planes gives, per C entry, the exponent the
//      reference's sequential mp_add chain ends with (src/arith/add.cuh:172: exp = min of the term
//      exponents).
//  Stage 2  (per-modulus multiply-accumulate)      k_limb_gemm
//      C_q(i,j) = sum_l A'_q(i,l) * B'_q(l,j) mod m_q as 16 (limb x limb) int8 tensor-core GEMMs with
//      exact int32 accumulation, recombined with 2^(8u) mod m_q.  Measured on B200 (tools/mma_bench.cu):
//      IMMA.16832.U8 sustains 1959 int8 MAC/clk/SM = 122 residue-MAC/clk/SM against 29.4 for
//      IMAD.WIDE.U32, which is why the limb split is used (BASELINE north_star: "only if it beats the
//      INT32 IMAD pipe").
//  Stage 3  (rounding / normalisation)      k_normalize_epilogue
//      Sign from the interval evaluation of the accumulated residues (the exact sum S satisfies
//      |S| < M/4 by the window guard), exact division by the power of two that separates our row+column
//      exponent base from the reference's result exponent, interval evaluation, one rounding
//      (power-of-two scaling) if the significand exceeds the working precision, then the
//      alpha/beta epilogue of src/blas/gemm.cuh:142-166 fused in.
//
// Exactness.  With A'(i,l) = +-Xa*2^sa(i,l), B'(l,j) = +-Xb*2^sb(l,j):  sum_l A'B' is the EXACT
// integer sum_l +-XaXb 2^(e_l - ra_i - cb_j) as long as it stays below M/4 in magnitude; the guard
// ua_i + ub_j + ceil(log2 k) <= log2(M) - 2 uses per-row/column bit bounds.  Whenever the reference's
// own k-loop never rounds or drops a term (its benchmark inputs, SURVEY 7 hard part 1) the digits,
// sign and exponent produced here are bit-identical to it; otherwise the result is the exact sum
// rounded once, which is at least as accurate as the reference's k roundings.  Elements whose guard
// fails are recomputed in reference order (k_gemm_ref_order with a todo list).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <climits>
#include <type_traits>

#include "ctx.hpp"
#include "kernels_ref_order.cuh"
#include "kernels_umma.cuh"

namespace mpres {

constexpr int kShiftSentinel = 15000;  // shift-plane value of an exact zero (ignored by the (min,+) product)
constexpr int kShiftMax = 8000;        // larger shifts cannot pass the window guard for any supported M

struct OuterInfo {   // per row of op(A) / column of op(B)
    int emin;        // min exponent over non-zero entries (0 if none)
    int win;         // max over non-zero entries of (e - emin + bit bound of X); <0 if the line is all zero
    int xb;          // upper bound of 1024 log2(X) over the non-zero entries (significand size, for the small-modulus path)
    int pad;
};

// ---- stage 1a: exponent base and magnitude window of every line -----------------------------------
// element (o, l) lives at index o*so + l*sl of X.  One warp per line.
__global__ void k_outer_info(const DevConsts *Cp, SoA X, long long so, long long sl, int outer, int inner, OuterInfo *info) {
    const int log2M = Cp->log2M;
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= outer) return;
    const long long len = X.len();
    int emin = INT_MAX;
    long long top = LLONG_MIN;
    // largest upper bound of X / M over the line, compared without floating point: (binary exponent of the bound, fraction bits) is
    // monotone in the value for positive doubles; its log2 is taken once per line
    long long be = LLONG_MIN;
    unsigned long long bm = 0;
    for (int l = lane; l < inner; l += 32) {
        const long long idx = warp * so + l * sl;
        const Er up = X.eval[idx + len];
        if (up.frac != 0) {
            const int e = X.exp[idx];
            emin = min(emin, e);
            // X/M < 2^(up.exp+1) and M < 2^(log2M+1)  =>  X < 2^(log2M + up.exp + 2)
            long long t = (long long) e + up.exp;
            top = t > top ? t : top;
            const unsigned long long fb = (unsigned long long) __double_as_longlong(fabs(up.frac));
            const long long ue = (up.exp > 100000 ? 100000 : (up.exp < -100000 ? -100000 : up.exp)) + (long long) (fb >> 52);   // + biased exponent of frac
            const unsigned long long fm = fb & 0xfffffffffffffull;
            if (ue > be || (ue == be && fm > bm)) { be = ue; bm = fm; }
        }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        emin = min(emin, __shfl_xor_sync(0xffffffffu, emin, o));
        long long t = __shfl_xor_sync(0xffffffffu, top, o);
        top = t > top ? t : top;
        const long long oe = __shfl_xor_sync(0xffffffffu, be, o);
        const unsigned long long om = __shfl_xor_sync(0xffffffffu, bm, o);
        if (oe > be || (oe == be && om > bm)) { be = oe; bm = om; }
    }
    // log2 of the largest bound: (be - 1023) + log2(1.fraction)
    const double lx = (double) (be - 1023) + log2(__longlong_as_double((long long) (bm | 0x3ff0000000000000ull)));
    if (lane == 0) {
        OuterInfo r;
        r.pad = 0;
        if (emin == INT_MAX) { r.emin = 0; r.win = -1; r.xb = 0; }
        else {
            long long w = top - emin + log2M + 2;
            r.emin = emin;
            r.win = w > 1000000 ? 1000000 : (w < 0 ? 0 : (int) w);
            // X <= up * M:  1024 log2 X <= 1024 (lx + log2 M), rounded up with a margin
            const double b = (lx + (Cp->small ? Cp->small->log2M_up : (double) (log2M + 1))) * 1024.0;
            r.xb = b > 1.0e8 ? 100000000 : (b < 0 ? 0 : (int) ceil(b) + 2);
        }
        info[warp] = r;
    }
}

// ---- stage 1a': how many moduli the exact sums need ---------------------------------------------------
// Every exact sum satisfies |S| < 2^(win_a + win_b + ceil(log2 k)); the first n' moduli determine it when
// their product M' obeys |S| < M'/4.  n' = smallest such count (or N when more than kMaxReducedBase would be
// needed).  Stage 1b then aligns the first ceil4(n') moduli (it works on groups of four), stage 2 multiplies
// n' of them and k_base_extend reconstructs the residues q >= n'.  One block.
constexpr int kMaxReducedBase = 48;
// sel[0] = number of small moduli (0: the small-modulus path is not used), sel[1] = reference moduli its input conversion reads.
// When the small base is selected *nprime is 0 and the kernels of the reference-moduli path leave at once.
__global__ void __launch_bounds__(256) k_choose_base(const DevConsts *Cp, const OuterInfo *ia, int m, const OuterInfo *ib, int n, int k,
                                                     int enabled, int small_enabled, int *nprime, int *sel) {
    __shared__ int sa[256], sb[256], sx[256];
    int wa = -1, wb = -1, xb = 0;
    for (int i = threadIdx.x; i < m; i += 256) { wa = max(wa, ia[i].win); xb = max(xb, ia[i].xb); }
    for (int j = threadIdx.x; j < n; j += 256) { wb = max(wb, ib[j].win); xb = max(xb, ib[j].xb); }
    sa[threadIdx.x] = wa; sb[threadIdx.x] = wb; sx[threadIdx.x] = xb;
    __syncthreads();
    for (int o = 128; o >= 1; o >>= 1) {
        if (threadIdx.x < o) {
            sa[threadIdx.x] = max(sa[threadIdx.x], sa[threadIdx.x + o]); sb[threadIdx.x] = max(sb[threadIdx.x], sb[threadIdx.x + o]);
            sx[threadIdx.x] = max(sx[threadIdx.x], sx[threadIdx.x + o]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const int N = Cp->N;
        int np = N;
        int lgk = 0;
        while ((1 << lgk) < k) ++lgk;
        const long long need = (sa[0] < 0 || sb[0] < 0) ? 0 : (long long) sa[0] + sb[0] + lgk + 2;
        if (enabled && (N & 3) == 0) {
            for (int c = 1; c <= N; ++c)
                if ((long long) Cp->prefix_log2[c] >= need) { np = c; break; }
            if (np > kMaxReducedBase) np = N;
        }
        int P = 0, nin = 0;
        const SmallDev *SD = Cp->small;
        if (small_enabled && SD && SD->usable && sa[0] >= 0 && sb[0] >= 0 && sa[0] <= kSmallShiftMax && sb[0] <= kSmallShiftMax) {
            for (int c = 1; c <= kSmallMax; ++c)
                if ((long long) SD->prefix_log2[c] >= need) { P = c; break; }
            const int cmax = min(kSmallNinMax, N - 1);
            for (int c = 1; c <= cmax; ++c)
                if (SD->in_log2_milli[c] >= sx[0]) { nin = c; break; }   // X < m_0 ... m_{c-1}
            if (P == 0 || nin == 0) { P = 0; nin = 0; }
        }
        sel[0] = P; sel[1] = nin;
        *nprime = P > 0 ? 0 : np;
    }
}

// ---- stage 1b: pre-shifted signed residues as u8 limb planes + shift plane -------------------------
// planes: [q][limb][outer_p][inner_p] bytes (inner contiguous); shifts: [outer_p][inner_p] int16.
// One block handles one line `o` and a run of kRun inner positions.
constexpr int kRun = 128;
__global__ void __launch_bounds__(256) k_align_planes(const DevConsts *Cp, SoA X, long long so, long long sl, int outer, int inner,
                                                      const OuterInfo *info, uint8_t *planes, int16_t *shifts,
                                                      long long outer_p, long long inner_p) {
    extern __shared__ uint8_t sm_stage[];   // [N*4][kRun + 4]
    const DevConsts &C = *Cp;
    const int N = C.N;
    const int o = blockIdx.x;
    const int l0 = blockIdx.y * kRun;
    const int pitch = kRun + 4;
    const long long len = X.len();
    const bool line_ok = o < outer;
    const OuterInfo oi = line_ok ? info[o] : OuterInfo{0, -1};
    for (int t = threadIdx.x; t < kRun * N; t += blockDim.x) {
        const int ll = t / N, q = t - ll * N;
        const int l = l0 + ll;
        unsigned r = 0;
        if (line_ok && l < inner) {
            const long long idx = (long long) o * so + (long long) l * sl;
            const int d = X.digits[idx * N + q];
            const Er up = X.eval[idx + len];
            int s = 0;
            const bool nz = up.frac != 0;
            if (nz) {
                long long sh = (long long) X.exp[idx] - oi.emin;
                s = sh > kShiftMax ? kShiftMax : (int) sh;
                const int m = C.moduli[q];
                const unsigned long long mu = C.barrett[q];
                int p2 = s <= C.log2M ? C.pow2[(long long) s * N + q] : 0;   // s > log2M: line fails the guard anyway
                int v = mulmod(d, p2, m, mu);
                if (X.sign[idx] && v) v = m - v;
                r = (unsigned) v;
            }
            if (q == 0) shifts[(long long) o * inner_p + l] = (int16_t) (nz ? s : kShiftSentinel);
        } else if (q == 0 && l < inner_p) {
            shifts[(long long) o * inner_p + l] = (int16_t) kShiftSentinel;
        }
#pragma unroll
        for (int b = 0; b < 4; ++b) sm_stage[(q * 4 + b) * pitch + ll] = (uint8_t) (r >> (8 * b));
    }
    __syncthreads();
    // write out: row (q, limb) -> kRun contiguous bytes
    const int rows = N * 4;
    const int words = kRun / 4;
    for (int t = threadIdx.x; t < rows * words; t += blockDim.x) {
        const int row = t / words, w = t - row * words;
        const int l = l0 + w * 4;
        if (l < inner_p) {
            const uint8_t *src = sm_stage + row * pitch + w * 4;
            uint32_t v = (uint32_t) src[0] | ((uint32_t) src[1] << 8) | ((uint32_t) src[2] << 16) | ((uint32_t) src[3] << 24);
            *(uint32_t *) (planes + ((long long) row * outer_p + o) * inner_p + l) = v;
        }
    }
}

// ---- stage 1b, vectorised: four residues per work item ------------------------------------------------
// Same output as k_align_planes for N % 4 == 0.  A thread owns a fixed group of four moduli (constants in
// registers) and walks entries; the exponent / sign / interval fields are read once per four residues, digits
// and the 2^s table row as 128-bit loads.  Bytes are staged in shared memory as [limb][q][l] (pitch 33 words:
// the eight modulus groups of a warp hit eight different banks, the four entries of a warp the same word)
// and leave as full 128-byte rows.
__global__ void __launch_bounds__(256) k_align_planes4(const DevConsts *Cp, SoA X, long long so, long long sl, int outer, int inner,
                                                       const OuterInfo *info, uint8_t *planes, int16_t *shifts,
                                                       long long outer_p, long long inner_p, const int *nprime) {
    extern __shared__ uint8_t sm_stage[];   // [4][N][kRun + 4]
    const DevConsts &C = *Cp;
    const int N = C.N;
    if (*nprime <= 0) return;                    // the small-modulus path was selected
    const int np = min(N, (*nprime + 3) & ~3);   // moduli q >= np are not needed (reduced base)
    const int Q4 = np >> 2;                 // active modulus groups
    const int EP = 256 / Q4;                // entries per pass
    const int o = blockIdx.x;
    constexpr int pitch = kRun + 4;
    const long long len = X.len();
    const bool line_ok = o < outer;
    const OuterInfo oi = line_ok ? info[o] : OuterInfo{0, -1};
    const int q4 = threadIdx.x % Q4, slot = threadIdx.x / Q4;
    int4 mq = make_int4(1, 1, 1, 1);
    unsigned long long mu[4] = {0ull, 0ull, 0ull, 0ull};
    if (slot < EP) {
        mq = *(const int4 *) (C.moduli + 4 * q4);
#pragma unroll
        for (int e = 0; e < 4; ++e) mu[e] = C.barrett[4 * q4 + e];
    }
    const int mv[4] = {mq.x, mq.y, mq.z, mq.w};
    const int log2M = C.log2M;
    // a block walks the runs blockIdx.y, blockIdx.y + gridDim.y, ... of its line (few, long-lived blocks: cheap to skip when the
    // small-modulus base was chosen)
    for (int run = blockIdx.y; run * kRun < inner_p; run += gridDim.y) {
        const int l0 = run * kRun;
        if (slot < EP) {
            for (int ll = slot; ll < kRun; ll += EP) {
                const int l = l0 + ll;
                unsigned r[4] = {0u, 0u, 0u, 0u};
                int sh16 = kShiftSentinel;
                if (line_ok && l < inner) {
                    const long long idx = (long long) o * so + (long long) l * sl;
                    if (X.eval[idx + len].frac != 0) {
                        long long sh = (long long) X.exp[idx] - oi.emin;
                        const int s = sh > kShiftMax ? kShiftMax : (int) sh;
                        sh16 = s;
                        if (s <= log2M) {   // s > log2M: the line fails the window guard anyway
                            const int4 dg = __ldg((const int4 *) (X.digits + idx * N) + q4);
                            const int4 pw = __ldg((const int4 *) (C.pow2 + (long long) s * N) + q4);
                            const int dv[4] = {dg.x, dg.y, dg.z, dg.w}, pv[4] = {pw.x, pw.y, pw.z, pw.w};
                            const int neg = X.sign[idx];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                int v = mulmod(dv[e], pv[e], mv[e], mu[e]);
                                if (neg && v) v = mv[e] - v;
                                r[e] = (unsigned) v;
                            }
                        }
                    }
                }
                if (q4 == 0) shifts[(long long) o * inner_p + l] = (int16_t) sh16;
#pragma unroll
                for (int b = 0; b < 4; ++b)
#pragma unroll
                    for (int e = 0; e < 4; ++e) sm_stage[(b * N + 4 * q4 + e) * pitch + ll] = (uint8_t) (r[e] >> (8 * b));
            }
        }
        __syncthreads();
        // write out: shared row (b, q) -> plane row (q, b), kRun contiguous bytes, one warp per row
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int r = warp; r < 4 * np; r += 8) {
            const int b = r / np, q = r - b * np;
            const int row = b * N + q;
            const uint32_t v = *(const uint32_t *) (sm_stage + row * pitch + lane * 4);
            *(uint32_t *) (planes + ((long long) (q * 4 + b) * outer_p + o) * inner_p + l0 + lane * 4) = v;
        }
        __syncthreads();
    }
}

// ---- stage 1c: (min,+) product of the shift planes ------------------------------------------------
// delta[j][i] = min_l (SA[i][l] + SB[j][l]); int16 SIMD pairs (DPX add-min on sm_90+).
// CTA tile 128 (i) x 64 (j), 256 threads, 8 x 4 outputs per thread; the shift words of a K slab sit in
// shared memory as [word][row] so that a thread's 8 + 4 operands are three 128-bit loads per 32 add-mins.
constexpr int kMpTI = 128, kMpTJ = 64;
constexpr int kMpK = 64;   // shift entries per stage (32 words)
__global__ void __launch_bounds__(256) k_minplus(const int16_t *SA, const int16_t *SB, int16_t *delta, long long inner_p,
                                                 long long m_p, long long n_p) {
    __shared__ __align__(16) uint32_t sa[kMpK / 2][kMpTI + 4];
    __shared__ __align__(16) uint32_t sb[kMpK / 2][kMpTJ + 4];
    const int i0 = blockIdx.x * kMpTI, j0 = blockIdx.y * kMpTJ;
    // rows of a thread: 4 l .. 4 l + 3 and 64 + 4 l .. 64 + 4 l + 3 (l = thread & 15): the eight lanes of a
    // 128-bit shared-memory phase then cover 32 distinct banks
    const int ti = (threadIdx.x & 15) * 4, tj = (threadIdx.x >> 4) * 4;
    uint32_t acc[8][4];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0x7fff7fffu;
    const uint32_t *A32 = (const uint32_t *) SA, *B32 = (const uint32_t *) SB;
    const long long wpl = inner_p / 2;   // words per line
    for (long long w0 = 0; w0 < wpl; w0 += kMpK / 2) {
        for (int t = threadIdx.x; t < kMpTI * (kMpK / 2); t += 256) {
            const int row = t / (kMpK / 2), w = t - row * (kMpK / 2);
            sa[w][row] = A32[(long long) (i0 + row) * wpl + w0 + w];
        }
        for (int t = threadIdx.x; t < kMpTJ * (kMpK / 2); t += 256) {
            const int row = t / (kMpK / 2), w = t - row * (kMpK / 2);
            sb[w][row] = B32[(long long) (j0 + row) * wpl + w0 + w];
        }
        __syncthreads();
#pragma unroll 4
        for (int w = 0; w < kMpK / 2; ++w) {
            const uint4 a0 = *(const uint4 *) &sa[w][ti], a1 = *(const uint4 *) &sa[w][64 + ti], b0 = *(const uint4 *) &sb[w][tj];
            const uint32_t av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w}, bv[4] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = __viaddmin_s16x2(av[a], bv[b], acc[a][b]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        __align__(16) short out[8];
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            int lo = (int) (short) (acc[a][b] & 0xffffu), hi = (int) (short) (acc[a][b] >> 16);
            out[a] = (short) (lo < hi ? lo : hi);
        }
        const uint2 *o2 = (const uint2 *) out;
        int16_t *drow = delta + (long long) (j0 + tj + b) * m_p + i0;
        *(uint2 *) (drow + ti) = o2[0];
        *(uint2 *) (drow + 64 + ti) = o2[1];
    }
}

// ---- stage 2: per-modulus limb GEMM on the int8 tensor cores ---------------------------------------
// A planes [q][limb][m_p][k_p], B planes [q][limb][n_p][k_p] (u8, K contiguous).  Output plane
// S[q][n_p][m_p] (canonical residues, i contiguous).
// CTA: 128 (i) x 64 (j) outputs, 8 warps as 4 x 2, warp tile 32 x 32 = 2 (m16) x 4 (n8) MMA tiles.
// K is consumed in 64-byte slabs through a 4-deep cp.async pipeline; rows are 64 B with the 16-byte
// chunk index XOR-swizzled by (row >> 1) & 3 so that ldmatrix is bank-conflict free.
// PASS 0 accumulates the limb pairs with s + t <= 3 (4 accumulators per output), PASS 1 those with
// s + t >= 4 (3 accumulators) and adds to the PASS 0 result.
constexpr int kBM = 128, kBN = 64, kBK = 64, kStages = 4;
constexpr int kStageBytesA = 4 * kBM * kBK, kStageBytesB = 4 * kBN * kBK;
constexpr int kStageBytes = kStageBytesA + kStageBytesB;
constexpr int kGemmSmem = kStages * kStageBytes;

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    unsigned sa = (unsigned) __cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void ldmatrix_x4(unsigned (&r)[4], const void *smem) {
    unsigned sa = (unsigned) __cvta_generic_to_shared(smem);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(sa));
}
__device__ __forceinline__ void mma_u8(int (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// byte offset of (row, 16-byte chunk c) inside a [rows][64 B] swizzled tile
__device__ __forceinline__ int swz(int row, int c) { return row * kBK + ((c ^ ((row >> 1) & 3)) << 4); }

template <int PASS>
__global__ void __launch_bounds__(256, 1) k_limb_gemm(const DevConsts *Cp, const uint8_t *PA, const uint8_t *PB, int *S,
                                                      long long m_p, long long n_p, long long k_p, long long k_begin, int k_len, bool add_to_S,
                                                      const int *nprime) {
    extern __shared__ __align__(128) uint8_t smem[];
    if ((int) blockIdx.z >= *nprime) return;   // modulus outside the reduced base
    constexpr int NACC = PASS == 0 ? 4 : 3;
    constexpr int UBASE = PASS == 0 ? 0 : 4;
    constexpr int LIMB0 = PASS == 0 ? 0 : 1;     // PASS 1 never touches limb 0
    constexpr int NL = 4 - LIMB0;
    const int q = blockIdx.z;
    const int i0 = blockIdx.y * kBM, j0 = blockIdx.x * kBN;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wm = (warp & 3) * 32, wn = (warp >> 2) * 32;
    const uint8_t *Aq = PA + (long long) q * 4 * m_p * k_p;
    const uint8_t *Bq = PB + (long long) q * 4 * n_p * k_p;

    int acc[NACC][2][4][4];
#pragma unroll
    for (int u = 0; u < NACC; ++u)
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[u][a][b][c] = 0;
    const int m_q = Cp->moduli[q];
    const unsigned long long mu_q = Cp->barrett[q];
    unsigned long long cu[NACC];   // 2^(8u) mod m
#pragma unroll
    for (int u = 0; u < NACC; ++u) cu[u] = (unsigned long long) (unsigned) Cp->pow2[(long long) (8 * (UBASE + u)) * Cp->N + q];

    const int nslab = k_len / kBK;   // host keeps k_len <= 8064 so that 4 pairs * 255^2 * k_len < 2^31
    // Per-thread copy assignment, fixed across slabs: 16-byte chunk c = tid & 3 of rows r0 and r0 + 64 of
    // every A limb tile and of row r0 of every B limb tile.
    const int cpc = threadIdx.x & 3, cpr = threadIdx.x >> 2;
    const uint8_t *gA = Aq + ((long long) i0 + cpr) * k_p + k_begin + cpc * 16;
    const uint8_t *gB = Bq + ((long long) j0 + cpr) * k_p + k_begin + cpc * 16;
    const long long limbA = m_p * k_p, limbB = n_p * k_p, half = 64 * k_p;
    const int dsw = swz(cpr, cpc);
    auto load_slab = [&](int slab, int stage) {
        uint8_t *sA = smem + stage * kStageBytes + dsw, *sB = smem + stage * kStageBytes + kStageBytesA + dsw;
        const long long kb = (long long) slab * kBK;
#pragma unroll
        for (int lb = LIMB0; lb < 4; ++lb) {
            cp_async16(sA + lb * kBM * kBK, gA + lb * limbA + kb);
            cp_async16(sA + lb * kBM * kBK + 64 * kBK, gA + lb * limbA + half + kb);
            cp_async16(sB + lb * kBN * kBK, gB + lb * limbB + kb);
        }
    };

#pragma unroll
    for (int s = 0; s < kStages - 1; ++s) {
        if (s < nslab) load_slab(s, s);
        cp_async_commit();
    }
    for (int slab = 0; slab < nslab; ++slab) {
        cp_async_wait<kStages - 2>();
        __syncthreads();
        if (slab + kStages - 1 < nslab) load_slab(slab + kStages - 1, (slab + kStages - 1) % kStages);
        cp_async_commit();
        const uint8_t *sA = smem + (slab % kStages) * kStageBytes, *sB = sA + kStageBytesA;
#pragma unroll
        for (int ks = 0; ks < kBK / 32; ++ks) {
            // A fragments: all limbs of the two m16 tiles
            unsigned af[4][2][4];
#pragma unroll
            for (int lb = LIMB0; lb < 4; ++lb)
#pragma unroll
                for (int a = 0; a < 2; ++a) {
                    const int row = wm + a * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                    const int chunk = ks * 2 + (lane >> 4);
                    ldmatrix_x4(af[lb][a], sA + lb * kBM * kBK + swz(row, chunk));
                }
#pragma unroll
            for (int t = LIMB0; t < 4; ++t) {
                // B fragments of limb t: four n8 tiles, (b0, b1) each
                unsigned bf[2][4];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int row = wn + h * 16 + (lane & 7) + (lane >> 4) * 8;
                    const int chunk = ks * 2 + ((lane >> 3) & 1);
                    ldmatrix_x4(bf[h], sB + t * kBN * kBK + swz(row, chunk));
                }
#pragma unroll
                for (int s = LIMB0; s < 4; ++s) {
                    const int u = s + t - UBASE;
                    if (u < 0 || u >= NACC) continue;
#pragma unroll
                    for (int a = 0; a < 2; ++a)
#pragma unroll
                        for (int b = 0; b < 4; ++b) mma_u8(acc[u][a][b], af[s][a], bf[b >> 1][(b & 1) * 2], bf[b >> 1][(b & 1) * 2 + 1]);
                }
            }
        }
    }
    cp_async_wait<0>();
    // C fragment: c0,c1 -> row g, cols 2t,2t+1; c2,c3 -> row g+8
    int *Sq = S + (long long) q * n_p * m_p;
    const int g = lane >> 2, tq = lane & 3;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int i = i0 + wm + a * 16 + g + (c >> 1) * 8;
                const int j = j0 + wn + b * 8 + tq * 2 + (c & 1);
                int *dst = Sq + (long long) j * m_p + i;
                unsigned long long v = (PASS == 1 || add_to_S) ? (unsigned long long) (unsigned) *dst : 0ull;
#pragma unroll
                for (int u = 0; u < NACC; u += 2) {   // two terms < 2^62 each plus v < 2^31 stay below 2^64
                    v += (unsigned long long) (unsigned) acc[u][a][b][c] * cu[u];
                    if (u + 1 < NACC) v += (unsigned long long) (unsigned) acc[u + 1][a][b][c] * cu[u + 1];
                    v = (unsigned long long) (unsigned) reduce64(v, m_q, mu_q);
                }
                *dst = (int) v;
            }
}

}  // namespace mpres

#include "kernels_norm.cuh"
#include "kernels_small.cuh"
#include "kernels_minplus.cuh"

namespace mpres {

// reference-order recomputation of the elements in `todo`, epilogue fused
template <int G, int R>
__global__ void k_gemm_todo(const DevConsts *Cp, bool ta, bool tb, int m, int n, int k, SoA A, int lda, SoA B, int ldb,
                            SoA alpha, SoA beta, SoA Cm, int ldc, const long long *todo, const int *todo_count) {
    const DevConsts &C = *Cp;
    Lane<R> L;
    lane_init<G, R>(C, L);
    long long grp = ((long long) blockIdx.x * blockDim.x + threadIdx.x) / G;
    const long long ngrp = (long long) gridDim.x * blockDim.x / G;
    const long long total = *todo_count;
    if (grp >= total) return;
    Num<R> al, be;
    load_num<G, R>(C, L, alpha, 0, al);
    load_num<G, R>(C, L, beta, 0, be);
    for (; grp < total; grp += ngrp) {
        const long long e = todo[grp];
        const int row = (int) (e % m), col = (int) (e / m);
        Num<R> sum, prod, a, b, c, t1, t2;
        num_zero(sum);
        for (int l = 0; l < k; ++l) {
            load_num<G, R>(C, L, A, mat_index(ta, row, l, lda), a);
            load_num<G, R>(C, L, B, mat_index(tb, l, col, ldb), b);
            mp_mul<G, R, true>(C, L, prod, a, b);
            mp_add<G, R, true>(C, L, sum, sum, prod);
        }
        const long long ic = row + (long long) col * ldc;
        load_num<G, R>(C, L, Cm, ic, c);
        mp_mul<G, R, true>(C, L, t1, sum, al);
        mp_mul<G, R, true>(C, L, t2, c, be);
        mp_add<G, R, true>(C, L, c, t2, t1);
        store_num<G, R>(C, L, Cm, ic, c);
    }
}

}  // namespace mpres

// ---- host launchers ----------------------------------------------------------------------------------

static inline long long round_up(long long v, long long a) { return (v + a - 1) / a * a; }

// S is unused by the fast path (the epilogue is fused); *done tells the caller whether C is final.
inline int gemm_fast_full(mpres_ctx *c, bool ta, bool tb, int m, int n, int k, SoA A, int lda, SoA B, int ldb,
                          SoA alpha, SoA beta, SoA Cm, int ldc, cudaStream_t st, bool *done) {
    *done = false;
    const int N = c->hc.N;
    const long long m_p = round_up(m, kBM), n_p = round_up(n, kBN), k_p = round_up(k, 128);
    if (k_p > 32000 * 128ll) return 0;
    // the small-modulus stage 2 (kernels_small.cuh) is offered whenever its tables exist; k_choose_base decides per call
    const bool small_on = c->stage2 == MPRES_STAGE2_SMALL && c->sc.usable;
    const long long m_ps = round_up(m, kSN), n_ps = round_up(n, 256);   // rows of the one-byte planes: 256 x (128 | 256) tiles of k_small_umma
    // workspace: planes A/B (u8), S (int), shifts, delta, infos, todo
    const size_t bytesPA = (size_t) N * 4 * m_p * k_p, bytesPB = (size_t) N * 4 * n_p * k_p;
    const size_t bytesS = (size_t) N * n_p * m_p * 4;
    const size_t bytesSA = (size_t) m_ps * k_p * 2, bytesSB = (size_t) n_ps * k_p * 2, bytesD = (size_t) n_p * m_p * 2;
    const size_t bytesInfo = (size_t) (m_p + n_p) * sizeof(OuterInfo);
    const size_t bytesTab = (size_t) (3 * c->hc.log2M + 2) * N * sizeof(int);   // alpha * 2^j, beta * 2^j (stage 3)
    const size_t bytesTodo = (size_t) m * n * sizeof(long long);   // per list: reference-order todo, stage-3 slow list
    void *pPA, *pPB, *pS, *pMisc, *pQA = nullptr, *pQB = nullptr, *pS8 = nullptr;
    int rc;
    if ((rc = ws_reserve(c, 3, bytesPA, &pPA))) return rc;
    if ((rc = ws_reserve(c, 4, bytesPB, &pPB))) return rc;
    if ((rc = ws_reserve(c, 5, bytesS, &pS))) return rc;
    if ((rc = ws_reserve(c, 6, bytesSA + bytesSB + bytesD + bytesInfo + 2 * bytesTodo + bytesTab + 1024, &pMisc))) return rc;
    if (small_on) {
        if ((rc = ws_reserve(c, 8, (size_t) kSmallMax * m_ps * k_p, &pQA))) return rc;
        if ((rc = ws_reserve(c, 9, (size_t) kSmallMax * n_ps * k_p, &pQB))) return rc;
        if ((rc = ws_reserve(c, 10, (size_t) kSmallMax * n_ps * m_ps, &pS8))) return rc;
    }
    char *pm = (char *) pMisc;
    int16_t *SA = (int16_t *) pm; pm += bytesSA;
    int16_t *SB = (int16_t *) pm; pm += bytesSB;
    int16_t *D = (int16_t *) pm; pm += (bytesD + 15) / 16 * 16;
    OuterInfo *IA = (OuterInfo *) pm; pm += (size_t) m_p * sizeof(OuterInfo);
    OuterInfo *IB = (OuterInfo *) pm; pm += (size_t) n_p * sizeof(OuterInfo);
    long long *todo = (long long *) pm; pm += bytesTodo;
    long long *slow = (long long *) pm; pm += bytesTodo;
    pm = (char *) (((uintptr_t) pm + 15) & ~(uintptr_t) 15);   // the table rows are read as 128-bit loads
    int *scal_tab = (int *) pm;
    int *nprime = c->d_counter + 2, *sel = c->d_counter + 4;

    // element (o, l) index strides: op(A)(i, l) and op(B)(l, j)
    const long long soA = ta ? lda : 1, slA = ta ? 1 : lda;
    const long long soB = tb ? 1 : ldb, slB = tb ? ldb : 1;
    auto mark = [&](int i) { if (c->profiling) { if (!c->ev[i]) cudaEventCreate(&c->ev[i]); cudaEventRecord(c->ev[i], st); } };
    mark(0);
    int extra_launches = 0;
    k_outer_info<<<(unsigned) ((m * 32ll + 255) / 256), 256, 0, st>>>(c->dconsts, A, soA, slA, m, k, IA);
    k_outer_info<<<(unsigned) ((n * 32ll + 255) / 256), 256, 0, st>>>(c->dconsts, B, soB, slB, n, k, IB);
    k_choose_base<<<1, 256, 0, st>>>(c->dconsts, IA, m, IB, n, k, c->reduced_base, small_on ? 1 : 0, nprime, sel);
    const size_t smem_align = (size_t) N * 4 * (kRun + 4);
    if (!c->attr_fast) {
        cudaFuncSetAttribute(k_align_planes, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 4 * (kRun + 4));
        cudaFuncSetAttribute(k_align_planes4, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 4 * (kRun + 4));
        cudaFuncSetAttribute(k_base_extend, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) base_extend_smem(128));
        cudaFuncSetAttribute(k_limb_gemm<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem);
        cudaFuncSetAttribute(k_limb_gemm<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem);
        cudaFuncSetAttribute(k_ext_small<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) ext_small_smem(512, 128));
        cudaFuncSetAttribute(k_ext_small<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) ext_small_smem(512, 128));
        c->attr_fast = true;
    }
    if (small_on) {
        const unsigned gA = (unsigned) std::min<long long>((m_ps / kASo) * (k_p / kASl), (long long) c->sm_count * MPRES_ALIGN_BLOCKS);
        const unsigned gB = (unsigned) std::min<long long>((n_ps / kASo) * (k_p / kASl), (long long) c->sm_count * MPRES_ALIGN_BLOCKS);
        if (c->align_mma) {
            if (!c->attr_align_mma) { cudaFuncSetAttribute(k_align_small<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) align_small_smem(true)); c->attr_align_mma = true; }
            k_align_small<true><<<gA, 256, align_small_smem(true), st>>>(c->dconsts, A, soA, slA, m, k, IA, (uint8_t *) pQA, SA, m_ps, k_p, sel);
            k_align_small<true><<<gB, 256, align_small_smem(true), st>>>(c->dconsts, B, soB, slB, n, k, IB, (uint8_t *) pQB, SB, n_ps, k_p, sel);
        } else {
            if (!c->attr_align) { cudaFuncSetAttribute(k_align_small<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) align_small_smem(false)); c->attr_align = true; }
            k_align_small<false><<<gA, 256, align_small_smem(false), st>>>(c->dconsts, A, soA, slA, m, k, IA, (uint8_t *) pQA, SA, m_ps, k_p, sel);
            k_align_small<false><<<gB, 256, align_small_smem(false), st>>>(c->dconsts, B, soB, slB, n, k, IB, (uint8_t *) pQB, SB, n_ps, k_p, sel);
        }
        extra_launches += 2;
    }
    if (N % 4 == 0 && N <= 128 && c->stage1 == 0) {
        k_align_planes4<<<dim3((unsigned) m_p, (unsigned) std::min<long long>(k_p / kRun, 4)), 256, smem_align, st>>>(c->dconsts, A, soA, slA, m, k, IA, (uint8_t *) pPA, SA, m_p, k_p, nprime);
        k_align_planes4<<<dim3((unsigned) n_p, (unsigned) std::min<long long>(k_p / kRun, 4)), 256, smem_align, st>>>(c->dconsts, B, soB, slB, n, k, IB, (uint8_t *) pPB, SB, n_p, k_p, nprime);
    } else {
        k_align_planes<<<dim3((unsigned) m_p, (unsigned) (k_p / kRun)), 256, smem_align, st>>>(c->dconsts, A, soA, slA, m, k, IA, (uint8_t *) pPA, SA, m_p, k_p);
        k_align_planes<<<dim3((unsigned) n_p, (unsigned) (k_p / kRun)), 256, smem_align, st>>>(c->dconsts, B, soB, slB, n, k, IB, (uint8_t *) pPB, SB, n_p, k_p);
    }
    if (c->minplus_sparse && k_p >= 512) {
        // (min,+) from candidate lists (kernels_minplus.cuh); SA / SB have m_ps / n_ps allocated rows
        const size_t bT = (size_t) k_p * (m_ps + n_ps) * 2, bD = (size_t) m_p * n_p * 2 * 2, bC = (size_t) (m_p + n_p) * kMcT * 8 + (size_t) (m_p + n_p) * 4;
        void *pMp;
        if ((rc = ws_reserve(c, 11, bT + bD + bC + (size_t) m * n * 8 + 256, &pMp))) return rc;
        char *q = (char *) pMp;
        int16_t *SAT = (int16_t *) q; q += (size_t) k_p * m_ps * 2;
        int16_t *SBT = (int16_t *) q; q += (size_t) k_p * n_ps * 2;
        int16_t *D1 = (int16_t *) q; q += (size_t) m_p * n_p * 2;
        int16_t *D2 = (int16_t *) q; q += (size_t) m_p * n_p * 2;
        int *cposA = (int *) q; q += (size_t) m_p * kMcT * 4;
        int *cvalA = (int *) q; q += (size_t) m_p * kMcT * 4;
        int *cposB = (int *) q; q += (size_t) n_p * kMcT * 4;
        int *cvalB = (int *) q; q += (size_t) n_p * kMcT * 4;
        int *thrA = (int *) q; q += (size_t) m_p * 4;
        int *thrB = (int *) q; q += (size_t) n_p * 4;
        q = (char *) (((uintptr_t) q + 15) & ~(uintptr_t) 15);
        long long *mplist = (long long *) q;
        k_mp_select<<<(unsigned) m, 256, 0, st>>>(SA, k_p, (int) k_p, m, cposA, cvalA, thrA);
        k_mp_select<<<(unsigned) n, 256, 0, st>>>(SB, k_p, (int) k_p, n, cposB, cvalB, thrB);
        k_mp_transpose<<<dim3((unsigned) (m_ps / 64), (unsigned) (k_p / 64)), 256, 0, st>>>(SA, k_p, SAT, m_ps);
        k_mp_transpose<<<dim3((unsigned) (n_ps / 64), (unsigned) (k_p / 64)), 256, 0, st>>>(SB, k_p, SBT, n_ps);
        k_mp_gather<<<(unsigned) m, 256, 0, st>>>(cposA, cvalA, SBT, n_ps, m, (int) n_p, D1, n_p);
        k_mp_gather<<<(unsigned) n, 256, 0, st>>>(cposB, cvalB, SAT, m_ps, n, (int) m_p, D2, m_p);
        k_mp_combine<<<dim3((unsigned) (m_p / 64), (unsigned) (n_p / 64)), 256, 0, st>>>(D1, n_p, D2, m_p, thrA, thrB, m, n, D, m_p, mplist, c->d_counter + 6);
        k_mp_fix<<<c->sm_count * 4, 256, 0, st>>>(SA, SB, k_p, D, m_p, mplist, c->d_counter + 6);
        extra_launches += 7;
    } else {
        k_minplus<<<dim3((unsigned) (m_p / kMpTI), (unsigned) (n_p / kMpTJ)), 256, 0, st>>>(SA, SB, D, k_p, m_p, n_p);
    }
    dim3 grid((unsigned) (n_p / kBN), (unsigned) (m_p / kBM), (unsigned) N);
    mark(1);
    int gemm_launches = 0;
    if (small_on) {
        for (long long kb = 0; kb < k_p; kb += kSmallKChunk) {
            const int kl = (int) std::min<long long>(kSmallKChunk, k_p - kb);
            if ((rc = launch_small_umma(c, (const uint8_t *) pQA, (const uint8_t *) pQB, (uint8_t *) pS8, m_ps, n_ps, k_p, kb, kl, kb > 0, sel, st))) return rc;
            gemm_launches += 1;
        }
    }
    for (long long kb = 0; kb < k_p; kb += 8064) {
        const int kl = (int) std::min<long long>(8064, k_p - kb);
        if (c->stage2 == MPRES_STAGE2_MMA_SYNC) {
            k_limb_gemm<0><<<grid, 256, kGemmSmem, st>>>(c->dconsts, (const uint8_t *) pPA, (const uint8_t *) pPB, (int *) pS, m_p, n_p, k_p, kb, kl, kb > 0, nprime);
            k_limb_gemm<1><<<grid, 256, kGemmSmem, st>>>(c->dconsts, (const uint8_t *) pPA, (const uint8_t *) pPB, (int *) pS, m_p, n_p, k_p, kb, kl, true, nprime);
            gemm_launches += 2;
        } else {
            if ((rc = launch_limb_umma(c, c->stage2 != MPRES_STAGE2_UMMA_UNSTACKED, (const uint8_t *) pPA, (const uint8_t *) pPB, (int *) pS, m_p, n_p, k_p, kb, kl,
                                       kb > 0, st))) return rc;
            gemm_launches += 1;
        }
    }
    mark(2);
    const bool allow_fb = c->mode == MPRES_MODE_AUTO;
    int stage3_launches = 0;
    bool have_fast = c->stage3 == 0;
    switch (N) { case 8: case 16: case 24: case 32: case 40: case 48: case 56: case 64: break; default: have_fast = false; }
    const bool f32 = c->sc.usable && c->sc.red_shift >= 24 && c->sc.red_shift <= 27 && c->norm32;
    const bool fused = small_on && have_fast && c->fuse_ext;
    if (small_on && !fused) {
        const unsigned gx = (unsigned) std::min<long long>((m_p / kXT) * n, (long long) c->sm_count * 4);     // persistent: four blocks per SM
        const size_t sm = ext_small_smem(c->sc.ext_cols, N);
        if (c->sc.red_shift) k_ext_small<true><<<gx, kXT, sm, st>>>(c->dconsts, m, n, (const uint8_t *) pS8, m_p, m_ps, n_ps, (int *) pS, n_p, sel);
        else k_ext_small<false><<<gx, kXT, sm, st>>>(c->dconsts, m, n, (const uint8_t *) pS8, m_p, m_ps, n_ps, (int *) pS, n_p, sel);
        ++extra_launches;
    }
    if (c->reduced_base && N % 4 == 0) {
        const dim3 gx((unsigned) ((m + kExtThreads - 1) / kExtThreads), (unsigned) std::min(n, 256));
        k_base_extend<<<gx, kExtThreads, base_extend_smem(N), st>>>(c->dconsts, m, n, (int *) pS, m_p, n_p, c->d_counter + 2);
        ++stage3_launches;
    }
    auto norm_fast = [&](auto tag) {
        constexpr int NQ = decltype(tag)::value;
        const unsigned g3 = (unsigned) ((long long) ((m + kNormFastThreads - 1) / kNormFastThreads) * n);
        const int rowsT = 3 * c->hc.log2M + 2;
        k_scalar_tables<<<(rowsT * NQ + 255) / 256, 256, 0, st>>>(c->dconsts, alpha, beta, scal_tab);
        const size_t sm_cds = (size_t) kNormFastThreads * (NQ + 1) * sizeof(int);
        if (fused) {
            // small-modulus path: base extension and normalisation in one kernel (leaves at once when the small base was not selected)
            const unsigned gx = (unsigned) ((m_p / kXT) * n);
            const size_t sm = ext_small_smem(c->sc.ext_cols, NQ) + (ext_norm_cds_aliased(NQ) ? 0 : sm_cds);
            if (!(c->attr_fused >> (NQ / 8) & 1ull)) {
                cudaFuncSetAttribute(k_ext_norm_small<NQ, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm);
                cudaFuncSetAttribute(k_ext_norm_small<NQ, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm);
                c->attr_fused |= 1ull << (NQ / 8);
            }
            if (f32)
                k_ext_norm_small<NQ, true><<<gx, kXT, sm, st>>>(c->dconsts, m, n, k, (const uint8_t *) pS8, m_p, m_ps, n_ps, (int *) pS, n_p, sel, D, IA, IB, alpha, beta, Cm, ldc,
                                                               scal_tab, todo, c->d_counter, slow, c->d_counter + 1, allow_fb);
            else
                k_ext_norm_small<NQ, false><<<gx, kXT, sm, st>>>(c->dconsts, m, n, k, (const uint8_t *) pS8, m_p, m_ps, n_ps, (int *) pS, n_p, sel, D, IA, IB, alpha, beta, Cm, ldc,
                                                                scal_tab, todo, c->d_counter, slow, c->d_counter + 1, allow_fb);
            ++extra_launches;
        }
        // limb-plane path (or unfused small path); `gate`: leave at once when the fused kernel did the work
        const int *gate = fused ? sel : nullptr;
        auto launch_norm = [&](auto kern, size_t smem) {
            kern<<<g3, kNormFastThreads, smem, st>>>(c->dconsts, m, n, k, (const int *) pS, D, m_p, n_p, IA, IB, alpha, beta, Cm, ldc,
                                                     scal_tab, todo, c->d_counter, slow, c->d_counter + 1, allow_fb, gate);
        };
        if (c->norm_staged) {
            const size_t sm_st = sm_cds + (size_t) NQ * kNormFastThreads * sizeof(int);
            if (sm_st > 48 * 1024 && !(c->attr_norm >> (NQ / 8) & 1ull)) {
                cudaFuncSetAttribute(k_norm_fast<NQ, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm_st);
                cudaFuncSetAttribute(k_norm_fast<NQ, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm_st);
                c->attr_norm |= 1ull << (NQ / 8);
            }
            if (f32) launch_norm(k_norm_fast<NQ, true, true>, sm_st); else launch_norm(k_norm_fast<NQ, false, true>, sm_st);
        } else {
            if (f32) launch_norm(k_norm_fast<NQ, true, false>, sm_cds); else launch_norm(k_norm_fast<NQ, false, false>, sm_cds);
        }
        ++stage3_launches;
    };
    if (have_fast) {
        switch (N) {
            case 8: norm_fast(std::integral_constant<int, 8>{}); break;
            case 16: norm_fast(std::integral_constant<int, 16>{}); break;
            case 24: norm_fast(std::integral_constant<int, 24>{}); break;
            case 32: norm_fast(std::integral_constant<int, 32>{}); break;
            case 40: norm_fast(std::integral_constant<int, 40>{}); break;
            case 48: norm_fast(std::integral_constant<int, 48>{}); break;
            case 56: norm_fast(std::integral_constant<int, 56>{}); break;
            case 64: norm_fast(std::integral_constant<int, 64>{}); break;
            default: have_fast = false;
        }
    }
    MPRES_DISPATCH(N, {
        if (have_fast) {
            k_norm_list<G, R><<<c->sm_count * 8, 256, 0, st>>>(c->dconsts, m, n, k, (const int *) pS, D, m_p, n_p, IA, IB, alpha, beta, Cm, ldc,
                                                              slow, c->d_counter + 1);
            stage3_launches = 2;
        } else {
            constexpr int kNormTile = 256 / G;
            const unsigned g3 = (unsigned) ((long long) ((m + kNormTile - 1) / kNormTile) * n);
            k_normalize_epilogue<G, R><<<g3, 256, (size_t) N * (kNormTile + 1) * 4, st>>>(
                c->dconsts, m, n, k, (const int *) pS, D, m_p, n_p, IA, IB, alpha, beta, Cm, ldc, todo, c->d_counter, allow_fb);
            stage3_launches = 1;
        }
        if (allow_fb) {
            k_gemm_todo<G, R><<<c->sm_count * 8, 128, 0, st>>>(c->dconsts, ta, tb, m, n, k, A, lda, B, ldb, alpha, beta, Cm, ldc, todo, c->d_counter);
            ++stage3_launches;
        }
    });
    mark(3);
    c->ev_valid = c->profiling;
    c->last_stage2_launches = gemm_launches;
    for (int i = 0; i < 6 + stage3_launches + gemm_launches + extra_launches; ++i) LAUNCHED(c);
    CUDA_TRY(cudaGetLastError());
    *done = true;
    return 0;
}

#include "kernels_vec.cuh"
