// kernels_small.cuh -- stage 2 of the fast mp_gemm path on a base of ONE-BYTE moduli.
//
// The exact sums S(i,j) = sum_l +-X_a X_b 2^(shift) of the fast path are plain integers with |S| < 2^need
// (need from the per-row / per-column magnitude windows).  The reference's 26.6-bit moduli cost sixteen
// limb x limb int8 GEMMs each on the tensor cores (kernels_umma.cuh): 1.7 bits of S per GEMM.  Any pairwise
// coprime moduli determine S, so here the multiply-accumulate runs modulo 256, 251, 243, 241, ... (host_consts.hpp:
// kSmallModuli): ONE u8 x u8 -> s32 GEMM per modulus, ~7.6 bits of S per GEMM, 4.5 times fewer tensor-core
// cycles for the same exact result.  (The reference has no counterpart: its k-loop multiplies residues one
// thread per entry, src/blas/gemm.cuh:39-58.)
//
//   k_align_small    stage 1: an entry's significand X is rebuilt in binary from its first n_in reference
//                    residues (CRT with a floating-point rank that is put right when it is off by one; n_in is
//                    the smallest count whose moduli product exceeds every significand, known from the interval
//                    evaluations), reduced modulo every small modulus with byte dot products (dp4a),
//                    multiplied by +-2^shift and stored as one u8 plane per modulus, K-major.  Persistent.
//   k_small_umma_p   stage 2 (default): persistent, per modulus 256 x 256 tiles of B'_p A'_p^T on tcgen05.mma
//                    kind::i8 (two M128 N256 MMAs per 32-byte K step), operands by TMA in 128-byte swizzled
//                    rows through an mbarrier ring, accumulators in TMEM, reduced mod p in the epilogue.
//   k_small_umma     the first version: one 128 x 256 tile per CTA, 64-byte rows (kept for A/B).
//   k_ext_small      stage 3a: CRT base extension to the reference moduli.  xi_i = x_i (M'/p_i)^-1 mod p_i, rank
//                    R = nearest integer of sum xi_i / p_i (|S| < M'/4), and
//                        S mod m_q = sum_i xi_i (M'/p_i mod m_q) + R (m_q - M' mod m_q)      (mod m_q)
//                    evaluated as an int8 MMA (mma.sync m16n8k32) of the byte vector (xi, R) with the four byte
//                    limbs of the constants.  Output: the same residue planes S[q][j][i] the limb kernels produce,
//                    so the normalisation kernels (kernels_norm.cuh) are shared and the results are identical.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ctx.hpp"
#include "kernels_umma.cuh"

namespace mpres {

// v mod p for v < 2^30, mu = floor(2^32 / p)
__device__ __forceinline__ unsigned small_mod(unsigned v, unsigned p, unsigned mu) {
    const unsigned r = v - __umulhi(v, mu) * p;   // in [0, 2p)
    return r >= p ? r - p : r;
}

__device__ __forceinline__ void mma_u8_frag(int (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// ---- stage 1: alignment into the small base ---------------------------------------------------------------
// planes: [j][outer_p][inner_p] u8 (inner contiguous), shifts: [outer_p][inner_p] int16 as in k_align_planes.
// A block handles kASo lines x kASl inner positions, one entry per thread.
constexpr int kASo = 8, kASl = 32;

// bits [sh, sh + width) of x, in place (width > 0)
template <int N>
__device__ __forceinline__ void slice_extract(unsigned (&x)[N], int sh, int width) {
    const int ws = sh >> 5, bs = sh & 31;
#pragma unroll
    for (int step = 1; step < N; step <<= 1)
        if (ws & step) {
#pragma unroll
            for (int w = 0; w < N; ++w) x[w] = (w + step < N) ? x[w + step] : 0u;
        }
    if (ws >= N) {
#pragma unroll
        for (int w = 0; w < N; ++w) x[w] = 0u;
    }
#pragma unroll
    for (int w = 0; w < N; ++w) {
        unsigned v = __funnelshift_r(x[w], w + 1 < N ? x[w + 1] : 0u, bs);
        const int rem = width - 32 * w;              // bits of this word inside the slice
        v = rem >= 32 ? v : (rem <= 0 ? 0u : (v & ((1u << rem) - 1u)));
        x[w] = v;
    }
}

// 16 bytes of a digit row of which only the leading residues are read: fetch 64 bytes of the line, not 128 (see the cp.async of the first quad)
__device__ __forceinline__ int4 ldg_row16(const int4 *p) {
    int4 v;
    asm volatile("ld.global.nc.L2::64B.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

#ifdef MPRES_ALIGN_UNROLL
constexpr int kAlignUnroll = MPRES_ALIGN_UNROLL;      // A/B knob: unroll factor of the loop over the one-byte moduli (quads)
#endif
// NW: words of the binary significand = reference residues read (>= n_in); CW = NW rounded up to a multiple of four
// is the row pitch of the staged tables.  width > 0: only bits [slice width, (slice + 1) width) of the significand are converted.
template <int NW, bool MMA>
__device__ __forceinline__ void align_small_entry(const int4 (&dg)[(NW + 3) / 4], int nin, int P, int srow, const unsigned *s_mi, const unsigned *s_negmp,
                                                  const int *s_m, const unsigned long long *s_bmu, const int *s_w, const double *s_rcpm,
                                                  const unsigned *s_cw, const unsigned *s_ppm, const uint8_t *pws,
                                                  uint8_t *out, uint8_t *xrow, int slice = 0, int width = 0) {
    constexpr int CW = (NW + 3) & ~3;
    // CRT over the first nin residues: X = sum xi_i M'_i - R M' with R = floor(sum xi_i / m_i).  The double sum can miss R
    // by one when X / M' is within 2^-48 of an integer; then the result is off by exactly M' and is put right below.
    unsigned xi[NW];
    double sum = 0.0;
#pragma unroll
    for (int i = 0; i < NW; ++i) {
        const int4 d4 = dg[i >> 2];
        const int d = (i & 3) == 0 ? d4.x : (i & 3) == 1 ? d4.y : (i & 3) == 2 ? d4.z : d4.w;
        unsigned v = 0;
        if (i < nin) {
            v = (unsigned) mulmod(d, s_w[i], s_m[i], s_bmu[i]);
            sum += (double) v * s_rcpm[i];
        }
        xi[i] = v;
    }
    const unsigned R = (unsigned) __double2int_rd(sum);
    unsigned x[NW];
    {
        unsigned long long carry = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            unsigned long long col = carry + (unsigned long long) R * s_negmp[w];
#pragma unroll
            for (int i = 0; i < NW; ++i) col += (unsigned long long) xi[i] * s_mi[i * CW + w];
            x[w] = w < nin ? (unsigned) col : 0u;
            carry = col >> 32;
        }
        // t = x - M' (= x + negmp mod 2^(32 nin)); no borrow <=> x >= M'
        unsigned t[NW];
        unsigned long long c2 = 0;
        unsigned top = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            if (w < nin) {
                c2 += (unsigned long long) x[w] + s_negmp[w];
                t[w] = (unsigned) c2;
                c2 >>= 32;
                top = x[w];
            } else {
                t[w] = 0;
            }
        }
        if ((int) top < 0) {                        // x "negative" (M' < 2^(32 nin - 1)): the rank was one too large, x += M' (= x - negmp)
            unsigned long long bw = 0;
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                if (w < nin) {
                    const unsigned long long df = (unsigned long long) x[w] - s_negmp[w] - bw;
                    x[w] = (unsigned) df;
                    bw = (df >> 32) & 1ull;
                }
            }
        } else if (c2) {                            // x >= M': the rank was one too small
#pragma unroll
            for (int w = 0; w < NW; ++w) x[w] = t[w];
        }
    }
    if (width > 0) slice_extract<NW>(x, slice * width, width);
    if (MMA) {
        // the binary significand goes to this entry's operand row; the residues follow on the tensor cores (k_align_small<true>)
#pragma unroll
        for (int w4 = 0; w4 < (NW + 7) / 8 * 2; ++w4)
            *(uint4 *) (xrow + 16 * w4) = make_uint4(4 * w4 < NW ? x[4 * w4] : 0u, 4 * w4 + 1 < NW ? x[4 * w4 + 1] : 0u,
                                                     4 * w4 + 2 < NW ? x[4 * w4 + 2] : 0u, 4 * w4 + 3 < NW ? x[4 * w4 + 3] : 0u);
        return;
    }
    // residues modulo the small moduli, times +-2^shift (the entry's 64-byte row of the table: L1-resident, and L1 is what this kernel
    // lives on -- staging the rows in shared memory shrank it and cost 5x)
    const unsigned *mrow = (const unsigned *) (pws + (size_t) srow * 64);
    const int nq4 = width > 0 ? (((width + 31) >> 5) + 3) >> 2 : CW / 4;      // word quads that can be non-zero (a slice is short)
#ifdef MPRES_ALIGN_UNROLL
#pragma unroll kAlignUnroll
#endif
    for (int jg = 0; 4 * jg < P; ++jg) {
        const unsigned mult4 = __ldg(mrow + jg);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int j = 4 * jg + e;
            unsigned v = 0;
#pragma unroll
            for (int w4 = 0; w4 < CW / 4; ++w4) {
                if (w4 >= nq4) break;
                const uint4 c4 = *(const uint4 *) (s_cw + j * CW + 4 * w4);
                v = __dp4a(x[4 * w4], c4.x, v);
                if (4 * w4 + 1 < NW) v = __dp4a(x[4 * w4 + 1], c4.y, v);
                if (4 * w4 + 2 < NW) v = __dp4a(x[4 * w4 + 2], c4.z, v);
                if (4 * w4 + 3 < NW) v = __dp4a(x[4 * w4 + 3], c4.w, v);
            }
            const uint2 pm = *(const uint2 *) (s_ppm + 2 * j);            // (p, floor(2^32 / p)); rows j >= P hold (1, 0) and are never written out
            const unsigned np_ = 0u - pm.x;
            const unsigned t = v * __byte_perm(mult4, 0u, 0x4440u + e);   // times byte e of the +-2^shift row
            const unsigned r = t + __umulhi(t, pm.y) * np_;               // t - floor(t mu / 2^32) p, in [0, 2p)
            out[j * (kASo * kASl)] = (uint8_t) min(r, r + np_);
        }
    }
}

template <int NW, bool MMA>
__device__ __forceinline__ void align_small_dispatch(const int *dig, const int4 &d0, int nin, int P, int srow, const unsigned *s_mi, const unsigned *s_negmp,
                                                     const int *s_m, const unsigned long long *s_bmu, const int *s_w, const double *s_rcpm,
                                                     const unsigned *s_cw, const unsigned *s_ppm, const uint8_t *pws, uint8_t *out, uint8_t *xrow,
                                                     int slice = 0, int width = 0) {
    int4 dg[(NW + 3) / 4];
    dg[0] = d0;                                   // prefetched with the entry's other fields
#pragma unroll
    for (int g = 1; g < (NW + 3) / 4; ++g) dg[g] = (4 * g < nin) ? ldg_row16((const int4 *) dig + g) : make_int4(0, 0, 0, 0);
    align_small_entry<NW, MMA>(dg, nin, P, srow, s_mi, s_negmp, s_m, s_bmu, s_w, s_rcpm, s_cw, s_ppm, pws, out, xrow, slice, width);
}

// MMA = true: the residues modulo the one-byte moduli are computed on the tensor cores -- per warp, the 32 binary significands
// (operand rows of 32 or 64 bytes) times the byte table 256^b mod p_j (mma.sync m16n8k32, u8 x u8 -> s32), each thread then finishing
// the (entry, modulus) pairs of its accumulator fragment: times +-2^shift, one Barrett step, one byte store.  MMA = false: byte dot
// products (dp4a) per entry and modulus -- the default: on B200 the tensor-core variant measured 2.8 ms against 2.3 ms for the two
// operands of config 3 (its byte stores and table gathers cost more than the dot products it saves).  Same planes either way.
__host__ __device__ constexpr size_t align_small_smem_base() {   // output tile, shifts and the tables at their largest (16 words)
    return 64 * kASo * kASl + kASo * kASl * 2 + 64 * 16 * 4 + 16 * 16 * 4 + 16 * 4 + 16 * 8 + 16 * 8 + 16 * 4 * 2 + 64 * 4 * 2 + 64;
}
constexpr int kAXPitch = 96;      // bytes per operand row (conflict-free 64-bit fragment loads)
#ifndef MPRES_ALIGN_BLOCKS
#define MPRES_ALIGN_BLOCKS 4
#endif
// resident blocks per SM the alignment kernel is compiled for (and its persistent grid is sized by).  Measured on B200 (round 2): the plain
// instantiation gains 3-5 % from four blocks of 64 registers (config 3: 2 x 1.04 -> 2 x 0.99 ms); the sliced one loses 30 % (17.9 -> 23.6 ms at
// 4096^3 / 424-bit full-precision inputs: its slice extraction spills), and the tensor-core variant's shared memory allows three at most
__host__ __device__ constexpr int align_small_blocks(bool mma, bool sliced) { return (mma || sliced) ? 3 : MPRES_ALIGN_BLOCKS; }
// SLICED: the instantiation for significands cut into pieces (kept apart: the default one is tuned to its register budget)
template <bool MMA, bool SLICED = false>
__global__ void __launch_bounds__(256, align_small_blocks(MMA, SLICED)) k_align_small(const DevConsts *Cp, SoA X, long long so, long long sl, int outer, int inner,
                                                        const OuterInfo *info, uint8_t *planes, int16_t *shifts,
                                                        long long outer_p, long long inner_p, const int *sel, long long plane_rows = 0) {
    extern __shared__ __align__(16) uint8_t as_smem[];
    const int P = sel[0], nin = sel[1];
    if (P <= 0) return;
    if (plane_rows == 0) plane_rows = outer_p;      // rows between consecutive planes (> outer_p: the lines are a block of a larger plane set)
    const int slices = SLICED ? max(1, sel[kSelSlices]) : 1;   // > 1: planes [slice][P] of the significand's pieces (k_choose_base)
    const int width = (SLICED && slices > 1) ? sel[kSelWidth] : 0;
    const DevConsts &C = *Cp;
    const SmallDev &SD = *C.small;
    const int N = C.N;
    const int NWr = nin <= 8 ? nin : (nin <= 12 ? 12 : 16);   // instantiated word counts
    const int CW = (NWr + 3) & ~3;
    // shared: out [64][256] u8 | shifts [8][32] i16 | cw [64][CW] | mi [CW][CW] | negmp [CW] | bmu [CW] u64 | rcpm [CW] f64 | m, w [CW] | p, pmu [64]
    uint8_t *s_out = as_smem;
    int16_t *s_sh = (int16_t *) (s_out + 64 * kASo * kASl);
    unsigned *s_cw = (unsigned *) (s_sh + kASo * kASl);
    unsigned *s_mi = s_cw + 64 * CW;
    unsigned *s_negmp = s_mi + CW * CW;
    unsigned long long *s_bmu = (unsigned long long *) (s_negmp + CW);
    double *s_rcpm = (double *) (s_bmu + CW);
    int *s_m = (int *) (s_rcpm + CW);
    int *s_w = s_m + CW;
    unsigned *s_ppm = (unsigned *) (s_w + CW);      // [64] x (p, floor(2^32 / p))
    // MMA only (behind the tables of the largest configuration): operand rows of the entries, their table rows, the byte table
    uint8_t *s_x = as_smem + align_small_smem_base();          // [256][kAXPitch]
    int *s_srow = (int *) (s_x + 256 * kAXPitch);                // [256]
    uint8_t *s_cwB = (uint8_t *) (s_srow + 256);                 // [56][kAXPitch]
    // prefetch area (cp.async targets): the fields of the NEXT tile's entries, one slot per thread, and the exponent bases of its lines.
    // Kept to 8.2 KB: three blocks must stay within the 100 KB shared-memory configuration -- the rest of the SM's 256 KB is the L1 this
    // kernel's table rows and register spills live in (with 58 KB per block the L1 hit rate fell from 55 % to 6 % and the kernel took 5x).
    uint8_t *pf = as_smem + align_small_smem_base() + (MMA ? 256 * kAXPitch + 256 * 4 + 56 * kAXPitch : 0);
    int4 *s_fd0 = (int4 *) pf;                                    // [256]
    double *s_fup = (double *) (s_fd0 + 256);                     // [256]
    int *s_fex = (int *) (s_fup + 256), *s_fsg = s_fex + 256, *s_fem = s_fsg + 256;   // [256], [256], [kASo]
    // Persistent: a block walks the tiles blockIdx.x, blockIdx.x + gridDim.x, ... (line-block fastest); the tables are staged once and
    // the entry fields of the NEXT tile (interval bound, exponent, sign, line base, first four residues) are loaded while the current
    // tile is converted, so their DRAM latency is off the critical path.
    // The fields travel by cp.async into this thread's shared-memory slot: held in registers across the conversion they were spilled,
    // and the spill store waited for the load (ncu: 14 % of the kernel's stall samples on one STL).
    const int tiles_o = (int) (outer_p / kASo);
    const long long ntiles = (long long) tiles_o * (inner_p / kASl);
    int ol, ll;
    if (so == 1) { ol = threadIdx.x & (kASo - 1); ll = threadIdx.x / kASo; }     // lines are the contiguous direction of the input
    else { ll = threadIdx.x & (kASl - 1); ol = threadIdx.x / kASl; }
    const int slot = ol * kASl + ll;
    const long long xlen = X.len();
    auto entry_index = [&](long long tile, long long &idx) {
        const int o = (int) (tile % tiles_o) * kASo + ol, l = (int) (tile / tiles_o) * kASl + ll;
        const bool inside = tile < ntiles && o < outer && l < inner;
        idx = inside ? (long long) o * so + (long long) l * sl : 0;
        return inside;
    };
    auto cp_async = [](void *dst, const void *src, auto bytes) {
        const unsigned sa = (unsigned) __cvta_generic_to_shared(dst);
        // (.L2::64B: a 16-byte request into an otherwise untouched 128-byte digit row costs a full line of DRAM time by default and half of it with
        // this prefetch size -- tools/row_read_probe.cu, profiles/r02_row_read_probe.json; cudaLimitMaxL2FetchGranularity changes nothing)
        if (decltype(bytes)::value == 16) asm volatile("cp.async.cg.shared.global.L2::64B [%0], [%1], 16;\n" ::"r"(sa), "l"(src));
        else if (decltype(bytes)::value == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(src));
        else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(sa), "l"(src));
    };
    auto prefetch_fields = [&](long long tile) {
        long long idx;
        if (entry_index(tile, idx)) {
            cp_async(s_fd0 + threadIdx.x, X.digits + idx * N, std::integral_constant<int, 16>{});
            cp_async(s_fup + threadIdx.x, &X.eval[idx + xlen].frac, std::integral_constant<int, 8>{});
            cp_async(s_fex + threadIdx.x, X.exp + idx, std::integral_constant<int, 4>{});
            cp_async(s_fsg + threadIdx.x, X.sign + idx, std::integral_constant<int, 4>{});
        } else {
            s_fup[threadIdx.x] = 0.0;                                // reads as an exact zero
        }
        if (threadIdx.x < kASo) {
            const int o = (int) (tile % tiles_o) * kASo + threadIdx.x;
            if (tile < ntiles && o < outer) cp_async(s_fem + threadIdx.x, &info[o].emin, std::integral_constant<int, 4>{});
        }
    };
    prefetch_fields(blockIdx.x);
    asm volatile("cp.async.wait_all;\n" ::: "memory");              // made visible to the block by the barrier below

    for (int t = threadIdx.x; t < ((P + 3) & ~3) * CW; t += 256) { const int j = t / CW, w = t - j * CW; s_cw[t] = SD.cw[j * 16 + w]; }
    for (int t = threadIdx.x; t < CW * CW; t += 256) { const int i = t / CW, w = t - i * CW; s_mi[t] = (i < nin && w < nin) ? SD.in_mi[((size_t) nin * 16 + i) * 16 + w] : 0u; }
    if (threadIdx.x < CW) {
        const int i = threadIdx.x;
        const bool on = i < nin;
        s_negmp[i] = on ? SD.in_negmp[nin * 16 + i] : 0u;
        s_m[i] = on ? C.moduli[i] : 1;
        s_bmu[i] = on ? C.barrett[i] : 0ull;
        s_rcpm[i] = on ? 1.0 / (double) C.moduli[i] : 0.0;
        s_w[i] = on ? C.ext_w[nin * N + i] : 0;
    }
    if (threadIdx.x >= 64 && threadIdx.x < 128) { const int j = threadIdx.x - 64; s_ppm[2 * j] = (unsigned) SD.p[j]; s_ppm[2 * j + 1] = SD.mu[j]; }
    if (MMA) {
        for (int t = threadIdx.x; t < 56 * 16; t += 256) { const int j = t >> 4, w = t & 15; *(unsigned *) (s_cwB + j * kAXPitch + 4 * w) = SD.cw[j * 16 + w]; }
    }
    __syncthreads();
    const int KS = (NWr + 7) / 8, NT = (P + 7) / 8;     // 32-byte K steps, 8-modulus column tiles

    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int o0 = (int) (tile % tiles_o) * kASo, l0 = (int) (tile / tiles_o) * kASl;
        long long cur_idx;
        const bool cur_inside = entry_index(tile, cur_idx);
        const double cur_upf = s_fup[threadIdx.x];
        const int4 cur_d0 = s_fd0[threadIdx.x];
        int sh16 = kShiftSentinel;
        const bool live = cur_inside && cur_upf != 0;
        int srow = 0;
        if (live) {
            const long long sh = (long long) s_fex[threadIdx.x] - s_fem[ol];
            const int s = sh > kSmallShiftMax ? kSmallShiftMax : (sh < 0 ? 0 : (int) sh);   // the selection guarantees sh <= kSmallShiftMax
            sh16 = s;
            srow = 2 * s + (s_fsg[threadIdx.x] ? 1 : 0);
        }
        __syncthreads();                                                      // every thread has read its fields (and the line bases)
        prefetch_fields(tile + gridDim.x);                                   // in flight through the conversion
        s_sh[slot] = (int16_t) sh16;
        uint8_t *xrow = s_x + threadIdx.x * kAXPitch;
        for (int slice = 0; slice < slices; ++slice) {
        if (slice > 0) __syncthreads();                                       // the previous slice's planes have left the output tile
        if (live) {
            const int *dig = X.digits + cur_idx * N;
            uint8_t *outp = s_out + slot;
#define MPRES_AS_CASE(NW_) case NW_: align_small_dispatch<NW_, MMA>(dig, cur_d0, nin, P, srow, s_mi, s_negmp, s_m, s_bmu, s_w, s_rcpm, s_cw, s_ppm, SD.pws, outp, xrow, SLICED ? slice : 0, SLICED ? width : 0); break;
            switch (NWr) {
                MPRES_AS_CASE(1) MPRES_AS_CASE(2) MPRES_AS_CASE(3) MPRES_AS_CASE(4) MPRES_AS_CASE(5) MPRES_AS_CASE(6) MPRES_AS_CASE(7) MPRES_AS_CASE(8)
                MPRES_AS_CASE(12)
                default: align_small_dispatch<16, MMA>(dig, cur_d0, nin, P, srow, s_mi, s_negmp, s_m, s_bmu, s_w, s_rcpm, s_cw, s_ppm, SD.pws, outp, xrow, SLICED ? slice : 0, SLICED ? width : 0); break;
            }
#undef MPRES_AS_CASE
        } else if (MMA) {
#pragma unroll
            for (int w4 = 0; w4 < 4; ++w4) *(uint4 *) (xrow + 16 * w4) = make_uint4(0u, 0u, 0u, 0u);   // a zero significand gives zero residues
        } else {
            for (int j = 0; j < P; ++j) s_out[j * (kASo * kASl) + slot] = 0;
        }
        if (MMA) {
            s_srow[threadIdx.x] = srow;
            __syncwarp();                                   // a warp consumes the 32 operand rows it wrote
            const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
            int acc[2][7][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 7; ++nt)
#pragma unroll
                    for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0;
            for (int ks = 0; ks < KS; ++ks) {
                unsigned a[2][4];
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    const uint8_t *r0 = s_x + (warp * 32 + mt * 16 + g) * kAXPitch + ks * 32 + 8 * t;
                    const uint2 a02 = *(const uint2 *) r0, a13 = *(const uint2 *) (r0 + 8 * kAXPitch);
                    a[mt][0] = a02.x; a[mt][1] = a13.x; a[mt][2] = a02.y; a[mt][3] = a13.y;
                }
#pragma unroll
                for (int nt = 0; nt < 7; ++nt) {
                    if (nt < NT) {
                        const uint2 b = *(const uint2 *) (s_cwB + (nt * 8 + g) * kAXPitch + ks * 32 + 8 * t);
                        mma_u8_frag(acc[0][nt], a[0], b.x, b.y);
                        mma_u8_frag(acc[1][nt], a[1], b.x, b.y);
                    }
                }
            }
            // the four entries of this thread's fragment rows: table row and position in the output tile
            int e_srow[4], e_slot[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int te = warp * 32 + (q >> 1) * 16 + g + 8 * (q & 1);
                e_srow[q] = s_srow[te];
                e_slot[q] = so == 1 ? ((te & (kASo - 1)) * kASl + te / kASo) : te;
            }
#pragma unroll
            for (int nt = 0; nt < 7; ++nt) {
                if (nt < NT) {
                    const int j = nt * 8 + 2 * t;
                    const uint4 pm = *(const uint4 *) (s_ppm + 2 * j);            // (p_j, mu_j, p_j+1, mu_j+1)
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const unsigned mult2 = __ldg((const unsigned short *) (SD.pws + (size_t) e_srow[q] * 64 + j));
                        const unsigned t0 = (unsigned) acc[q >> 1][nt][2 * (q & 1)] * (mult2 & 0xffu);
                        const unsigned t1 = (unsigned) acc[q >> 1][nt][2 * (q & 1) + 1] * (mult2 >> 8);
                        const unsigned r0 = t0 - __umulhi(t0, pm.y) * pm.x, r1 = t1 - __umulhi(t1, pm.w) * pm.z;
                        s_out[j * (kASo * kASl) + e_slot[q]] = (uint8_t) min(r0, r0 - pm.x);
                        s_out[(j + 1) * (kASo * kASl) + e_slot[q]] = (uint8_t) min(r1, r1 - pm.z);
                    }
                }
            }
        }
        __syncthreads();
        // write out: (j, line) -> 32 contiguous bytes, two 16-byte halves
        for (int v = threadIdx.x; v < P * kASo * 2; v += 256) {
            const int j = v / (kASo * 2), rem = v - j * (kASo * 2);
            const int oo = rem >> 1, h = rem & 1;
            const uint4 val = *(const uint4 *) (s_out + j * (kASo * kASl) + oo * kASl + h * 16);
            *(uint4 *) (planes + ((long long) (slice * P + j) * plane_rows + o0 + oo) * inner_p + l0 + h * 16) = val;
        }
        }   // slices
        if (threadIdx.x < kASo * 4) {
            const int oo = threadIdx.x >> 2, part = threadIdx.x & 3;
            *(uint4 *) (shifts + (long long) (o0 + oo) * inner_p + l0 + part * 8) = *(const uint4 *) (s_sh + oo * kASl + part * 8);
        }
        asm volatile("cp.async.wait_all;\n" ::: "memory");                 // the next tile's fields have landed (issued a whole conversion ago)
        __syncthreads();
    }
}
inline size_t align_small_smem(bool mma) {
    return align_small_smem_base() + (mma ? 256 * kAXPitch + 256 * 4 + 56 * kAXPitch : 0) + 256 * 16 + 256 * 8 + 2 * 256 * 4 + kASo * 4;   // + the prefetch area
}

// ---- stage 2: one u8 GEMM per small modulus on tcgen05 -----------------------------------------------------
constexpr int kSM = 128, kSN = 256, kSK = 64, kSStages = 4;
constexpr int kSABytes = kSM * kSK, kSBBytes = kSN * kSK, kSStageBytes = kSABytes + kSBBytes;
constexpr int kSSmem = kSStages * kSStageBytes + 1024 + 256;
constexpr int kSThreads = 320;        // warps 0-7 epilogue, 8 TMA producer, 9 MMA issuer
constexpr int kSTmemCols = 256;
constexpr int kSmallKChunk = 32768;   // 250^2 * 32768 + 255 < 2^31

// The A operand of the MMA (128 TMEM lanes) is the B' tile (rows j), the B operand (256 columns) the A' tile (rows i): the
// accumulator holds S^T, so a thread of the epilogue owns one j and 128 consecutive i -- whole 16-byte runs of the output plane
// S8[z][j][i] (i contiguous, pitch m_ps).
__global__ void __launch_bounds__(kSThreads, 2)
k_small_umma(const __grid_constant__ CUtensorMap tmJ, const __grid_constant__ CUtensorMap tmI, const DevConsts *Cp, uint8_t *S8,
             long long m_ps, long long n_ps, int k_byte0, int nk, int add_to_S, const int *sel) {
    extern __shared__ uint8_t smem_raw[];
    const int z = blockIdx.z;
    if (z >= sel[0]) return;   // modulus outside the selected base (whole CTA leaves before any setup)
    uint8_t *smem = (uint8_t *) (((uintptr_t) smem_raw + 1023) & ~(uintptr_t) 1023);
    uint64_t *full = (uint64_t *) (smem + kSStages * kSStageBytes);
    uint64_t *empty = full + kSStages;
    uint64_t *accum_bar = empty + kSStages;
    uint32_t *tmem_slot = (uint32_t *) (accum_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i0 = blockIdx.x * kSN, j0 = blockIdx.y * kSM;

    if (warp == 8 && lane == 0) {
        ptx::prefetch_tmap(&tmJ);
        ptx::prefetch_tmap(&tmI);
        for (int s = 0; s < kSStages; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
        ptx::mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 9) ptx::tmem_alloc(tmem_slot, kSTmemCols);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 8) {
        // ===== TMA producer =====
        if (lane == 0) {
            for (int it = 0; it < nk; ++it) {
                const int s = it % kSStages;
                const uint32_t ph = (uint32_t) (it / kSStages) & 1u;
                ptx::mbar_wait(&empty[s], ph ^ 1u);
                ptx::mbar_expect_tx(&full[s], kSStageBytes);
                uint8_t *dst = smem + s * kSStageBytes;
                ptx::tma_load_3d(dst, &tmJ, &full[s], k_byte0 + it * kSK, j0, z);
                ptx::tma_load_3d(dst + kSABytes, &tmI, &full[s], k_byte0 + it * kSK, i0, z);
            }
        }
        __syncwarp();
    } else if (warp == 9) {
        // ===== MMA issuer (one thread): D[128 x 256] += B'[128 x 32] A'[256 x 32]^T per 32-byte K step =====
        if (lane == 0) {
            constexpr uint32_t idesc = ptx::idesc_u8(kSM, kSN);
            for (int it = 0; it < nk; ++it) {
                const int s = it % kSStages;
                const uint32_t ph = (uint32_t) (it / kSStages) & 1u;
                ptx::mbar_wait(&full[s], ph);
                ptx::tc_fence_after();
                const uint32_t a_addr = ptx::smem_u32(smem + s * kSStageBytes), b_addr = a_addr + kSABytes;
#pragma unroll
                for (int ks = 0; ks < kSK / 32; ++ks)
                    ptx::umma_i8(tmem, ptx::smem_desc_sw64(a_addr + ks * 32), ptx::smem_desc_sw64(b_addr + ks * 32), idesc, (it | ks) ? 1u : 0u);
                ptx::umma_commit(&empty[s]);   // frees the stage once these MMAs have read it
            }
            ptx::umma_commit(accum_bar);       // accumulator complete
        }
        __syncwarp();
    } else {
        // ===== epilogue: TMEM -> registers -> mod p -> 16-byte runs of the u8 plane =====
        const SmallDev &SD = *Cp->small;
        const unsigned p = (unsigned) SD.p[z], mu = SD.mu[z];
        ptx::mbar_wait(accum_bar, 0);
        ptx::tc_fence_after();
        const int quad = warp & 3, half = warp >> 2;
        const int j = j0 + quad * 32 + lane;
        uint4 *dst = (uint4 *) (S8 + ((long long) z * n_ps + j) * m_ps + i0 + half * 128);
        const uint32_t tbase = tmem + ((uint32_t) (quad * 32) << 16) + (uint32_t) (half * 128);
#pragma unroll 1
        for (int jc = 0; jc < 4; ++jc) {        // 32 columns = two 16-byte runs per pass
            uint32_t d[4][8];
#pragma unroll
            for (int u = 0; u < 4; ++u) ptx::tmem_ld8(tbase + (uint32_t) (jc * 32 + u * 8), d[u]);
            uint4 prev[2] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
            if (add_to_S) { prev[0] = dst[2 * jc]; prev[1] = dst[2 * jc + 1]; }
            ptx::tmem_ld_wait();
            unsigned wd[8];
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int hw = 0; hw < 2; ++hw) {
                    const uint4 pv = prev[u >> 1];
                    const unsigned pw = ((u & 1) * 2 + hw) == 0 ? pv.x : ((u & 1) * 2 + hw) == 1 ? pv.y : ((u & 1) * 2 + hw) == 2 ? pv.z : pv.w;
                    unsigned o = 0;
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const unsigned v = d[u][hw * 4 + e] + ((pw >> (8 * e)) & 0xffu);   // < 2^31 + 2^8
                        unsigned r = v - __umulhi(v, mu) * p;
                        r = r >= p ? r - p : r;
                        o |= r << (8 * e);
                    }
                    wd[u * 2 + hw] = o;
                }
            dst[2 * jc] = make_uint4(wd[0], wd[1], wd[2], wd[3]);
            dst[2 * jc + 1] = make_uint4(wd[4], wd[5], wd[6], wd[7]);
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 9) ptx::tmem_dealloc(tmem, kSTmemCols);
}

// Persistent variant: one CTA per SM walks the (modulus, tile) list; two TMEM accumulators, so the epilogue of a tile
// (TMEM -> mod p -> global) runs while the MMAs of the next tile are issued, and the TMA ring (8 stages) keeps prefetching
// across tile boundaries.  Tiles are ordered i-fastest within a modulus: the CTAs running at the same time share their
// operand tiles in L2.  The number of moduli is read from `sel`, so no CTA is launched for work that does not exist.
// KB = bytes of K per pipeline stage = width of the TMA box and of the shared-memory swizzle: 128 (SWIZZLE_128B, 4 stages) moves
// the operands in full 128-byte rows -- ncu showed the 64-byte rows of the first version saturating L1TEX (93 %) at 57 % tensor
// activity; 64 (SWIZZLE_64B, 8 stages) is kept for comparison.
constexpr int kPRingBytes = 8 * kSStageBytes;     // 192 KB of operand stages either way
constexpr int kPSmem = kPRingBytes + 1024 + 256;
constexpr int kPTmemCols = 512;

namespace ptx {
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
}  // namespace ptx

namespace ptx {
// K-major operand in the 128B-swizzle canonical layout (rows of 128 bytes, 8-row atoms of 1024 bytes): SBO = 1024 B
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t) ((saddr & 0x3ffffu) >> 4);
    d |= (uint64_t) 1 << 16;
    d |= (uint64_t) (1024 >> 4) << 32;
    d |= (uint64_t) 1 << 46;
    d |= (uint64_t) 2 << 61;                           // layout type SWIZZLE_128B
    return d;
}
}  // namespace ptx

// TJ = rows of B' per tile: 128 (one MMA per K step, two accumulators: the epilogue of a tile overlaps the next tile's MMAs) or 256
// (two MMAs per K step share the 256-row A' operand: a third less operand traffic and twice the work per pipeline stage; both
// accumulators belong to one tile, so its epilogue runs while the TMA ring prefetches the next tile).
// Column panels (sharded calls): the B' planes of panel g are a separate [P][n_ps][k_p] range (4-D tensor map: k, row, plane, panel),
// the sums go to S8 + g * s8_panel.  Panels are walked in ring order from pan.first; before the first load of a panel that another rank
// produces, the TMA thread waits for that panel's arrival flag (written by the producing rank after its copy landed here), so the
// multiplication of the panels already here overlaps the transfer of the others.
struct SmallPanels {
    int count, first, own;       // panels of this launch: first, first + 1, ... (mod ring)
    unsigned epoch;
    const unsigned *flags;       // [ring] arrival epochs (nullptr: every panel is local)
    long long s8_panel;          // bytes between the S8 ranges of consecutive panels
    int ring;                    // panels of the call (0: count)
    long long plane_rows;        // rows of one S8 plane (0: n_ps -- every panel has its own plane set)
    int slices;                  // > 1: operand planes [slice][P]; sums S_d, d = t + u, into planes [d][P] (K runs over the pairs of d)
    int pair;                    // >= 0: this launch multiplies only the pair-th slice pair of every d (the launches add up in S8) -- two planes
                                 // per (d, modulus) in flight instead of 2 npairs: the tiles of a modulus stay in L2
};
__device__ __forceinline__ void wait_arrival(const unsigned *flag, unsigned epoch) {
    const long long t0 = clock64();
    for (;;) {
        unsigned v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if ((int) (v - epoch) >= 0) break;
        if (clock64() - t0 > (1ll << 34)) asm volatile("trap;");      // ~ 8 s: a rank never delivered -- fail instead of hanging
        __nanosleep(100);
    }
    asm volatile("fence.proxy.async;" ::: "memory");                   // the planes are read through the async proxy (TMA)
}
namespace ptx {
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"((uint64_t) map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
}  // namespace ptx

template <int KB, int TJ>
__global__ void __launch_bounds__(kSThreads, 1)
k_small_umma_p(const __grid_constant__ CUtensorMap tmJ, const __grid_constant__ CUtensorMap tmI, const DevConsts *Cp, uint8_t *S8,
               long long m_ps, long long n_ps, int k_byte0, int nk, int add_to_S, const int *sel, const SmallPanels pan) {
    extern __shared__ uint8_t smem_raw[];
    const int P = sel[0];
    if (P <= 0) return;
    const int tiles_i = (int) (m_ps / kSN), tiles_j = (int) (n_ps / TJ);
    const int per_z = tiles_i * tiles_j;
    const int SL = pan.slices > 1 ? pan.slices : 1, ND = 2 * SL - 1;
    const long long per_d = (long long) P * per_z;
    const long long per_panel = per_d * ND;
    const long long total = per_panel * pan.count;
    if ((long long) blockIdx.x >= total) return;
    uint8_t *smem = (uint8_t *) (((uintptr_t) smem_raw + 1023) & ~(uintptr_t) 1023);
    constexpr int kStageA = TJ * KB, kStageB = kSN * KB, kStage = kStageA + kStageB, kPStages = kPRingBytes / kStage;
    constexpr int NH = TJ / kSM;                  // MMAs (128-row halves of the B' tile) per K step
    nk = nk * kSK / KB;                           // the caller counts 64-byte steps
    uint64_t *full = (uint64_t *) (smem + kPRingBytes);
    uint64_t *empty = full + kPStages;
    uint64_t *acc_full = empty + kPStages;      // [2]
    uint64_t *acc_empty = acc_full + 2;         // [2]
    uint32_t *tmem_slot = (uint32_t *) (acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 8 && lane == 0) {
        ptx::prefetch_tmap(&tmJ);
        ptx::prefetch_tmap(&tmI);
        for (int s = 0; s < kPStages; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { ptx::mbar_init(&acc_full[b], 1); ptx::mbar_init(&acc_empty[b], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 9) ptx::tmem_alloc(tmem_slot, kPTmemCols);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 8) {
        // ===== TMA producer =====
        if (lane == 0) {
            long long g = 0;   // stage uses so far
            int arrived = -1;
            for (long long tile = blockIdx.x; tile < total; tile += gridDim.x) {
                const int pi = (int) (tile / per_panel);
                const long long tp = tile - (long long) pi * per_panel;
                const int pg = (pan.first + pi) % (pan.ring ? pan.ring : pan.count);
                const int dd = (int) (tp / per_d);
                const long long tq = tp - (long long) dd * per_d;
                const int z = (int) (tq / per_z), r = (int) (tq - (long long) z * per_z);
                const int j0 = (r / tiles_i) * TJ, i0 = (r % tiles_i) * kSN;
                if (pg != arrived) {
                    if (pan.flags && pg != pan.own) wait_arrival(pan.flags + pg, pan.epoch);
                    arrived = pg;
                }
                int t0 = max(0, dd - (SL - 1)), t1 = min(dd, SL - 1);      // slice pairs (t, dd - t) of this sum
                if (pan.pair >= 0) { t0 += pan.pair; if (t0 > t1) continue; t1 = t0; }
                for (int t = t0; t <= t1; ++t) {
                    const int za = t * P + z, zb = (dd - t) * P + z;
                    for (int it = 0; it < nk; ++it, ++g) {
                        const int s = (int) (g % kPStages);
                        const uint32_t ph = (uint32_t) (g / kPStages) & 1u;
                        ptx::mbar_wait(&empty[s], ph ^ 1u);
                        ptx::mbar_expect_tx(&full[s], kStage);
                        uint8_t *dst = smem + s * kStage;
                        if (pan.plane_rows) ptx::tma_load_4d(dst, &tmJ, &full[s], k_byte0 + it * KB, j0, pg, zb);     // (k, row, panel, plane): one plane set
                        else ptx::tma_load_4d(dst, &tmJ, &full[s], k_byte0 + it * KB, j0, zb, pg);
                        ptx::tma_load_3d(dst + kStageA, &tmI, &full[s], k_byte0 + it * KB, i0, za);
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 9) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            constexpr uint32_t idesc = ptx::idesc_u8(kSM, kSN);
            long long g = 0;
            int lt = 0;
            for (long long tile = blockIdx.x; tile < total; tile += gridDim.x) {
                int nkt = nk;
                if (SL > 1) {
                    const int dd = (int) ((tile % per_panel) / per_d);
                    const int np = min(dd, SL - 1) - max(0, dd - (SL - 1)) + 1;
                    if (pan.pair >= 0) { if (pan.pair >= np) continue; }
                    else nkt = nk * np;
                }
                const int b = NH == 1 ? (lt & 1) : 0, use = NH == 1 ? (lt >> 1) : lt;
                ++lt;
                ptx::mbar_wait(&acc_empty[b], (uint32_t) ((use & 1) ^ 1));   // the epilogue has drained this accumulator
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem + (uint32_t) (b * kSN);
                for (int it = 0; it < nkt; ++it, ++g) {
                    const int s = (int) (g % kPStages);
                    const uint32_t ph = (uint32_t) (g / kPStages) & 1u;
                    ptx::mbar_wait(&full[s], ph);
                    ptx::tc_fence_after();
                    const uint32_t a_addr = ptx::smem_u32(smem + s * kStage), b_addr = a_addr + kStageA;
#pragma unroll
                    for (int ks = 0; ks < KB / 32; ++ks) {
                        const uint64_t bd = KB == 128 ? ptx::smem_desc_sw128(b_addr + ks * 32) : ptx::smem_desc_sw64(b_addr + ks * 32);
#pragma unroll
                        for (int h = 0; h < NH; ++h) {
                            const uint32_t ah = a_addr + h * (kSM * KB) + ks * 32;
                            const uint64_t ad = KB == 128 ? ptx::smem_desc_sw128(ah) : ptx::smem_desc_sw64(ah);
                            ptx::umma_i8(d_tmem + (uint32_t) (h * kSN), ad, bd, idesc, (it | ks) ? 1u : 0u);
                        }
                    }
                    ptx::umma_commit(&empty[s]);
                }
                ptx::umma_commit(&acc_full[b]);
            }
        }
        __syncwarp();
    } else {
        // ===== epilogue warps =====
        const SmallDev &SD = *Cp->small;
        const int quad = warp & 3, half = warp >> 2;
        int lt = 0;
        for (long long tile = blockIdx.x; tile < total; tile += gridDim.x) {
            const int pi = (int) (tile / per_panel);
            const long long tp = tile - (long long) pi * per_panel;
            const int pg = (pan.first + pi) % (pan.ring ? pan.ring : pan.count);
            const int dd = (int) (tp / per_d);
            if (SL > 1 && pan.pair >= 0 && pan.pair >= min(dd, SL - 1) - max(0, dd - (SL - 1)) + 1) continue;
            const long long tq = tp - (long long) dd * per_d;
            const int zm = (int) (tq / per_z), r = (int) (tq - (long long) zm * per_z);
            const int z = dd * P + zm;                                   // output plane
            const int j0 = (r / tiles_i) * TJ, i0 = (r % tiles_i) * kSN;
            const unsigned p = (unsigned) SD.p[zm], mu = SD.mu[zm];
            uint8_t *S8p = S8 + (long long) pg * pan.s8_panel;
            const long long s8_rows = pan.plane_rows ? pan.plane_rows : n_ps;
            const int b = NH == 1 ? (lt & 1) : 0, use = NH == 1 ? (lt >> 1) : lt;
            ++lt;
            ptx::mbar_wait(&acc_full[b], (uint32_t) (use & 1));
            ptx::tc_fence_after();
#pragma unroll 1
            for (int h = 0; h < NH; ++h) {
                const int j = j0 + h * kSM + quad * 32 + lane;
                uint4 *dst = (uint4 *) (S8p + ((long long) z * s8_rows + j) * m_ps + i0 + half * 128);
                const uint32_t tbase = tmem + ((uint32_t) (quad * 32) << 16) + (uint32_t) ((NH == 1 ? b : h) * kSN + half * 128);
                // all 128 columns of this thread into registers first, so the accumulator can be handed back early
                uint32_t d[16][8];
#pragma unroll
                for (int u = 0; u < 16; ++u) ptx::tmem_ld8(tbase + (uint32_t) (u * 8), d[u]);
                ptx::tmem_ld_wait();
                if (h == NH - 1) {
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&acc_empty[b]);
                }
#pragma unroll
                for (int c16 = 0; c16 < 8; ++c16) {     // 16 columns = one 16-byte run per pass
                    uint4 pv = make_uint4(0, 0, 0, 0);
                    if (add_to_S) pv = dst[c16];
                    unsigned wd[4];
#pragma unroll
                    for (int w = 0; w < 4; ++w) {
                        const unsigned pw = w == 0 ? pv.x : w == 1 ? pv.y : w == 2 ? pv.z : pv.w;
                        unsigned o = 0;
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const unsigned v = d[c16 * 2 + (w >> 1)][(w & 1) * 4 + e] + ((pw >> (8 * e)) & 0xffu);
                            unsigned rr = v - __umulhi(v, mu) * p;
                            rr = rr >= p ? rr - p : rr;
                            o |= rr << (8 * e);
                        }
                        wd[w] = o;
                    }
                    dst[c16] = make_uint4(wd[0], wd[1], wd[2], wd[3]);
                }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 9) ptx::tmem_dealloc(tmem, kPTmemCols);
}

// ---- stage 3a: CRT base extension from the small base to the reference moduli ------------------------------
constexpr int kXT = 128;        // entries per block (one per thread in the first phase)
constexpr int kXPitch = 96;     // bytes per operand row in shared memory: 64-bit fragment loads are bank-conflict free
constexpr int kXSPitch = 136;   // ints per residue row of the result tile: fragment-order writes and per-entry reads are conflict free


// shared memory of one extension block: operand rows of the entries | operand rows of the constants | staged one-byte
// residues [56][128] | per-modulus constants | result tile [N][kXSPitch]
struct ExtSmem {
    uint8_t *At, *Bt, *X8;
    uint4 *s_c4;       // [64] per small modulus: (p, floor(2^32 / p), (M'/p)^-1 mod p, bits of 1 / p)
    int *s_S;
};
__host__ __device__ inline size_t ext_small_smem(int ext_cols, int N) {
    return (size_t) kXT * kXPitch + (size_t) ext_cols * kXPitch + 56 * kXT + 64 * 16 + (size_t) N * kXSPitch * 4;
}
__device__ __forceinline__ ExtSmem ext_carve(uint8_t *base, int ext_cols) {
    ExtSmem e;
    e.At = base;
    e.X8 = e.At + kXT * kXPitch;              // At and X8 are dead once the block function returns: the fused kernel reuses them
    e.Bt = e.X8 + 56 * kXT;
    e.s_c4 = (uint4 *) (e.Bt + (size_t) ext_cols * kXPitch);
    e.s_S = (int *) (e.s_c4 + 64);
    return e;
}

__host__ __device__ constexpr bool ext_norm_cds_aliased(int NQ) { return (size_t) kXT * (NQ + 1) * 4 <= (size_t) kXT * kXPitch + 56 * kXT; }

// One block: the kXT consecutive rows from row0 of column col.  Leaves S mod m_q of entry e at s_S[q * kXSPitch + e].
// Ends with a __syncthreads().
// `first`: the block's first tile -- its one-byte residues are fetched here and the constant tables staged; later tiles of a
// persistent block find both in place.  `next_off` >= 0: offset (in a plane of S8) of the tile this block handles next; its residues
// are fetched as soon as the first phase has consumed the current ones, so the copy runs under the tensor-core phase.
template <bool FASTRED>
__device__ __forceinline__ void ext_small_block(const DevConsts &C, const SmallDev &SD, const ExtSmem &E, int P, const uint8_t *S8, long long m_ps,
                                                long long n_ps, int col, int row0, bool first = true, long long next_off = -1) {
    const int N = C.N, cols = SD.ext_cols;
    // a tile's one-byte residues: P runs of 128 contiguous bytes, asynchronous 16-byte copies (all in flight at once)
    auto fetch = [&](const uint8_t *src) {
        const long long plane = n_ps * m_ps;
        for (int v = threadIdx.x; v < P * (kXT / 16); v += kXT) {
            const int j = v >> 3, part = v & 7;
            cp_async16(E.X8 + j * kXT + part * 16, src + (long long) j * plane + part * 16);
        }
        cp_async_commit();
    };
    if (first) {
        fetch(S8 + (long long) col * m_ps + row0);
        const uint4 *bsrc = (const uint4 *) (SD.ext_b + (size_t) P * cols * 64);
        for (int v = threadIdx.x; v < cols * 4; v += kXT) *(uint4 *) (E.Bt + (v >> 2) * kXPitch + (v & 3) * 16) = __ldg(bsrc + v);
        if (threadIdx.x < 64) {
            const int j = threadIdx.x;
            E.s_c4[j] = make_uint4((unsigned) SD.p[j], SD.mu[j], (unsigned) SD.inv[P * 64 + j], __float_as_uint(SD.rcp[j]));
        }
    }
    cp_async_wait<0>();
    __syncthreads();
    // ---- phase 1: xi_i and the rank, one entry per thread ----
    {
        const uint8_t *xsrc = E.X8 + threadIdx.x;
        unsigned wds[16];
        float sum = 0.f;
#pragma unroll
        for (int g = 0; g < 16; ++g) {
            unsigned wd = 0;
            if (4 * g < P) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int j = 4 * g + e;
                    const unsigned xv = j < P ? (unsigned) xsrc[j * kXT] : 0u;
                    const uint4 cj = E.s_c4[j];                                   // (p, floor(2^32 / p), (M'/p)^-1 mod p, 1 / p)
                    const unsigned tj = xv * cj.z;
                    const unsigned rj = tj - __umulhi(tj, cj.y) * cj.x;
                    const unsigned xi = min(rj, rj - cj.x);
                    sum = fmaf((float) xi, __uint_as_float(cj.w), sum);
                    wd |= xi << (8 * e);
                }
            }
            wds[g] = wd;
        }
        const unsigned R = (unsigned) __float2int_rn(sum);
        wds[15] |= R << 24;
        uint4 *dst = (uint4 *) (E.At + threadIdx.x * kXPitch);
#pragma unroll
        for (int g = 0; g < 4; ++g) dst[g] = make_uint4(wds[4 * g], wds[4 * g + 1], wds[4 * g + 2], wds[4 * g + 3]);
    }
    if (next_off >= 0) {
        __syncthreads();                    // every thread has read its residues: the staging area is free for the next tile
        fetch(S8 + next_off);
    } else {
        __syncwarp();                       // a warp only reads the 32 operand rows it wrote
    }
    // ---- phase 2: (xi, R) x limbs on the tensor cores, limbs recombined and reduced per reference modulus ----
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int kred = SD.red_shift;
#pragma unroll 1
    for (int hp = 0; hp * 64 < cols; ++hp) {     // 64 operand columns = 16 reference moduli per pass
        int mq[4];
        unsigned long long muq[4];
        unsigned rmu[4];
#pragma unroll
        for (int pp = 0; pp < 4; ++pp) {
            const int q = hp * 16 + pp * 4 + t;
            const bool on = q < N;
            mq[pp] = on ? C.moduli[q] : 1;
            muq[pp] = (!FASTRED && on) ? C.barrett[q] : 0ull;
            rmu[pp] = (FASTRED && on) ? SD.red_mu[q] : 0u;
        }
#pragma unroll 1
        for (int mt = 0; mt < 2; ++mt) {
            int acc[8][4];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[nt][e] = 0;
            const uint8_t *ar0 = E.At + (warp * 32 + mt * 16 + g) * kXPitch + 8 * t, *ar1 = ar0 + 8 * kXPitch;
            const uint8_t *br = E.Bt + (size_t) (hp * 64 + g) * kXPitch + 8 * t;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                const uint2 a02 = *(const uint2 *) (ar0 + ks * 32), a13 = *(const uint2 *) (ar1 + ks * 32);
                const unsigned a[4] = {a02.x, a13.x, a02.y, a13.y};
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
                    const uint2 b = *(const uint2 *) (br + (size_t) nt * 8 * kXPitch + ks * 32);
                    mma_u8_frag(acc[nt], a, b.x, b.y);
                }
            }
#pragma unroll
            for (int pp = 0; pp < 4; ++pp) {
                const int q = hp * 16 + pp * 4 + t;
#pragma unroll
                for (int rh = 0; rh < 2; ++rh) {
                    const unsigned long long v = (unsigned long long) (unsigned) acc[2 * pp][2 * rh] + ((unsigned long long) (unsigned) acc[2 * pp][2 * rh + 1] << 8) +
                                                 ((unsigned long long) (unsigned) acc[2 * pp + 1][2 * rh] << 16) +
                                                 ((unsigned long long) (unsigned) acc[2 * pp + 1][2 * rh + 1] << 24);   // < 2^47
                    unsigned r;
                    if (FASTRED) {
                        r = barrett_k(v, (unsigned) mq[pp], rmu[pp], kred);
                    } else {
                        r = (unsigned) reduce64(v, mq[pp], muq[pp]);
                    }
                    if (q < N) E.s_S[q * kXSPitch + warp * 32 + mt * 16 + g + 8 * rh] = (int) r;
                }
            }
        }
    }
    __syncthreads();
}

// stand-alone extension: writes the residue planes S[q][col][row] (what the limb kernels produce)
template <bool FASTRED>
__global__ void __launch_bounds__(kXT, 4) k_ext_small(const DevConsts *Cp, int m, int n, const uint8_t *S8, long long m_p, long long m_ps, long long n_ps,
                                                      int *S, long long n_p, const int *sel) {
    extern __shared__ __align__(16) uint8_t xs_smem[];
    const int P = sel[0];
    if (P <= 0) return;
    const DevConsts &C = *Cp;
    const SmallDev &SD = *C.small;
    const ExtSmem E = ext_carve(xs_smem, SD.ext_cols);
    const int tiles = (int) (m_p / kXT);
    const long long total = (long long) tiles * n;
    const int N = C.N;
    // persistent: the constant tables (8 KB at 32 moduli) are staged once per block instead of once per tile
    bool first = true;
    for (long long tl = blockIdx.x; tl < total; tl += gridDim.x) {
        const int col = (int) (tl / tiles);
        const int row0 = (int) (tl - (long long) col * tiles) * kXT;
        const long long nx = tl + gridDim.x;
        const long long next_off = nx < total ? (nx / tiles) * m_ps + (nx % tiles) * kXT : -1;
        ext_small_block<FASTRED>(C, SD, E, P, S8, m_ps, n_ps, col, row0, first, next_off);
        first = false;
        for (int v = threadIdx.x; v < N * (kXT / 4); v += kXT) {
            const int q = v >> 5, part = v & 31;
            *(int4 *) (S + ((long long) q * n_p + col) * m_p + row0 + part * 4) = *(const int4 *) (E.s_S + q * kXSPitch + part * 4);
        }
    }
}

// fused: base extension + entry-per-thread normalisation and alpha/beta epilogue (kernels_norm.cuh: norm_fast_body).  The residue
// planes are only written for the entries handed to the list kernel.
template <int NQ, bool F32>
__global__ void __launch_bounds__(kXT, 4) k_ext_norm_small(const DevConsts *Cp, int m, int n, int k, const uint8_t *S8, long long m_p, long long m_ps, long long n_ps,
                                                           int *S, long long n_p, const int *sel, const int16_t *delta, const OuterInfo *ia, const OuterInfo *ib,
                                                           SoA alpha, SoA beta, SoA Cm, int ldc, const int *scal_tab, long long *todo, int *todo_count,
                                                           long long *slow, int *slow_count, bool fallback_allowed) {
    extern __shared__ __align__(16) uint8_t xs_smem[];
    const int P = sel[0];
    if (P <= 0) return;
    const DevConsts &C = *Cp;
    const SmallDev &SD = *C.small;
    const ExtSmem E = ext_carve(xs_smem, SD.ext_cols);
    // the staged digits of C: in the operand rows of the extension when they fit, else behind everything
    int *cds = ext_norm_cds_aliased(NQ) ? (int *) E.At : (int *) (xs_smem + ext_small_smem(SD.ext_cols, NQ));
    const int tiles = (int) (m_p / kXT);
    const int col = blockIdx.x / tiles;
    const int row0 = (blockIdx.x - col * tiles) * kXT;
    ext_small_block<F32>(C, SD, E, P, S8, m_ps, n_ps, col, row0);
    norm_fast_body<NQ, F32, true>(C, cds, m, n, k, col, row0, E.s_S + threadIdx.x, kXSPitch, S + (long long) col * m_p + row0 + threadIdx.x, n_p * m_p,
                                  delta, m_p, ia, ib, alpha, beta, Cm, ldc, scal_tab, todo, todo_count, slow, slow_count, fallback_allowed);
}

}  // namespace mpres

// map (k, row, plane[, panel]) over u8 planes [panels][planes][rows_p][k_p] (panel_stride bytes between panels), box box_k B x box_rows x 1 x 1
// plane_rows > 0: the planes are plane_rows rows apart (the panels are row blocks of ONE plane set: panel_stride = rows_p * k_p)
inline int small_make_map(CUtensorMap *map, const void *planes, int nplanes, long long rows_p, long long k_p, int box_rows, int box_k = mpres::kSK,
                          int npanels = 0, long long panel_stride = 0, long long plane_rows = 0) {
    mpres_encode_tiled_fn enc = umma_encode_fn();
    if (!enc) return -30;
    cuuint64_t dims[4] = {(cuuint64_t) k_p, (cuuint64_t) rows_p, (cuuint64_t) nplanes, (cuuint64_t) (npanels > 0 ? npanels : 1)};
    cuuint64_t strides[3] = {(cuuint64_t) k_p, (cuuint64_t) (rows_p * k_p), (cuuint64_t) (npanels > 1 ? panel_stride : rows_p * k_p * nplanes)};
    if (plane_rows > 0) {     // (k, row, panel, plane)
        dims[2] = (cuuint64_t) (npanels > 0 ? npanels : 1); dims[3] = (cuuint64_t) nplanes;
        strides[1] = (cuuint64_t) (rows_p * k_p); strides[2] = (cuuint64_t) (plane_rows * k_p);
    }
    cuuint32_t box[4] = {(cuuint32_t) box_k, (cuuint32_t) box_rows, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, npanels > 0 ? 4 : 3, const_cast<void *>(planes), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     box_k == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -31;
}

// One launch = the P one-byte moduli of the call, all tiles of all column panels, K range [k_begin, k_begin + k_len).
// PA: planes of A' [P][m_ps][k_p] (m_ps % 256 == 0); PB: per panel, planes of B' [P][n_ps][k_p] (n_ps % 256 == 0), consecutive panels
// pb_panel bytes apart; S8: per panel [P][n_ps][m_ps].  P is the host's copy of sel[0].
inline int launch_small_umma(mpres_ctx *c, int P, const uint8_t *PA, const uint8_t *PB, long long pb_panel, uint8_t *S8, long long m_ps, long long n_ps, long long k_p,
                             long long k_begin, int k_len, bool add_to_S, const int *sel, const mpres::SmallPanels &pan, cudaStream_t st, long long pb_plane_rows = 0) {
    const int SLc = pan.slices > 1 ? pan.slices : 1;
    CUtensorMap tmJ, tmI;
    int rc;
    const int box_k = (c->small_persistent && c->small_kb == 128 && k_len % 128 == 0 && k_begin % 128 == 0) ? 128 : mpres::kSK;
    const int tj = (c->small_persistent && box_k == 128 && c->small_tj == 256 && n_ps % 256 == 0) ? 256 : mpres::kSM;
    if (!c->attr_small) {
        CUDA_TRY(cudaFuncSetAttribute(mpres::k_small_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, mpres::kSSmem));
        CUDA_TRY(cudaFuncSetAttribute(mpres::k_small_umma_p<64, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, mpres::kPSmem));
        CUDA_TRY(cudaFuncSetAttribute(mpres::k_small_umma_p<128, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, mpres::kPSmem));
        CUDA_TRY(cudaFuncSetAttribute(mpres::k_small_umma_p<128, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, mpres::kPSmem));
        c->attr_small = true;
    }
    if ((rc = small_make_map(&tmI, PA, P * SLc, m_ps, k_p, mpres::kSN, box_k))) return rc;
    if (c->small_persistent) {
        if ((rc = small_make_map(&tmJ, PB, P * SLc, n_ps, k_p, tj, box_k, pan.ring ? pan.ring : pan.count, pb_panel, pb_plane_rows))) return rc;
        const long long max_tiles = (long long) P * (2 * SLc - 1) * (m_ps / mpres::kSN) * (n_ps / tj) * pan.count;
        const unsigned gx = (unsigned) std::min<long long>(max_tiles, (long long) c->sm_count);
        if (tj == 256)
            mpres::k_small_umma_p<128, 256><<<gx, mpres::kSThreads, mpres::kPSmem, st>>>(tmJ, tmI, c->dconsts, S8, m_ps, n_ps, (int) k_begin, k_len / mpres::kSK, add_to_S ? 1 : 0, sel, pan);
        else if (box_k == 128)
            mpres::k_small_umma_p<128, 128><<<gx, mpres::kSThreads, mpres::kPSmem, st>>>(tmJ, tmI, c->dconsts, S8, m_ps, n_ps, (int) k_begin, k_len / mpres::kSK, add_to_S ? 1 : 0, sel, pan);
        else
            mpres::k_small_umma_p<64, 128><<<gx, mpres::kSThreads, mpres::kPSmem, st>>>(tmJ, tmI, c->dconsts, S8, m_ps, n_ps, (int) k_begin, k_len / mpres::kSK, add_to_S ? 1 : 0, sel, pan);
        return 0;
    }
    // one tile per CTA (A/B measurement): local panels only, one launch per panel
    for (int g = 0; g < pan.count; ++g) {
        if ((rc = small_make_map(&tmJ, PB + (long long) g * pb_panel, P, n_ps, k_p, tj, box_k))) return rc;
        dim3 grid((unsigned) (m_ps / mpres::kSN), (unsigned) (n_ps / mpres::kSM), (unsigned) P);
        mpres::k_small_umma<<<grid, mpres::kSThreads, mpres::kSSmem, st>>>(tmJ, tmI, c->dconsts, S8 + (long long) g * pan.s8_panel, m_ps, n_ps, (int) k_begin, k_len / mpres::kSK, add_to_S ? 1 : 0, sel);
    }
    return 0;
}
