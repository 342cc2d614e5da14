// kernels_umma.cuh -- stage 2 of the fast mp_gemm path on the 5th-generation tensor cores:
// tcgen05.mma kind::i8 (u8 x u8 -> s32 in TMEM), operands staged by TMA (cp.async.bulk.tensor) into
// 64B-swizzled shared memory through a 4-stage mbarrier pipeline, warp-specialised (1 TMA warp, 1 MMA
// warp, 8 epilogue warps).  Replaces the reference's per-thread serial residue loop
// (src/blas/gemm.cuh:39-58 -> src/modular.cuh:250-256) for the exact-window regime; the legacy
// mma.sync kernel (k_limb_gemm) stays selectable for A/B comparison.
//
// Work per CTA: one modulus q, a 128 (i) x 64 (j) tile of S_q = sum_l A'_q(i,l) B'_q(l,j) mod m_q.
// A'_q and B'_q are stored as four u8 limb planes each (stage 1), so the product is 16 limb x limb
// integer GEMMs whose results are combined by anti-diagonal u = s + t with weight 2^(8u) mod m_q.
//
// "Stacked-B" formulation.  The four B limb tiles (64 columns each) sit back to back in shared memory
// and form ONE K-major operand with N = 256 "columns" c = 64 t + j.  The MMA of A limb s is issued
// with its accumulator base at TMEM column 64 s, so its output column 64 s + c = 64 (s + t) + j lands
// exactly in the accumulator of anti-diagonal u = s + t.  Four N = 256 MMAs per 32-byte K step do the
// work of sixteen N = 64 MMAs with a third of the shared-memory operand traffic (A 4 KB + B 8 KB per
// 128 tensor cycles instead of A 4 KB + B 2 KB per 32), and the seven accumulators occupy TMEM columns
// [0, 448).  TMEM is zeroed first (tcgen05.st) so every MMA accumulates.
// Exactness: every accumulator receives at most 4 limb pairs, 4 * 255^2 * k_len < 2^31 for k_len <= 8064.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ctx.hpp"

namespace mpres {

constexpr int kUM = 128;                       // tile rows (TMEM lanes)
constexpr int kUN = 64;                        // tile columns
constexpr int kUK = 64;                        // K bytes per pipeline stage (one 64B swizzle row)
constexpr int kUStages = 4;
constexpr int kUABytes = 4 * kUM * kUK;        // 32 KB: [limb][row][k]
constexpr int kUBBytes = 4 * kUN * kUK;        // 16 KB: [limb][col][k]  == one 256-row K-major tile
constexpr int kUStageBytes = kUABytes + kUBBytes;
constexpr int kUSmem = kUStages * kUStageBytes + 1024 /* alignment slack */ + 256 /* barriers */;
constexpr int kUThreads = 320;                 // warps 0-7 epilogue, 8 TMA producer, 9 MMA issuer
constexpr int kUTmemCols = 512;
constexpr int kUTilesPerCta = 8;

namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Bounded spin: a protocol error traps (launch failure) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    for (uint32_t spin = 0;; ++spin) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
        if (ok) return;
        if (spin > (1u << 26)) asm volatile("trap;");
    }
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"((uint64_t) map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t) map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t *slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], u8 x u8 -> s32
__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// mbarrier arrive once all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st16_zero(uint32_t taddr) {
    const uint32_t z = 0;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(z) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor, K-major operand in the 64B-swizzle canonical layout
// (rows of 64 bytes, 8-row atoms of 512 bytes stacked contiguously): SBO = 512 B, LBO unused.
// Field layout as in the CUTLASS UMMA::SmemDescriptor (cute/arch/mma_sm100_desc.hpp).
__device__ __forceinline__ uint64_t smem_desc_sw64(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t) ((saddr & 0x3ffffu) >> 4);        // start address, bits [0,14)
    d |= (uint64_t) 1 << 16;                           // leading byte offset (ignored for swizzled K-major)
    d |= (uint64_t) (512 >> 4) << 32;                  // stride byte offset, bits [32,46)
    d |= (uint64_t) 1 << 46;                           // descriptor version (Blackwell)
    d |= (uint64_t) 4 << 61;                           // layout type SWIZZLE_64B
    return d;
}
// Instruction descriptor (UMMA::InstrDescriptor): dense, no saturate, D = S32, A = B = unsigned 8 bit,
// both K-major, N at bits [17,23) >> 3, M at bits [24,29) >> 4.
__host__ __device__ constexpr uint32_t idesc_u8(int M, int N) {
    return (2u << 4) | (0u << 7) | (0u << 10) | ((uint32_t) (N >> 3) << 17) | ((uint32_t) (M >> 4) << 24);
}

}  // namespace ptx

// STACKED = true : four N = 256 MMAs per K step (see header);  false: sixteen N = 64 MMAs (one per limb pair).
template <bool STACKED>
__global__ void __launch_bounds__(kUThreads, 1)
k_limb_umma(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const DevConsts *Cp, int *S,
            long long m_p, long long n_p, int k_byte0, int nk, int add_to_S, const int *nprime) {
    extern __shared__ uint8_t smem_raw[];
    if ((int) blockIdx.z >= *nprime) return;   // modulus outside the reduced base (whole CTA leaves before any setup)
    uint8_t *smem = (uint8_t *) (((uintptr_t) smem_raw + 1023) & ~(uintptr_t) 1023);
    uint64_t *full = (uint64_t *) (smem + kUStages * kUStageBytes);
    uint64_t *empty = full + kUStages;
    uint64_t *accum_bar = empty + kUStages;
    uint32_t *tmem_slot = (uint32_t *) (accum_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = blockIdx.z;
    const int i0 = blockIdx.y * kUM;
    // A CTA walks kUTilesPerCta consecutive column tiles (one after the other, no overlap between them): eight times fewer CTAs,
    // which is what the launch costs when every CTA leaves at once because the small-modulus base was chosen.
    const int jt0 = blockIdx.x * kUTilesPerCta;
    const int ntile = min(kUTilesPerCta, (int) (n_p / kUN) - jt0);

    if (warp == 8 && lane == 0) {
        ptx::prefetch_tmap(&tmA);
        ptx::prefetch_tmap(&tmB);
        for (int s = 0; s < kUStages; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
        ptx::mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 9) ptx::tmem_alloc(tmem_slot, kUTmemCols);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

  for (int tt = 0; tt < ntile; ++tt) {
    const int j0 = (jt0 + tt) * kUN;
    const int itb = tt * nk;                        // pipeline steps issued before this tile (stage index and phase continue)
    // zero the seven accumulators: warp w < 8 owns lanes 32 (w & 3) .. +31, columns 224 (w >> 2) .. +223
    if (warp < 8) {
        const uint32_t base = tmem + ((uint32_t) ((warp & 3) * 32) << 16) + (uint32_t) ((warp >> 2) * 224);
#pragma unroll
        for (int c = 0; c < 14; ++c) ptx::tmem_st16_zero(base + c * 16);
        ptx::tmem_st_wait();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();

    if (warp == 8) {
        // ===== TMA producer =====
        if (lane == 0) {
            for (int it = 0; it < nk; ++it) {
                const int s = (itb + it) % kUStages;
                const uint32_t ph = (uint32_t) ((itb + it) / kUStages) & 1u;
                ptx::mbar_wait(&empty[s], ph ^ 1u);
                ptx::mbar_expect_tx(&full[s], kUStageBytes);
                uint8_t *dst = smem + s * kUStageBytes;
                ptx::tma_load_3d(dst, &tmA, &full[s], k_byte0 + it * kUK, i0, 4 * q);
                ptx::tma_load_3d(dst + kUABytes, &tmB, &full[s], k_byte0 + it * kUK, j0, 4 * q);
            }
        }
        __syncwarp();
    } else if (warp == 9) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            constexpr uint32_t idesc = STACKED ? ptx::idesc_u8(kUM, 256) : ptx::idesc_u8(kUM, kUN);
            for (int it = 0; it < nk; ++it) {
                const int s = (itb + it) % kUStages;
                const uint32_t ph = (uint32_t) ((itb + it) / kUStages) & 1u;
                ptx::mbar_wait(&full[s], ph);
                ptx::tc_fence_after();
                const uint32_t a_addr = ptx::smem_u32(smem + s * kUStageBytes), b_addr = a_addr + kUABytes;
#pragma unroll
                for (int ks = 0; ks < kUK / 32; ++ks) {
#pragma unroll
                    for (int sl = 0; sl < 4; ++sl) {
                        const uint64_t adesc = ptx::smem_desc_sw64(a_addr + sl * (kUM * kUK) + ks * 32);
                        if (STACKED) {
                            ptx::umma_i8(tmem + sl * kUN, adesc, ptx::smem_desc_sw64(b_addr + ks * 32), idesc, 1u);
                        } else {
#pragma unroll
                            for (int tl = 0; tl < 4; ++tl)
                                ptx::umma_i8(tmem + (sl + tl) * kUN, adesc, ptx::smem_desc_sw64(b_addr + tl * (kUN * kUK) + ks * 32), idesc, 1u);
                        }
                    }
                }
                ptx::umma_commit(&empty[s]);   // frees the stage once these MMAs have read it
            }
            ptx::umma_commit(accum_bar);       // accumulators complete
        }
        __syncwarp();
    } else {
        // ===== epilogue: TMEM -> registers -> sum_u D_u 2^(8u) mod m_q -> S plane =====
        const int m_q = Cp->moduli[q];
        const unsigned long long mu_q = Cp->barrett[q];
        unsigned long long cu[7];
#pragma unroll
        for (int u = 0; u < 7; ++u) cu[u] = (unsigned long long) (unsigned) Cp->pow2[(long long) (8 * u) * Cp->N + q];
        ptx::mbar_wait(accum_bar, (uint32_t) (tt & 1));
        ptx::tc_fence_after();
        const int quad = warp & 3, half = warp >> 2;
        const int i = i0 + quad * 32 + lane;
        int *Sq = S + (long long) q * n_p * m_p;
        const uint32_t tbase = tmem + ((uint32_t) (quad * 32) << 16);
#pragma unroll 1
        for (int jc = 0; jc < 4; ++jc) {
            const int jb = half * 32 + jc * 8;
            uint32_t d[7][8];
#pragma unroll
            for (int u = 0; u < 7; ++u) ptx::tmem_ld8(tbase + (uint32_t) (u * kUN + jb), d[u]);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                int *dst = Sq + (long long) (j0 + jb + e) * m_p + i;
                unsigned long long v = add_to_S ? (unsigned long long) (unsigned) *dst : 0ull;
#pragma unroll
                for (int u = 0; u < 7; ++u) v += (unsigned long long) d[u][e] * cu[u];   // 7 terms < 2^58 each
                *dst = reduce64(v, m_q, mu_q);
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();          // the epilogue has read the accumulators: the next tile may zero them
    ptx::tc_fence_after();
  }
    if (warp == 9) ptx::tmem_dealloc(tmem, kUTmemCols);
}

}  // namespace mpres

// ---- host side: tensor maps --------------------------------------------------------------------------

typedef CUresult (*mpres_encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                          const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline mpres_encode_tiled_fn umma_encode_fn() {
    static mpres_encode_tiled_fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = (mpres_encode_tiled_fn) p;
    }
    return fn;
}

// limb planes [4N][rows_p][k_p] u8 -> 3-D map (k, row, plane), box 64 B x box_rows x 4 planes, 64B swizzle
inline int umma_make_map(CUtensorMap *map, const void *planes, int N, long long rows_p, long long k_p, int box_rows) {
    mpres_encode_tiled_fn enc = umma_encode_fn();
    if (!enc) return -30;
    cuuint64_t dims[3] = {(cuuint64_t) k_p, (cuuint64_t) rows_p, (cuuint64_t) (4 * N)};
    cuuint64_t strides[2] = {(cuuint64_t) k_p, (cuuint64_t) (rows_p * k_p)};
    cuuint32_t box[3] = {(cuuint32_t) mpres::kUK, (cuuint32_t) box_rows, 4};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void *>(planes), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -31;
}

// One launch = all moduli, all tiles, K range [k_begin, k_begin + k_len) (k_len % 64 == 0, <= 8064).
inline int launch_limb_umma(mpres_ctx *c, bool stacked, const uint8_t *PA, const uint8_t *PB, int *S, long long m_p, long long n_p, long long k_p,
                            long long k_begin, int k_len, bool add_to_S, cudaStream_t st) {
    const int N = c->hc.N;
    CUtensorMap tmA, tmB;
    int rc;
    if ((rc = umma_make_map(&tmA, PA, N, m_p, k_p, mpres::kUM))) return rc;
    if ((rc = umma_make_map(&tmB, PB, N, n_p, k_p, mpres::kUN))) return rc;
    if (!c->attr_umma) {
        CUDA_TRY(cudaFuncSetAttribute(mpres::k_limb_umma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mpres::kUSmem));
        CUDA_TRY(cudaFuncSetAttribute(mpres::k_limb_umma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mpres::kUSmem));
        c->attr_umma = true;
    }
    dim3 grid((unsigned) ((n_p / mpres::kUN + mpres::kUTilesPerCta - 1) / mpres::kUTilesPerCta), (unsigned) (m_p / mpres::kUM), (unsigned) N);
    if (stacked)
        mpres::k_limb_umma<true><<<grid, mpres::kUThreads, mpres::kUSmem, st>>>(tmA, tmB, c->dconsts, S, m_p, n_p, (int) k_begin, k_len / mpres::kUK, add_to_S ? 1 : 0, c->d_counter + 2);
    else
        mpres::k_limb_umma<false><<<grid, mpres::kUThreads, mpres::kUSmem, st>>>(tmA, tmB, c->dconsts, S, m_p, n_p, (int) k_begin, k_len / mpres::kUK, add_to_S ? 1 : 0, c->d_counter + 2);
    return 0;
}
