// kernels_minplus.cuh -- the (min,+) product of the shift planes from candidate lists.
//
// delta(i,j) = min_l (sa(i,l) + sb(l,j)) is the exponent the reference's mp_add chain ends with (src/arith/add.cuh:172).
// The dense kernel (k_minplus, kernels_fast.cuh) spends k/2 packed add-min operations per entry and sits at the DPX pipe's
// rate.  Here every line keeps its T smallest shifts as candidates (position, value) and the T-th smallest value as its
// threshold.  For a pair (i,j), every l that is a candidate of neither line has sa >= thr_a(i) and sb >= thr_b(j), so
//      c(i,j) = min( min over candidates of row i, min over candidates of column j )
// equals delta(i,j) whenever c(i,j) <= thr_a(i) + thr_b(j) -- 2 T add-mins instead of k.  The pairs that fail the test
// (none with the reference's benchmark inputs: the shifts of a line are geometrically distributed) are listed and
// recomputed densely.  The candidates of a row are applied to whole rows of the TRANSPOSED other plane, so all reads are
// contiguous (and L2-resident: 33 MB per plane at 4096^2).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace mpres {

constexpr int kMcT = 64;          // candidates per line

// ---- T smallest entries of every line (16-bit radix select, one block per line) ---------------------------
__global__ void __launch_bounds__(256) k_mp_select(const int16_t *S, long long pitch, int len, int lines, int *cpos, int *cval, int *thr) {
    __shared__ int hist[256];
    __shared__ int s_bin, s_rem, s_cnt, s_tie;
    const int line = blockIdx.x;
    if (line >= lines) return;
    const int16_t *row = S + (long long) line * pitch;
    // pass 1: high byte
    hist[threadIdx.x] = 0;
    __syncthreads();
    for (int l = threadIdx.x; l < len; l += 256) atomicAdd(&hist[((int) row[l] >> 8) & 255], 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0, b = 0;
        for (; b < 256; ++b) { if (acc + hist[b] >= kMcT) break; acc += hist[b]; }
        s_bin = b; s_rem = kMcT - acc;
    }
    __syncthreads();
    const int b1 = s_bin, rem1 = s_rem;
    __syncthreads();
    // pass 2: low byte inside that bin
    hist[threadIdx.x] = 0;
    __syncthreads();
    for (int l = threadIdx.x; l < len; l += 256) { const int v = row[l]; if (((v >> 8) & 255) == b1) atomicAdd(&hist[v & 255], 1); }
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0, b = 0;
        for (; b < 256; ++b) { if (acc + hist[b] >= rem1) break; acc += hist[b]; }
        s_bin = (b1 << 8) | b; s_rem = rem1 - acc; s_cnt = 0; s_tie = 0;
    }
    __syncthreads();
    const int vstar = s_bin, ties = s_rem;
    for (int l = threadIdx.x; l < len; l += 256) {
        const int v = row[l];
        bool take = v < vstar;
        if (v == vstar) take = atomicAdd(&s_tie, 1) < ties;
        if (take) {
            const int slot = atomicAdd(&s_cnt, 1);
            cpos[(long long) line * kMcT + slot] = l;
            cval[(long long) line * kMcT + slot] = v;
        }
    }
    if (threadIdx.x == 0) thr[line] = vstar;
}

// ---- transpose of a shift plane: T[l][o] = S[o][l] --------------------------------------------------------
__global__ void __launch_bounds__(256) k_mp_transpose(const int16_t *S, long long pitch_s, int16_t *T, long long pitch_t) {
    __shared__ int16_t tile[64][66];
    const int o0 = blockIdx.x * 64, l0 = blockIdx.y * 64;
    for (int t = threadIdx.x; t < 64 * 32; t += 256) {
        const int r = t >> 5, w = t & 31;
        const uint32_t v = *(const uint32_t *) (S + (long long) (o0 + r) * pitch_s + l0 + 2 * w);
        tile[r][2 * w] = (int16_t) (v & 0xffffu);
        tile[r][2 * w + 1] = (int16_t) (v >> 16);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < 64 * 32; t += 256) {
        const int r = t >> 5, w = t & 31;   // output row l0 + r, columns o0 + 2w, 2w + 1
        const uint32_t v = (uint32_t) (uint16_t) tile[2 * w][r] | ((uint32_t) (uint16_t) tile[2 * w + 1][r] << 16);
        *(uint32_t *) (T + (long long) (l0 + r) * pitch_t + o0 + 2 * w) = v;
    }
}

// ---- candidates of a line applied to rows of the transposed other plane ---------------------------------
// D[line][c] = min_t (cval[line][t] + OT[cpos[line][t]][c]) for c < cols (cols % 8 == 0).  One block per line.
__global__ void __launch_bounds__(256) k_mp_gather(const int *cpos, const int *cval, const int16_t *OT, long long pitch_ot, int lines, int cols,
                                                   int16_t *D, long long pitch_d) {
    __shared__ int s_pos[kMcT];
    __shared__ uint32_t s_add[kMcT];
    const int line = blockIdx.x;
    if (line >= lines) return;
    if (threadIdx.x < kMcT) {
        s_pos[threadIdx.x] = cpos[(long long) line * kMcT + threadIdx.x];
        const uint32_t v = (uint32_t) cval[(long long) line * kMcT + threadIdx.x] & 0xffffu;
        s_add[threadIdx.x] = v | (v << 16);
    }
    __syncthreads();
    for (int c8 = threadIdx.x; c8 * 8 < cols; c8 += 256) {
        uint32_t acc[4] = {0x7fff7fffu, 0x7fff7fffu, 0x7fff7fffu, 0x7fff7fffu};
#pragma unroll 8
        for (int t = 0; t < kMcT; ++t) {
            const uint4 v = __ldg((const uint4 *) (OT + (long long) s_pos[t] * pitch_ot) + c8);
            const uint32_t ad = s_add[t];
            acc[0] = __viaddmin_s16x2(v.x, ad, acc[0]);
            acc[1] = __viaddmin_s16x2(v.y, ad, acc[1]);
            acc[2] = __viaddmin_s16x2(v.z, ad, acc[2]);
            acc[3] = __viaddmin_s16x2(v.w, ad, acc[3]);
        }
        *((uint4 *) (D + (long long) line * pitch_d) + c8) = make_uint4(acc[0], acc[1], acc[2], acc[3]);
    }
}

// ---- delta[j][i] = min(D2[j][i], D1[i][j]); pairs above thr_a(i) + thr_b(j) are listed for the dense kernel ---------
__global__ void __launch_bounds__(256) k_mp_combine(const int16_t *D1, long long pitch1, const int16_t *D2, long long pitch2, const int *thrA, const int *thrB,
                                                    int m, int n, int16_t *delta, long long m_p, long long *list, int *count) {
    __shared__ int16_t tile[64][66];
    const int i0 = blockIdx.x * 64, j0 = blockIdx.y * 64;
    for (int t = threadIdx.x; t < 64 * 32; t += 256) {      // D1 tile: rows i, columns j
        const int r = t >> 5, w = t & 31;
        uint32_t v = 0x7fff7fffu;
        if (i0 + r < m) v = *(const uint32_t *) (D1 + (long long) (i0 + r) * pitch1 + j0 + 2 * w);
        tile[r][2 * w] = (int16_t) (v & 0xffffu);
        tile[r][2 * w + 1] = (int16_t) (v >> 16);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < 64 * 32; t += 256) {
        const int r = t >> 5, w = t & 31;                   // row j0 + r, columns i0 + 2w, 2w + 1
        const int j = j0 + r;
        if (j >= n) continue;
        const uint32_t v2 = *(const uint32_t *) (D2 + (long long) j * pitch2 + i0 + 2 * w);
        const int tb = thrB[j];
        int out[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int i = i0 + 2 * w + h;
            const int a = (int) (int16_t) (h ? (v2 >> 16) : (v2 & 0xffffu));
            const int b = (int) tile[2 * w + h][r];
            int c = a < b ? a : b;
            if (i < m && c > thrA[i] + tb) {                // not covered by the candidates: dense recomputation
                const int slot = atomicAdd(count, 1);
                list[slot] = ((long long) j << 32) | (unsigned) i;
            }
            out[h] = c;
        }
        *(uint32_t *) (delta + (long long) j * m_p + i0 + 2 * w) = ((uint32_t) out[0] & 0xffffu) | ((uint32_t) out[1] << 16);
    }
}

// ---- dense recomputation of the listed pairs, one warp each ----------------------------------------------------
__global__ void __launch_bounds__(256) k_mp_fix(const int16_t *SA, const int16_t *SB, long long inner_p, int16_t *delta, long long m_p,
                                                const long long *list, const int *count) {
    const int lane = threadIdx.x & 31;
    const long long nw = (long long) gridDim.x * (blockDim.x >> 5);
    const int total = *count;
    for (long long e = ((long long) blockIdx.x * blockDim.x + threadIdx.x) >> 5; e < total; e += nw) {
        const long long pr = list[e];
        const int j = (int) (pr >> 32), i = (int) (pr & 0xffffffffll);
        const uint32_t *a = (const uint32_t *) (SA + (long long) i * inner_p), *b = (const uint32_t *) (SB + (long long) j * inner_p);
        uint32_t acc = 0x7fff7fffu;
        for (long long w = lane; w < inner_p / 2; w += 32) acc = __viaddmin_s16x2(a[w], b[w], acc);
        int v = min((int) (int16_t) (acc & 0xffffu), (int) (int16_t) (acc >> 16));
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
        if (lane == 0) delta[(long long) j * m_p + i] = (int16_t) v;
    }
}

}  // namespace mpres
