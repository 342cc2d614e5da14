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
    lo.frac = __dmul_rd(lo.frac, C.unit_low.frac); lo.exp += C.unit_low.exp;
    up.frac = __dmul_ru(up.frac, C.unit_upp.frac); up.exp += C.unit_upp.exp;
    er_adjust(lo); er_adjust(up);
}

template <int G, int R>
__global__ void __launch_bounds__(kBinT) k_bin_norm(const DevConsts *Cp, int m, int n, const uint8_t *S8, long long m_ps, long long n_ps, const int *sel,
                                                    const OuterInfo *ia, const OuterInfo *ib, SoA alpha, SoA beta, SoA Cm, int ldc) {
    extern __shared__ __align__(16) uint8_t bin_smem[];
    const int P = sel[0];
    if (P <= 0) return;
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
