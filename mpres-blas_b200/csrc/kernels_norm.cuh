// kernels_norm.cuh -- stage 3 of the fast mp_gemm path: normalisation of the exact per-modulus sums and
// the alpha/beta epilogue of src/blas/gemm.cuh:142-166, fused.
//
// Three kernels share one definition of the arithmetic:
//   k_norm_fast<NQ>         ONE THREAD per C entry, moduli walked serially.  The sign / exponent /
//                           interval work of mp_mul and mp_add (four FP64 divisions, the alignment
//                           logic) is done once per entry instead of once per lane of a lane group,
//                           which is what made the residue-parallel kernel issue-bound (~1.2 k warp
//                           instructions per entry, round-1 ncu).  It handles the common case -- one
//                           magnification resolves sign and interval, no rounding fires, the sign of
//                           alpha*S + beta*C is decided by the intervals -- and appends every other
//                           entry to a list.  The FP64 sums use the same balanced tree as the lane
//                           butterfly, so its results are the same bits as the residue-parallel code.
//   k_norm_list<G, R>       the listed entries, one lane group each, full generality (refinement
//                           rounds, rns_scale2pow rounding, mixed-radix sign resolution).
//   k_normalize_epilogue    the round-1 tile kernel (all entries residue-parallel); still used for
//                           moduli counts without a k_norm_fast instantiation.
#pragma once

namespace mpres {

// Sign and interval evaluation of an exact sum S known only through its residues X = S mod M, given
// |S| < 2^bound <= M/4.  Instead of the reference's generic refinement (src/rns.cuh:911-921: up to
// log2(M)/49 rounds, each with a log2 and a ceil) the known bound gives the magnification directly:
// X * 2^K mod M has its fractional value in (0, 1/8) for S > 0 and in (7/8, 1) for S < 0.  Further
// rounds (only after heavy cancellation) magnify by what the current upper bound allows.
// Returns 0 for S == 0, else +1 / -1, with [lo, up] enclosing |S| / M.
template <int G, int R>
__device__ __forceinline__ int sign_eval_window(const DevConsts &C, const Lane<R> &L, const int (&x)[R], int bound, Er &lo, Er &up) {
    int nzbits = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) nzbits |= x[r];
    if (gor<G>(nzbits) == 0) { lo.frac = 0; lo.exp = 0; up.frac = 0; up.exp = 0; return 0; }
    int K = C.log2M - bound - 3;
    K = K < 0 ? 0 : K;
    int s[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        int c = mulmod(x[r], L.w[r], L.m[r], L.mu[r]);
        s[r] = L.act[r] ? mulmod(c, __ldg(C.pow2 + (long long) K * C.N + L.idx[r]), L.m[r], L.mu[r]) : 0;
    }
    for (int iter = 0; iter < 200; ++iter) {
        double fl[R], fu[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            fl[r] = __dmul_rd((double) s[r], L.rrd[r]);
            fu[r] = __dmul_ru((double) s[r], L.rru[r]);
        }
        const double suml = gsum_dir<G, R, false>(fl), sumu = gsum_dir<G, R, true>(fu);
        const double wl = floor(suml), wu = floor(sumu);
        const double dl = __dsub_rd(suml, wl), du = __dsub_ru(sumu, wu);   // exact
        double dist;   // upper bound of the distance of the fraction to the nearest integer
        if (wl == wu) {
            if (du < 0.25 && dl >= C.accuracy) {
                lo = er_from_double(dl); up = er_from_double(du);
                lo.exp -= K; up.exp -= K;
                return 1;
            }
            const double ml = __dsub_rd(1.0, du), mh = __dsub_ru(1.0, dl);
            if (dl > 0.75 && ml >= C.accuracy) {
                lo = er_from_double(ml); up = er_from_double(mh);
                lo.exp -= K; up.exp -= K;
                return -1;
            }
            if (du >= 0.25 && dl <= 0.75) {   // outside both windows: the guard did not hold (MODE_FAST only)
                lo = er_from_double(dl); up = er_from_double(du);
                lo.exp -= K; up.exp -= K;
                return 1;
            }
            dist = du < 0.25 ? du : mh;
        } else {
            dist = __dadd_ru(du, __dsub_ru(1.0, dl));   // straddles an integer
        }
        // |frac| <= dist < 2^(e+1): magnify by -(e+1) - 3 bits, keeping the value below 1/8
        int e = (int) (((unsigned long long) __double_as_longlong(dist) >> 52) & 0x7ff) - 1023;
        int kk = -(e + 1) - 3;
        kk = kk < 1 ? 1 : (kk > 60 ? 60 : kk);
        if (K + kk > C.log2M) kk = C.log2M - K;
        if (kk <= 0) break;
#pragma unroll
        for (int r = 0; r < R; ++r)
            s[r] = L.act[r] ? mulmod(s[r], __ldg(C.pow2 + (long long) kk * C.N + L.idx[r]), L.m[r], L.mu[r]) : 0;
        K += kk;
    }
    // not resolvable by magnification (cannot happen for |S| >= 1): fall back to the generic evaluation
    eval_compute<G, R, false>(C, L, x, lo, up);
    return (lo.frac != 0 && lo.exp >= -1) ? -1 : 1;
}


// Everything after the residues of S(row, col) are in registers: sign / interval by magnification,
// exact division by 2^d, one rounding if needed, then C = round(round(beta*C) + round(alpha*S)).
// `bound`: |S| < 2^bound.  d: the power of two separating our exponent base from the reference's.
template <int G, int R>
__device__ __forceinline__ void normalize_entry(const DevConsts &C, const Lane<R> &L, Num<R> &s, long long bound, int d, int exp_base,
                                                const SoA &alpha, const SoA &beta, const SoA &Cm, long long ic) {
    const int N = C.N;
    if (d < kShiftSentinel) {
        d = d > C.log2M ? C.log2M : d;   // only reachable in MODE_FAST with a failed guard
        bound = bound > C.log2M - 2 ? C.log2M - 2 : bound;
        Er lo, up;
        const int sg = sign_eval_window<G, R>(C, L, s.d, (int) bound, lo, up);
        if (sg != 0) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                int v = s.d[r];
                if (sg < 0 && v) v = L.m[r] - v;
                // exact division by 2^d (every term carries at least d trailing zero bits)
                s.d[r] = L.act[r] ? mulmod(v, __ldg(C.inv_pow2 + (long long) d * N + L.idx[r]), L.m[r], L.mu[r]) : 0;
            }
            s.sign = sg < 0 ? 1 : 0;
            s.exp = exp_base + d;
            s.lo = lo; s.up = up;
            s.lo.exp -= d; s.up.exp -= d;
            round_if_needed<G, R>(C, L, s);
        } else {
            num_zero(s);
        }
    } else {
        num_zero(s);
    }
    Num<R> al, be, c, t1, t2;
    load_num<G, R>(C, L, alpha, 0, al);
    load_num<G, R>(C, L, beta, 0, be);
    load_num<G, R>(C, L, Cm, ic, c);
    mp_mul<G, R, true>(C, L, t1, s, al);
    mp_mul<G, R, true>(C, L, t2, c, be);
    mp_add<G, R, true>(C, L, c, t2, t1);
    store_num<G, R>(C, L, Cm, ic, c);
}

__device__ __forceinline__ int ceil_log2(int k) {
    int lgk = 0;
    while ((1 << lgk) < k) ++lgk;
    return lgk;
}

// ---- base extension (reduced-base fast path) -----------------------------------------------------------
// Stage 2 produced S mod m_q for the first np moduli only.  With M' = m_0 ... m_{np-1}, M'_i = M'/m_i and
// xi_i = x_i (M'_i)^-1 mod m_i, the Chinese remainder theorem gives  S = sum_i xi_i M'_i - R M'  for an
// integer R.  Because |S| < M'/4 (k_choose_base), R is the integer NEAREST to sum_i xi_i / m_i -- a double
// sum (error < np 2^-52) decides it with a margin of 1/4, whatever the sign and however small S is.  Hence
//      S mod m_q = ( sum_i xi_i (M'_i mod m_q) + R (m_q - M' mod m_q) ) mod m_q          for q >= np.
// (The reference has no base extension; its rank-based CRT reconstruction in rns_scale2pow,
// src/rns.cuh:1061-1124, is the same identity with a floor instead of the nearest integer.)
// One thread per entry, the xi_i in registers, tables staged per block.  Writes the planes q >= np of S.
constexpr int kExtThreads = 128;
inline size_t base_extend_smem(int N) { return (size_t) N * N * sizeof(int) + (size_t) N * 40; }

template <int NP>
__device__ __forceinline__ void base_extend_body(int np, int N, bool lazy, int *Sp, long long plane, const int *s_m, const unsigned long long *s_mu,
                                                 const int *s_w, const double *s_rcp, const int *s_negmp, const int *s_t) {
    unsigned xi[NP];
    double sum = 0.0;
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        xi[i] = 0;
        if (i < np) {
            xi[i] = (unsigned) mulmod(Sp[i * plane], s_w[i], s_m[i], s_mu[i]);
            sum += (double) xi[i] * s_rcp[i];
        }
    }
    const unsigned R = (unsigned) __double2int_rn(sum);
    for (int q = np; q < N; ++q) {
        const int mq = s_m[q];
        const unsigned long long muq = s_mu[q];
        const int4 *t4 = (const int4 *) (s_t + (q - np) * NP);
        unsigned long long acc = 0;
#pragma unroll
        for (int i4 = 0; i4 < NP / 4; ++i4) {
            const int4 t = t4[i4];
            acc += (unsigned long long) xi[4 * i4] * (unsigned) t.x;
            acc += (unsigned long long) xi[4 * i4 + 1] * (unsigned) t.y;
            acc += (unsigned long long) xi[4 * i4 + 2] * (unsigned) t.z;
            acc += (unsigned long long) xi[4 * i4 + 3] * (unsigned) t.w;
            if (!lazy) acc = (unsigned long long) (unsigned) reduce64(acc, mq, muq);   // moduli up to 2^31: four terms fit
        }
        if (!lazy) acc = (unsigned long long) (unsigned) reduce64(acc, mq, muq);
        acc += (unsigned long long) R * (unsigned) s_negmp[q];
        Sp[q * plane] = reduce64(acc, mq, muq);
    }
}

__global__ void __launch_bounds__(kExtThreads, 6) k_base_extend(const DevConsts *Cp, int m, int n, int *S, long long m_p, long long n_p, const int *nprime) {
    extern __shared__ __align__(16) unsigned char ext_smem[];
    const DevConsts &C = *Cp;
    const int N = C.N;
    const int np = *nprime;
    if (np >= N || np <= 0) return;   // full base, or the small-modulus path was selected
    const int np4 = (np + 3) & ~3;
    unsigned long long *s_mu = (unsigned long long *) ext_smem;      // [N]
    double *s_rcp = (double *) (s_mu + N);                             // [N]   1 / m_i
    int *s_m = (int *) (s_rcp + N);                                    // [N]
    int *s_w = s_m + N;                                                // [N]   (M'_i)^-1 mod m_i
    int *s_negmp = s_w + N;                                            // [N]   m_q - M' mod m_q
    int *s_t = s_negmp + N + ((4 - ((3 * N) & 3)) & 3);               // [N - np][np4]  M'_i mod m_q, 16-byte aligned
    for (int t = threadIdx.x; t < N; t += kExtThreads) {
        const int mq = C.moduli[t];
        s_mu[t] = C.barrett[t]; s_m[t] = mq; s_rcp[t] = 1.0 / (double) mq;
        s_w[t] = C.ext_w[np * N + t];
        const int mp = C.prefix_mod[np * N + t];
        s_negmp[t] = mp ? mq - mp : 0;
    }
    for (int t = threadIdx.x; t < (N - np) * np4; t += kExtThreads) {
        const int q = np + t / np4, i = t % np4;
        s_t[t] = i < np ? C.ext_t[((long long) np * N + q) * N + i] : 0;
    }
    __syncthreads();
    const int row = blockIdx.x * kExtThreads + threadIdx.x;
    if (row >= m) return;
    const long long plane = n_p * m_p;
    const bool lazy = C.ext_lazy != 0;
    // a block walks the columns blockIdx.y, blockIdx.y + gridDim.y, ... (the tables are staged once)
    for (int col = blockIdx.y; col < n; col += gridDim.y) {
        int *Sp = S + (long long) col * m_p + row;
        switch (np4) {
            case 4: base_extend_body<4>(np, N, lazy, Sp, plane, s_m, s_mu, s_w, s_rcp, s_negmp, s_t); break;
            case 8: base_extend_body<8>(np, N, lazy, Sp, plane, s_m, s_mu, s_w, s_rcp, s_negmp, s_t); break;
            case 12: base_extend_body<12>(np, N, lazy, Sp, plane, s_m, s_mu, s_w, s_rcp, s_negmp, s_t); break;
            case 16: base_extend_body<16>(np, N, lazy, Sp, plane, s_m, s_mu, s_w, s_rcp, s_negmp, s_t); break;
            case 20: base_extend_body<20>(np, N, lazy, Sp, plane, s_m, s_mu, s_w, s_rcp, s_negmp, s_t); break;
            case 24: base_extend_body<24>(np, N, lazy, Sp, plane, s_m, s_mu, s_w, s_rcp, s_negmp, s_t); break;
            case 28: base_extend_body<28>(np, N, lazy, Sp, plane, s_m, s_mu, s_w, s_rcp, s_negmp, s_t); break;
            case 32: base_extend_body<32>(np, N, lazy, Sp, plane, s_m, s_mu, s_w, s_rcp, s_negmp, s_t); break;
            case 36: base_extend_body<36>(np, N, lazy, Sp, plane, s_m, s_mu, s_w, s_rcp, s_negmp, s_t); break;
            case 40: base_extend_body<40>(np, N, lazy, Sp, plane, s_m, s_mu, s_w, s_rcp, s_negmp, s_t); break;
            case 44: base_extend_body<44>(np, N, lazy, Sp, plane, s_m, s_mu, s_w, s_rcp, s_negmp, s_t); break;
            case 48: base_extend_body<48>(np, N, lazy, Sp, plane, s_m, s_mu, s_w, s_rcp, s_negmp, s_t); break;
            default: break;   // k_choose_base never selects a reduced base above kMaxReducedBase
        }
    }
}

// ---- round-1 tile kernel: every entry residue-parallel ---------------------------------------------
// One lane group per C entry; a block covers kNormTile consecutive rows of one column so the
// per-modulus planes are read as contiguous runs and transposed through shared memory (256 / G rows).
template <int G, int R>
__global__ void __launch_bounds__(256) k_normalize_epilogue(const DevConsts *Cp, int m, int n, int k, const int *S, const int16_t *delta,
                                                 long long m_p, long long n_p, const OuterInfo *ia, const OuterInfo *ib,
                                                 SoA alpha, SoA beta, SoA Cm, int ldc, long long *todo, int *todo_count, bool fallback_allowed) {
    extern __shared__ int sm_res[];   // [N][kNormTile + 1]
    constexpr int kNormTile = 256 / G;
    const DevConsts &C = *Cp;
    const int N = C.N;
    Lane<R> L;
    lane_init<G, R>(C, L);
    const int tiles = (m + kNormTile - 1) / kNormTile;
    const int col = blockIdx.x / tiles;
    const int row0 = (blockIdx.x - col * tiles) * kNormTile;
    for (int t = threadIdx.x; t < N * kNormTile; t += blockDim.x) {
        const int q = t / kNormTile, r = t - q * kNormTile;
        sm_res[q * (kNormTile + 1) + r] = S[((long long) q * n_p + col) * m_p + row0 + r];
    }
    __syncthreads();
    const int grp = threadIdx.x / G;
    const int row = row0 + grp;
    if (row >= m) return;
    const OuterInfo ra = ia[row], cb = ib[col];
    const int lgk = ceil_log2(k);
    Num<R> s;
    num_zero(s);
    int d = kShiftSentinel;
    long long bound = 0;
    if (ra.win >= 0 && cb.win >= 0) {
        bound = (long long) ra.win + cb.win + lgk;
        if (bound > (long long) C.log2M - 2) {
            // window guard failed: exact accumulation not guaranteed -> reference-order recomputation
            if ((threadIdx.x & (G - 1)) == 0) {
                int pos = atomicAdd(todo_count, 1);
                if (fallback_allowed) todo[pos] = row + (long long) col * m;
            }
            if (fallback_allowed) return;
        }
        d = delta[(long long) col * m_p + row];
        if (d < kShiftSentinel) {
#pragma unroll
            for (int r = 0; r < R; ++r) s.d[r] = L.act[r] ? sm_res[L.idx[r] * (kNormTile + 1) + grp] : 0;
        }
    }
    normalize_entry<G, R>(C, L, s, bound, d, ra.emin + cb.emin, alpha, beta, Cm, row + (long long) col * ldc);
}

// ---- listed entries, residue-parallel ----------------------------------------------------------------
template <int G, int R>
__global__ void __launch_bounds__(256) k_norm_list(const DevConsts *Cp, int m, int n, int k, const int *S, const int16_t *delta,
                                                   long long m_p, long long n_p, const OuterInfo *ia, const OuterInfo *ib,
                                                   SoA alpha, SoA beta, SoA Cm, int ldc, const long long *list, const int *list_count) {
    const DevConsts &C = *Cp;
    Lane<R> L;
    lane_init<G, R>(C, L);
    long long grp = ((long long) blockIdx.x * blockDim.x + threadIdx.x) / G;
    const long long ngrp = (long long) gridDim.x * blockDim.x / G;
    const long long total = *list_count;
    const int lgk = ceil_log2(k);
    for (; grp < total; grp += ngrp) {
        const long long e = list[grp];
        const int row = (int) (e % m), col = (int) (e / m);
        const OuterInfo ra = ia[row], cb = ib[col];
        Num<R> s;
        num_zero(s);
        int d = kShiftSentinel;
        long long bound = 0;
        if (ra.win >= 0 && cb.win >= 0) {
            bound = (long long) ra.win + cb.win + lgk;
            d = delta[(long long) col * m_p + row];
            if (d < kShiftSentinel) {
#pragma unroll
                for (int r = 0; r < R; ++r) s.d[r] = L.act[r] ? __ldg(S + ((long long) L.idx[r] * n_p + col) * m_p + row) : 0;
            }
        }
        normalize_entry<G, R>(C, L, s, bound, d, ra.emin + cb.emin, alpha, beta, Cm, row + (long long) col * ldc);
    }
}

// ---- entry-per-thread kernel ---------------------------------------------------------------------------
struct QConst {          // per-modulus constants, broadcast from shared memory
    int m;
    unsigned mu32;       // floor(2^(kb + 30) / m) when every modulus has the same bit length kb <= 27 (barrett_k), else 0
    unsigned long long mu;
    double rrd, rru;
};

// Per-call tables for k_norm_fast: alpha * 2^j mod m_q for j = -log2M .. log2M (row j + log2M) and
// beta * 2^j mod m_q for j = 0 .. log2M (rows 2 log2M + 1 + j).  They turn "multiply by the scalar, then by
// the alignment power of two" (mp_mul followed by the shift inside mp_add) into one modular multiplication.
__global__ void k_scalar_tables(const DevConsts *Cp, SoA alpha, SoA beta, int *tab) {
    const DevConsts &C = *Cp;
    const int N = C.N, L = C.log2M;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (3 * L + 2) * N) return;
    const int row = t / N, q = t - row * N;
    const int m = C.moduli[q];
    const unsigned long long mu = C.barrett[q];
    int v;
    if (row <= 2 * L) {
        const int j = row - L;
        v = mulmod(alpha.digits[q], j >= 0 ? C.pow2[(long long) j * N + q] : C.inv_pow2[(long long) (-j) * N + q], m, mu);
    } else {
        v = mulmod(beta.digits[q], C.pow2[(long long) (row - 2 * L - 1) * N + q], m, mu);
    }
    tab[t] = v;
}
struct ScalarEsi { int sign, exp; Er lo, up; };

// a * b mod m for canonical a, b.  F32: every modulus has bit length kb (24 <= kb <= 27), so the product is below
// 2^(2 kb) and the one-correction 32-bit Barrett step applies (mp_device.cuh: barrett_k); otherwise the generic 64-bit
// reduction.  Both return the canonical residue, hence identical results.
template <bool F32>
__device__ __forceinline__ int mulmod_q(int a, int b, const QConst &c, int kb) {
    if (F32) return (int) barrett_k((unsigned long long) (unsigned) a * (unsigned) b, (unsigned) c.m, c.mu32, kb);
    return mulmod(a, b, c.m, c.mu);
}
__host__ __device__ constexpr int pow2ceil_c(int n) { int p = 1; while (p < n) p <<= 1; return p; }
__host__ __device__ constexpr int log2_c(int p) { int l = 0; while ((1 << l) < p) ++l; return l; }
__host__ __device__ constexpr int trailing_ones_c(int q) { int t = 0; while (q & 1) { ++t; q >>= 1; } return t; }

constexpr int kNormFastThreads = 128;
#ifndef MPRES_NORM_ROUNDS
#define MPRES_NORM_ROUNDS 2
#endif
constexpr int kNormFastRounds = MPRES_NORM_ROUNDS;

// The body of the entry-per-thread kernel.  One block = kNormFastThreads consecutive rows (from row0) of column col.
// Sp[q * plane] is residue q of the exact sum of this thread's entry: in the residue planes of stage 2 (global memory),
// or -- SMEM_S -- in shared memory where the fused base-extension kernel (kernels_small.cuh) left it; then the residues of the
// entries handed to the list kernel are copied to the global planes (Sg, stride gplane), which k_norm_list reads.
template <int NQ, bool F32, bool SMEM_S>
__device__ __forceinline__ void norm_fast_body(const DevConsts &C, int *cds, int m, int n, int k, int col, int row0, const int *Sp, long long plane,
                                               int *Sg, long long gplane, const int16_t *delta, long long m_p, const OuterInfo *ia, const OuterInfo *ib,
                                               SoA alpha, SoA beta, SoA Cm, int ldc, const int *scal_tab, long long *todo, int *todo_count,
                                               long long *slow, int *slow_count, bool fallback_allowed) {
    constexpr int P = pow2ceil_c(NQ), LOGP = log2_c(P);
    constexpr int CP = NQ + 1;              // pitch of the staged C digits (conflict-free per-thread rows)
    // cds: [kNormFastThreads][CP]: digits of C in, digits of the result out
    __shared__ QConst qc[NQ];
    __shared__ ScalarEsi s_al, s_be;
    __shared__ unsigned char okf[kNormFastThreads];
    for (int q = threadIdx.x; q < NQ; q += blockDim.x) {
        QConst c;
        c.m = C.moduli[q]; c.mu32 = F32 ? C.small->red_mu[q] : 0u; c.mu = C.barrett[q]; c.rrd = C.recip_rd[q]; c.rru = C.recip_ru[q];
        qc[q] = c;
    }
    const int kb = F32 ? C.small->red_shift : 0;
    if (threadIdx.x == 0) { s_al.sign = alpha.sign[0]; s_al.exp = alpha.exp[0]; s_al.lo = alpha.eval[0]; s_al.up = alpha.eval[alpha.len()]; }
    if (threadIdx.x == 32) { s_be.sign = beta.sign[0]; s_be.exp = beta.exp[0]; s_be.lo = beta.eval[0]; s_be.up = beta.eval[beta.len()]; }
    const int row = row0 + threadIdx.x;
    const bool live = row < m;
    // the digits of this block's C entries are one contiguous run: asynchronous 4-byte copies into the padded rows (coalesced: a
    // warp copies 128 consecutive bytes per instruction); they are only waited for before the digit pass, so the DRAM latency hides
    // behind the sign / interval work
    const int rows_live = min(kNormFastThreads, m - row0);
    int4 *cd4 = (int4 *) (Cm.digits + (row0 + (long long) col * ldc) * NQ);
    {
        const int *csrc = (const int *) cd4;
        for (int v = threadIdx.x; v < rows_live * NQ; v += kNormFastThreads) {
            const unsigned sa = (unsigned) __cvta_generic_to_shared(cds + (v / NQ) * CP + v % NQ);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(sa), "l"(csrc + v));
        }
        asm volatile("cp.async.commit_group;\n" ::);
    }
    okf[threadIdx.x] = 0;
    // the entry's line descriptors and shift go out before the wait for the staged residues, and the scalar fields of C (needed after
    // the first pass) are pulled towards L1 meanwhile: their latencies overlap instead of following one another
    OuterInfo ra = {}, cb = {};
    int d = 0;
    long long ic = 0;
    if (live) {
        ra = ia[row]; cb = ib[col];
        d = delta[(long long) col * m_p + row];
        ic = row + (long long) col * ldc;
        asm volatile("prefetch.global.L1 [%0];\n" ::"l"(Cm.eval + ic));
        asm volatile("prefetch.global.L1 [%0];\n" ::"l"(Cm.eval + ic + Cm.len()));
        asm volatile("prefetch.global.L1 [%0];\n" ::"l"(Cm.exp + ic));
        asm volatile("prefetch.global.L1 [%0];\n" ::"l"(Cm.sign + ic));
    }
    if (SMEM_S) asm volatile("cp.async.wait_group 1;\n" ::: "memory");        // the staged residues of S (an older group than the digits of C)
    __syncthreads();
    bool to_slow = false, to_todo = false, go = false;
    int sg = 0, sign = 0;
    AddEsi p;
    Er rlo, rup;
    const int log2M = C.log2M;
    if (live) do {
        const int mp_h = C.mp_h;
        if (ra.win < 0 || cb.win < 0) { to_slow = true; break; }             // a line of exact zeros: S == 0
        const long long bound = (long long) ra.win + cb.win + ceil_log2(k);
        if (bound > (long long) log2M - 2) { to_todo = true; to_slow = !fallback_allowed; break; }   // window guard failed
        if (d >= kShiftSentinel) { to_slow = true; break; }                   // every term is an exact zero
        int K = log2M - (int) bound - 3;
        K = K < 0 ? 0 : K;
        // ---- pass 1 over the moduli: residues, magnified fractions, directed sums (balanced tree) ----
        // kNormFastRounds magnification rounds with exactly the decisions of sign_eval_window (further rounds are needed after
        // cancellation, when the first magnified fraction is below the accuracy threshold); whatever is still open goes to the list.
        // Measured on B200 at config 3: two unrolled rounds take the list from 2.2 % of the entries to none and the step from 13.07 to
        // 12.58 ms; a four-round loop spills and costs more than the list kernel did.
        Er lo, up;
        bool open = true;
#pragma unroll
        for (int round = 0; round < kNormFastRounds; ++round) {
            int nz = 0;
            double stl[LOGP + 1], stu[LOGP + 1];
            {
                const int4 *p2 = (const int4 *) (C.wpow2 + (long long) K * NQ);   // w_q 2^K mod m_q, four moduli per load
                int4 t4 = make_int4(0, 0, 0, 0);
#pragma unroll
                for (int q = 0; q < P; ++q) {
                    double vl = 0.0, vu = 0.0;
                    if (q < NQ) {
                        if ((q & 3) == 0) t4 = __ldg(p2 + (q >> 2));
                        const QConst c = qc[q];
                        const int xq = SMEM_S ? Sp[q * plane] : __ldg(Sp + q * plane);
                        nz |= xq;
                        const int sq = mulmod_q<F32>(xq, (q & 3) == 0 ? t4.x : (q & 3) == 1 ? t4.y : (q & 3) == 2 ? t4.z : t4.w, c, kb);
                        vl = __dmul_rd((double) sq, c.rrd);
                        vu = __dmul_ru((double) sq, c.rru);
                    }
                    const int t1s = trailing_ones_c(q);
#pragma unroll
                    for (int b = 0; b < LOGP; ++b)
                        if (b < t1s) { vl = __dadd_rd(stl[b], vl); vu = __dadd_ru(stu[b], vu); }
                    stl[t1s] = vl; stu[t1s] = vu;
                }
            }
            if (nz == 0) break;                                                   // S == 0: the list kernel writes the zero
            const double suml = stl[LOGP], sumu = stu[LOGP];
            const double wl = floor(suml), wu = floor(sumu);
            const double dl = __dsub_rd(suml, wl), du = __dsub_ru(sumu, wu);
            double dist;
            if (wl == wu) {
                if (du < 0.25 && dl >= C.accuracy) { sg = 1; lo = er_from_double(dl); up = er_from_double(du); open = false; break; }
                const double ml = __dsub_rd(1.0, du), mh = __dsub_ru(1.0, dl);
                if (dl > 0.75 && ml >= C.accuracy) { sg = -1; lo = er_from_double(ml); up = er_from_double(mh); open = false; break; }
                if (du >= 0.25 && dl <= 0.75) break;                              // outside both windows (MODE_FAST with a failed guard)
                dist = du < 0.25 ? du : mh;
            } else {
                dist = __dadd_ru(du, __dsub_ru(1.0, dl));                         // straddles an integer
            }
            const int e = (int) (((unsigned long long) __double_as_longlong(dist) >> 52) & 0x7ff) - 1023;
            int kk = -(e + 1) - 3;
            kk = kk < 1 ? 1 : (kk > 60 ? 60 : kk);
            if (K + kk > log2M) kk = log2M - K;
            if (kk <= 0) break;
            K += kk;
        }
        if (open) { to_slow = true; break; }
        lo.exp -= K + d; up.exp -= K + d;
        if (up.exp >= mp_h) { to_slow = true; break; }                        // S itself needs a rounding
        const int s_sign = sg < 0 ? 1 : 0, s_exp = ra.emin + cb.emin + d;
        // ---- t1 = alpha * S, t2 = beta * C (mp_mul, src/arith/mul.cuh:53-61), exponent/sign/interval part ----
        const ScalarEsi al = s_al, be = s_be;
        const Er t1lo = er_md_dir<false>(lo, al.lo, C.unit_upp), t1up = er_md_dir<true>(up, al.up, C.unit_low);
        if (t1up.frac != 0 && t1up.exp >= mp_h) { to_slow = true; break; }
        const Er clo = Cm.eval[ic], cup = Cm.eval[ic + Cm.len()];
        const Er t2lo = er_md_dir<false>(clo, be.lo, C.unit_upp), t2up = er_md_dir<true>(cup, be.up, C.unit_low);
        if (t2up.frac != 0 && t2up.exp >= mp_h) { to_slow = true; break; }
        // ---- C = t2 + t1 (mp_add, src/arith/add.cuh:126-184) ----
        p = add_esi(C, t2lo, t2up, t1lo, t1up, Cm.exp[ic] + be.exp, s_exp + al.exp, Cm.sign[ic] ^ be.sign, s_sign ^ al.sign);
        if (p.gamma > log2M || p.theta > log2M) { to_slow = true; break; }   // shift outside the power table
        sign = p.lo.frac < 0;
        if (sign != (p.up.frac < 0)) { to_slow = true; break; }              // sign needs the mixed-radix comparison
        rlo = p.lo; rup = p.up;
        if (sign) { rlo.frac = -p.up.frac; rlo.exp = p.up.exp; rup.frac = -p.lo.frac; rup.exp = p.lo.exp; }
        if (rup.frac != 0 && rup.exp >= mp_h) { to_slow = true; break; }    // result needs a rounding
        go = true;
    } while (0);
    asm volatile("cp.async.wait_all;\n" ::: "memory");
    __syncthreads();                                                            // the staged digits of C are in place
    if (go) {
        // ---- pass 2 over the moduli: digits ----
        const int4 *ty = (const int4 *) (scal_tab + (long long) (p.theta - d + log2M) * NQ);          // alpha 2^(theta - d)
        const int4 *tx = (const int4 *) (scal_tab + (long long) (2 * log2M + 1 + p.gamma) * NQ);      // beta 2^gamma
        int *mycd = cds + threadIdx.x * CP;
        const int *sp2 = Sp;
        int4 y4 = make_int4(0, 0, 0, 0), x4 = make_int4(0, 0, 0, 0);
        const bool negx = (p.sx != 0) != (sign != 0), negy = ((p.sy != 0) != (sign != 0)) != (sg < 0);
#pragma unroll 8
        for (int q = 0; q < NQ; ++q, sp2 += plane) {
            if ((q & 3) == 0) { y4 = __ldg(ty + (q >> 2)); x4 = __ldg(tx + (q >> 2)); }
            const int tyq = (q & 3) == 0 ? y4.x : (q & 3) == 1 ? y4.y : (q & 3) == 2 ? y4.z : y4.w;
            const int txq = (q & 3) == 0 ? x4.x : (q & 3) == 1 ? x4.y : (q & 3) == 2 ? x4.z : x4.w;
            const QConst c = qc[q];
            const int v = SMEM_S ? *sp2 : __ldg(sp2);
            // +-(beta 2^gamma C) +- (alpha 2^(theta-d) S) with every sign (of S, of the two products, of the result) folded into two flags:
            // negation commutes with the modular products, so the digits are those of the reference's sequence of operations
            const unsigned ay = p.nzy ? (unsigned) mulmod_q<F32>(v, tyq, c, kb) : 0u;
            const unsigned ax = p.nzx ? (unsigned) mulmod_q<F32>(mycd[q], txq, c, kb) : 0u;
            const unsigned um = (unsigned) c.m;
            unsigned r = (negx ? um - ax : ax) + (negy ? um - ay : ay);     // in [0, 2 m]
            r = min(r, r - um);
            r = min(r, r - um);                                              // r == m -> 0
            mycd[q] = (int) r;
        }
        okf[threadIdx.x] = 1;
        Cm.sign[ic] = sign;
        Cm.exp[ic] = (p.ex == 0) ? p.ey : p.ex;
        Cm.eval[ic] = rlo;
        Cm.eval[ic + Cm.len()] = rup;
    }
    if (SMEM_S && Sg && to_slow && live) {
#pragma unroll 8
        for (int q = 0; q < NQ; ++q) Sg[q * gplane] = Sp[q * plane];
    }
    __syncthreads();
    for (int v = threadIdx.x; v < rows_live * (NQ / 4); v += kNormFastThreads) {
        const int ent = (4 * v) / NQ;
        if (okf[ent]) {
            const int *src = cds + ent * CP + (4 * v) % NQ;
            cd4[v] = make_int4(src[0], src[1], src[2], src[3]);
        }
    }
    // warp-aggregated appends
    const unsigned lane = threadIdx.x & 31u;
    unsigned bal = __ballot_sync(0xffffffffu, to_todo);
    if (bal) {
        int base = 0;
        if (lane == 0) base = atomicAdd(todo_count, __popc(bal));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (to_todo && fallback_allowed) todo[base + __popc(bal & ((1u << lane) - 1u))] = row + (long long) col * m;
    }
    bal = __ballot_sync(0xffffffffu, to_slow);
    if (bal) {
        int base = 0;
        if (lane == 0) base = atomicAdd(slow_count, __popc(bal));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (to_slow) slow[base + __popc(bal & ((1u << lane) - 1u))] = row + (long long) col * m;
    }
}

#ifndef MPRES_NORM_BLOCKS
#define MPRES_NORM_BLOCKS 6
#endif
// STAGED: the block's residues of S (NQ runs of kNormFastThreads consecutive ints) are brought into shared memory with 16-byte
// cp.async copies up front -- one exposed DRAM latency per block (hidden by the other resident blocks) instead of one per batch of
// loads in each of the two passes, and the second pass reads them from shared memory.
template <int NQ, bool F32, bool STAGED>
__global__ void __launch_bounds__(kNormFastThreads, MPRES_NORM_BLOCKS) k_norm_fast(const DevConsts *Cp, int m, int n, int k, const int *S, const int16_t *delta,
                                                                long long m_p, long long n_p, const OuterInfo *ia, const OuterInfo *ib,
                                                                SoA alpha, SoA beta, SoA Cm, int ldc, const int *scal_tab, long long *todo, int *todo_count,
                                                                long long *slow, int *slow_count, bool fallback_allowed, const int *gate) {
    extern __shared__ __align__(16) int cds[];   // [kNormFastThreads][NQ + 1], then (STAGED) [NQ][kNormFastThreads]
    if (gate && *gate > 0) return;          // the fused small-modulus kernel handled this call
    const int tiles = (m + kNormFastThreads - 1) / kNormFastThreads;
    const int col = blockIdx.x / tiles;
    const int row0 = (blockIdx.x - col * tiles) * kNormFastThreads;
    if (STAGED) {
        int *sst = cds + kNormFastThreads * (NQ + 1);
        const int *src = S + (long long) col * m_p + row0;             // 512-byte aligned: m_p and row0 are multiples of kNormFastThreads
        const long long plane = n_p * m_p;
        constexpr int CH = kNormFastThreads / 4;                         // 16-byte chunks per modulus
        for (int v = threadIdx.x; v < NQ * CH; v += kNormFastThreads) {
            const int q = v / CH, ch = v - q * CH;
            const unsigned sa = (unsigned) __cvta_generic_to_shared(sst + q * kNormFastThreads + 4 * ch);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(src + q * plane + 4 * ch));
        }
        asm volatile("cp.async.commit_group;\n" ::);
        norm_fast_body<NQ, F32, true>(*Cp, cds, m, n, k, col, row0, sst + threadIdx.x, kNormFastThreads, nullptr, 0, delta, m_p, ia, ib,
                                      alpha, beta, Cm, ldc, scal_tab, todo, todo_count, slow, slow_count, fallback_allowed);
        return;
    }
    norm_fast_body<NQ, F32, false>(*Cp, cds, m, n, k, col, row0, S + (long long) col * m_p + row0 + threadIdx.x, n_p * m_p, nullptr, 0, delta, m_p, ia, ib,
                                   alpha, beta, Cm, ldc, scal_tab, todo, todo_count, slow, slow_count, fallback_allowed);
}

}  // namespace mpres
