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
// Stage 2 produced S mod m_q for the first np moduli only.  |S| < M'/4 with M' = m_0 ... m_{np-1}, so the
// mixed-radix digits a_i of X = S mod M' (X = sum a_i m_0...m_{i-1}, src/rns.cuh:570-582 is the reference's
// mixed-radix conversion) determine S: the top digit tells the sign (X/M' lies in [0,1/4) or (3/4,1)) and
//      S mod m_q = sum_i a_i (m_0...m_{i-1} mod m_q) - [S < 0] (M' mod m_q)          for q >= np.
// One thread per entry; digits live in shared memory ([i][thread], conflict-free), the triangular inverse
// table and the weight table are staged per block.  Writes the planes q >= np of S.
constexpr int kExtThreads = 128;
inline size_t base_extend_smem(int N) { return (size_t) N * kExtThreads * sizeof(int) + (size_t) 2 * N * N * sizeof(int) + (size_t) N * 16; }

__global__ void __launch_bounds__(kExtThreads) k_base_extend(const DevConsts *Cp, int m, int n, int *S, long long m_p, long long n_p, const int *nprime) {
    extern __shared__ __align__(16) unsigned char ext_smem[];
    const DevConsts &C = *Cp;
    const int N = C.N;
    const int np = *nprime;
    if (np >= N) return;
    unsigned long long *s_mu = (unsigned long long *) ext_smem;      // [N]
    int *s_m = (int *) (s_mu + N);                                     // [N]
    int *s_inv = s_m + N;                                              // [np][np]   m_i^-1 mod m_j
    int *s_w = s_inv + np * np;                                        // [np + 1][N - np]  prefix weights
    int *xs = s_w + (np + 1) * (N - np);                               // [np][kExtThreads]
    for (int t = threadIdx.x; t < N; t += kExtThreads) { s_mu[t] = C.barrett[t]; s_m[t] = C.moduli[t]; }
    for (int t = threadIdx.x; t < np * np; t += kExtThreads) s_inv[t] = C.mrc_inv[(t / np) * N + t % np];
    for (int t = threadIdx.x; t < (np + 1) * (N - np); t += kExtThreads) s_w[t] = C.prefix_mod[(t / (N - np)) * N + np + t % (N - np)];
    __syncthreads();
    const int tiles = (m + kExtThreads - 1) / kExtThreads;
    const int col = blockIdx.x / tiles;
    const int row = (blockIdx.x - col * tiles) * kExtThreads + threadIdx.x;
    if (row >= m) return;
    int *Sp = S + (long long) col * m_p + row;
    const long long plane = n_p * m_p;
    int *xt = xs + threadIdx.x;
    for (int i = 0; i < np; ++i) xt[i * kExtThreads] = Sp[i * plane];
    // mixed-radix conversion: after step i every element j > i holds (x_j - a_i) / m_i mod m_j
    for (int i = 0; i < np - 1; ++i) {
        const int ai = xt[i * kExtThreads];
        for (int j = i + 1; j < np; ++j) {
            const int mj = s_m[j];
            int a = ai >= mj ? ai - mj : ai;       // moduli of one set differ by less than a factor two ...
            if (a >= mj) a %= mj;                  // ... but stay correct for any set
            int t = xt[j * kExtThreads] - a;
            t = t < 0 ? t + mj : t;
            xt[j * kExtThreads] = mulmod(t, s_inv[i * np + j], mj, s_mu[j]);
        }
    }
    const int top = xt[(np - 1) * kExtThreads];
    const bool neg = 2ll * top >= (long long) s_m[np - 1];
    for (int q = np; q < N; ++q) {
        const int mq = s_m[q];
        const unsigned long long muq = s_mu[q];
        unsigned long long acc = 0;
        for (int i = 0; i < np; ++i) {
            acc += (unsigned long long) (unsigned) xt[i * kExtThreads] * (unsigned) s_w[i * (N - np) + q - np];
            if ((i & 3) == 3) acc = (unsigned long long) (unsigned) reduce64(acc, mq, muq);   // any moduli < 2^31: 4 terms < 2^64
        }
        int r = reduce64(acc, mq, muq);
        if (neg) { r -= s_w[np * (N - np) + q - np]; r = r < 0 ? r + mq : r; }
        Sp[q * plane] = r;
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
    int m, w;
    unsigned long long mu;
    double rrd, rru;
    int al, be;          // digits of alpha and beta
    int pad[2];
};
struct ScalarEsi { int sign, exp; Er lo, up; };

__host__ __device__ constexpr int pow2ceil_c(int n) { int p = 1; while (p < n) p <<= 1; return p; }
__host__ __device__ constexpr int log2_c(int p) { int l = 0; while ((1 << l) < p) ++l; return l; }
__host__ __device__ constexpr int trailing_ones_c(int q) { int t = 0; while (q & 1) { ++t; q >>= 1; } return t; }

constexpr int kNormFastThreads = 128;

template <int NQ>
__global__ void __launch_bounds__(kNormFastThreads) k_norm_fast(const DevConsts *Cp, int m, int n, int k, const int *S, const int16_t *delta,
                                                                long long m_p, long long n_p, const OuterInfo *ia, const OuterInfo *ib,
                                                                SoA alpha, SoA beta, SoA Cm, int ldc, long long *todo, int *todo_count,
                                                                long long *slow, int *slow_count, bool fallback_allowed) {
    constexpr int P = pow2ceil_c(NQ), LOGP = log2_c(P);
    constexpr int CP = NQ + 1;              // pitch of the staged C digits (conflict-free per-thread rows)
    extern __shared__ int cds[];            // [kNormFastThreads][CP]: digits of C in, digits of the result out
    __shared__ QConst qc[NQ];
    __shared__ ScalarEsi s_al, s_be;
    __shared__ unsigned char okf[kNormFastThreads];
    const DevConsts &C = *Cp;
    for (int q = threadIdx.x; q < NQ; q += blockDim.x) {
        QConst c;
        c.m = C.moduli[q]; c.w = C.part_inverse[q]; c.mu = C.barrett[q]; c.rrd = C.recip_rd[q]; c.rru = C.recip_ru[q];
        c.al = alpha.digits[q]; c.be = beta.digits[q]; c.pad[0] = c.pad[1] = 0;
        qc[q] = c;
    }
    if (threadIdx.x == 0) { s_al.sign = alpha.sign[0]; s_al.exp = alpha.exp[0]; s_al.lo = alpha.eval[0]; s_al.up = alpha.eval[alpha.len()]; }
    if (threadIdx.x == 32) { s_be.sign = beta.sign[0]; s_be.exp = beta.exp[0]; s_be.lo = beta.eval[0]; s_be.up = beta.eval[beta.len()]; }
    const int tiles = (m + kNormFastThreads - 1) / kNormFastThreads;
    const int col = blockIdx.x / tiles;
    const int row0 = (blockIdx.x - col * tiles) * kNormFastThreads;
    const int row = row0 + threadIdx.x;
    const bool live = row < m;
    // the digits of this block's C entries are one contiguous run: stage them with coalesced 128-bit loads
    const int rows_live = min(kNormFastThreads, m - row0);
    int4 *cd4 = (int4 *) (Cm.digits + (row0 + (long long) col * ldc) * NQ);
    for (int v = threadIdx.x; v < rows_live * (NQ / 4); v += kNormFastThreads) {
        const int4 t = cd4[v];
        int *dst = cds + ((4 * v) / NQ) * CP + (4 * v) % NQ;
        dst[0] = t.x; dst[1] = t.y; dst[2] = t.z; dst[3] = t.w;
    }
    okf[threadIdx.x] = 0;
    __syncthreads();
    bool to_slow = false, to_todo = false;
    if (live) do {
        const int log2M = C.log2M, mp_h = C.mp_h;
        const OuterInfo ra = ia[row], cb = ib[col];
        if (ra.win < 0 || cb.win < 0) { to_slow = true; break; }             // a line of exact zeros: S == 0
        const long long bound = (long long) ra.win + cb.win + ceil_log2(k);
        if (bound > (long long) log2M - 2) { to_todo = true; to_slow = !fallback_allowed; break; }   // window guard failed
        const int d = delta[(long long) col * m_p + row];
        if (d >= kShiftSentinel) { to_slow = true; break; }                   // every term is an exact zero
        int K = log2M - (int) bound - 3;
        K = K < 0 ? 0 : K;
        // ---- pass 1 over the moduli: residues, magnified fractions, directed sums (balanced tree) ----
        int x[NQ];
        int nz = 0;
        double stl[LOGP + 1], stu[LOGP + 1];
        {
            const int *Sp = S + (long long) col * m_p + row;
            const long long plane = n_p * m_p;
            const int *p2 = C.pow2 + (long long) K * NQ;
#pragma unroll
            for (int q = 0; q < P; ++q) {
                double vl = 0.0, vu = 0.0;
                if (q < NQ) {
                    const QConst c = qc[q];
                    x[q] = __ldg(Sp + q * plane);
                    nz |= x[q];
                    const int sq = mulmod(mulmod(x[q], c.w, c.m, c.mu), __ldg(p2 + q), c.m, c.mu);
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
        if (nz == 0) { to_slow = true; break; }
        const double suml = stl[LOGP], sumu = stu[LOGP];
        const double wl = floor(suml), wu = floor(sumu);
        if (wl != wu) { to_slow = true; break; }
        const double dl = __dsub_rd(suml, wl), du = __dsub_ru(sumu, wu);
        int sg;
        Er lo, up;
        if (du < 0.25 && dl >= C.accuracy) { sg = 1; lo = er_from_double(dl); up = er_from_double(du); }
        else {
            const double ml = __dsub_rd(1.0, du), mh = __dsub_ru(1.0, dl);
            if (dl > 0.75 && ml >= C.accuracy) { sg = -1; lo = er_from_double(ml); up = er_from_double(mh); }
            else { to_slow = true; break; }                                   // needs further magnification rounds
        }
        lo.exp -= K + d; up.exp -= K + d;
        if (up.exp >= mp_h) { to_slow = true; break; }                        // S itself needs a rounding
        const int s_sign = sg < 0 ? 1 : 0, s_exp = ra.emin + cb.emin + d;
        // ---- t1 = alpha * S, t2 = beta * C (mp_mul, src/arith/mul.cuh:53-61), exponent/sign/interval part ----
        const ScalarEsi al = s_al, be = s_be;
        const Er t1lo = er_md_dir<false>(lo, al.lo, C.unit_upp), t1up = er_md_dir<true>(up, al.up, C.unit_low);
        if (t1up.frac != 0 && t1up.exp >= mp_h) { to_slow = true; break; }
        const long long ic = row + (long long) col * ldc;
        const Er clo = Cm.eval[ic], cup = Cm.eval[ic + Cm.len()];
        const Er t2lo = er_md_dir<false>(clo, be.lo, C.unit_upp), t2up = er_md_dir<true>(cup, be.up, C.unit_low);
        if (t2up.frac != 0 && t2up.exp >= mp_h) { to_slow = true; break; }
        // ---- C = t2 + t1 (mp_add, src/arith/add.cuh:126-184) ----
        const AddEsi p = add_esi(C, t2lo, t2up, t1lo, t1up, Cm.exp[ic] + be.exp, s_exp + al.exp, Cm.sign[ic] ^ be.sign, s_sign ^ al.sign);
        if (p.gamma > log2M || p.theta > log2M) { to_slow = true; break; }   // shift outside the power table
        const int sign = p.lo.frac < 0;
        if (sign != (p.up.frac < 0)) { to_slow = true; break; }              // sign needs the mixed-radix comparison
        Er rlo = p.lo, rup = p.up;
        if (sign) { rlo.frac = -p.up.frac; rlo.exp = p.up.exp; rup.frac = -p.lo.frac; rup.exp = p.lo.exp; }
        if (rup.frac != 0 && rup.exp >= mp_h) { to_slow = true; break; }    // result needs a rounding
        // ---- pass 2 over the moduli: digits ----
        const int ysh = p.theta - d;    // alpha*S carries 2^(theta - d)
        const int *ty = ysh >= 0 ? C.pow2 + (long long) ysh * NQ : C.inv_pow2 + (long long) (-ysh) * NQ;
        const int *tx = C.pow2 + (long long) p.gamma * NQ;
        int *mycd = cds + threadIdx.x * CP;
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const QConst c = qc[q];
            int v = x[q];
            if (sg < 0 && v) v = c.m - v;
            const int ay = p.nzy ? mulmod(mulmod(v, c.al, c.m, c.mu), __ldg(ty + q), c.m, c.mu) : 0;
            const int ax = p.nzx ? mulmod(mulmod(mycd[q], c.be, c.m, c.mu), __ldg(tx + q), c.m, c.mu) : 0;
            const int a = p.sx ? (ax ? c.m - ax : 0) : ax;
            const int b = p.sy ? (ay ? c.m - ay : 0) : ay;
            int r = a + b - c.m;
            r = r < 0 ? r + c.m : r;
            mycd[q] = sign ? (r ? c.m - r : 0) : r;
        }
        okf[threadIdx.x] = 1;
        Cm.sign[ic] = sign;
        Cm.exp[ic] = (p.ex == 0) ? p.ey : p.ex;
        Cm.eval[ic] = rlo;
        Cm.eval[ic + Cm.len()] = rup;
    } while (0);
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

}  // namespace mpres
