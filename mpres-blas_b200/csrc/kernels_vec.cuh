// kernels_vec.cuh -- exact-window fast path of mp_gemv and mp_dot: ONE pass over A (or x and y) at HBM speed.
//
// The reference materialises the element-wise products in an m x n (or n) buffer and re-reads it for the
// row / column sums (src/blas/gemv.cuh:199-218, src/blas/dot.cuh:96-105): at least four passes over the
// multiple-precision data.  Here every output is accumulated EXACTLY in residues while the operands stream
// through shared memory once:
//
//      S = sum_l +-X_l Y_l 2^(e_l - base)  mod m_q,        e_l = exp(x_l) + exp(y_l),  base = min_l e_l
//
// which is what the reference's mp_mul / mp_add chain produces (exp = min of the term exponents,
// src/arith/add.cuh:172) whenever none of its steps rounds or drops a term -- independent of the summation
// order, so a single streaming pass with any blocking gives the same digits, sign and exponent.  `base` is
// not known in advance: every accumulator carries a running base and is rescaled (one modular
// multiplication by 2^(old - new)) when a term with a smaller exponent arrives, an O(log n) event.
//
//   k_mv_acc_n    y = A x     rows of A in a block, threads = (row, group of four moduli); the reduction runs over
//                             columns.  The scaled vector entry is a per-column constant, so the residue product
//                             uses Shoup's precomputed-quotient multiplication (one IMAD.HI + two IMAD).
//   k_mv_acc_t    y = A^T x and x . y     the reduction runs along contiguous memory; thread-private accumulators,
//                             block-level combine.
//   k_mv_finalize             combines the partial sums of the splits, checks the magnitude window
//                             (|S| < M/4), resolves sign / interval evaluation by magnification (sign_eval_window),
//                             rounds once if needed and applies y += S (GEMV) or r = S (DOT).  Outputs whose window
//                             guard fails go to a todo list and are recomputed in reference order.
//
// Operands are staged with cp.async (16-byte chunks of the digit rows, the exponent / sign / upper-bound
// fields of the same elements) through a multi-stage shared-memory ring; a thread always consumes the digit
// chunks it copied itself.  Algorithmic traffic: (4N + 24) bytes per element (the lower interval bounds are
// never read).
#pragma once

namespace mpres {

constexpr int kVecNone = INT_MAX;      // running base of an accumulator that has seen no non-zero term
constexpr int kVecSent = INT_MAX;      // term code of an exact zero

__device__ __forceinline__ void cp_async4(void *smem, const void *gmem) {
    unsigned sa = (unsigned) __cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async16_ca(void *smem, const void *gmem) {
    unsigned sa = (unsigned) __cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}

// floor(x 2^32 / m) for x < m (Shoup's precomputed quotient); mu = floor((2^64-1)/m)
__device__ __forceinline__ unsigned shoup_pre(unsigned x, unsigned m, unsigned long long mu) {
    const unsigned long long p = (unsigned long long) x << 32;
    unsigned long long q = __umul64hi(p, mu);
    const unsigned long long r = p - q * m;   // < 2m
    if (r >= m) ++q;
    return (unsigned) q;
}
// d * x mod m with xs = shoup_pre(x): canonical result, d < 2^32, m < 2^31
__device__ __forceinline__ unsigned shoup_mul(unsigned d, unsigned x, unsigned xs, unsigned m) {
    const unsigned q = __umulhi(d, xs);
    const unsigned r = d * x - q * m;   // in [0, 2m), exact modulo 2^32
    return r >= m ? r - m : r;
}
// the same without the final conditional subtraction: result in [0, 2m), congruent to d * x
__device__ __forceinline__ unsigned shoup_mul_lazy(unsigned d, unsigned x, unsigned xs, unsigned m) {
    return d * x - __umulhi(d, xs) * m;
}
// 32-bit Barrett for m < 2^30: k = bit length of m, mu = floor(2^(2k) / m)
struct Barrett32 { unsigned m, mu; int s1, s2; };
__device__ __forceinline__ Barrett32 barrett32_init(unsigned m) {
    Barrett32 b;
    const int k = 32 - __clz(m);
    b.m = m; b.mu = (unsigned) ((1ull << (2 * k)) / m); b.s1 = k - 1; b.s2 = k + 1;
    return b;
}
__device__ __forceinline__ unsigned barrett32_mul(unsigned a, unsigned x, const Barrett32 &b) {   // a, x < m
    const unsigned long long p = (unsigned long long) a * x;
    const unsigned ph = (unsigned) (p >> b.s1);
    const unsigned q = (unsigned) (((unsigned long long) ph * b.mu) >> b.s2);
    unsigned r = (unsigned) p - q * b.m;   // < 3m
    r = r >= b.m ? r - b.m : r;
    return r >= b.m ? r - b.m : r;
}

__device__ __forceinline__ unsigned barrett32_mul_lazy(unsigned a, unsigned x, const Barrett32 &b) {   // result in [0, 3m)
    const unsigned long long p = (unsigned long long) a * x;
    const unsigned ph = (unsigned) (p >> b.s1);
    const unsigned q = (unsigned) (((unsigned long long) ph * b.mu) >> b.s2);
    return (unsigned) p - q * b.m;
}

// One accumulator per residue.  SMALL (every modulus < 2^27): 64-bit lazy sum of (term * 2^s) products, each
// < 2^54, reduced at least every 512 terms.  Otherwise a canonical 32-bit residue.
template <bool SMALL> struct VecAcc { typedef unsigned type; };
template <> struct VecAcc<true> { typedef unsigned long long type; };

template <bool SMALL>
__device__ __forceinline__ void vacc_add(typename VecAcc<SMALL>::type &acc, unsigned t, int neg, unsigned pw, unsigned m, unsigned long long mu) {
    if (SMALL) {
        const unsigned v = neg ? m - t : t;   // t in [0, m)  ->  v in (0, m], congruent to -t
        acc += (unsigned long long) v * pw;
    } else {
        unsigned v = (unsigned) mulmod((int) t, (int) pw, (int) m, mu);
        if (neg && v) v = m - v;
        const unsigned s = (unsigned) acc + v;   // < 2m < 2^32
        acc = s >= m ? s - m : s;
    }
}
template <bool SMALL>
__device__ __forceinline__ unsigned vacc_canon(typename VecAcc<SMALL>::type acc, unsigned m, unsigned long long mu) {
    if (SMALL) return (unsigned) reduce64((unsigned long long) acc, (int) m, mu);
    return (unsigned) acc;
}

// shared-memory bytes of one stage of k_mv_acc_n (16-byte fields first)
__host__ __device__ inline size_t mv_n_stage_bytes(int N, int RB, int T, int CB) {
    size_t off = 0;
    off += (size_t) CB * T * 16;            // digits
    off += (size_t) CB * (RB + 1) * 16;     // upper bounds of A
    off += (size_t) CB * 16;                // upper bounds of the vector entries
    off += (size_t) CB * N * 4;             // vector digits
    off += (size_t) CB * N * 4;             // their Shoup quotients
    off += (size_t) CB * (RB + 4) * 4 * 2;  // exponents, signs of A
    off += (size_t) CB * 4 * 2;             // exponent, sign of the vector entries
    return (off + 15) & ~(size_t) 15;
}
__host__ __device__ inline size_t mv_n_common_bytes(int RB, int CB) { return (size_t) (RB * (CB + 1) + 2 * RB) * 4; }

// An accumulator is labelled with an exponent `lab` <= every term exponent seen so far: value = acc * 2^lab.  The
// label is set kVecSlack below the running minimum, so that only a minimum more than kVecSlack below the
// previous one forces a rescale (one modular multiplication of the accumulator).
constexpr int kVecSlack = 24;
constexpr int kVecExpLimit = 1 << 28;   // exponents beyond this magnitude make the output fail the window guard

// ---- y = A x: partial sums of rows [i0, i0 + RB) over the columns of one split -------------------------------
// v = round(alpha x) (compact, inc 1).  RB = 2^lgRB rows, blockDim = RB * N/4.  Outputs per (split, row):
// pd[(split m + row) N + q] digits, plab label, pmin smallest term exponent, ptop window top.
// SMALL: every modulus has bit length kb <= 27.
// (forcing three blocks per SM through __launch_bounds__ was measured: 6.0 ms against 5.4 ms at config 4 -- the default allocation stays)
// ABS: the signs of both factors are ignored (sums of magnitudes: mp_ge_norm, src/blas/genorm.cuh:39-75).  vs = 0: every column
// uses v[0] (a vector of equal entries without the memory), vs = 1: v[j].
template <bool SMALL, int CB, int STAGES, bool ABS = false>
__global__ void __launch_bounds__(256) k_mv_acc_n(const DevConsts *Cp, SoA A, int lda, int m, int n, SoA v, int lgRB, int cols_per_split,
                                                  int *pd, int *plab, int *pmin, int *ptop, int vs = 1) {
    extern __shared__ __align__(16) unsigned char vsm[];
    typedef typename VecAcc<SMALL>::type acc_t;
    static_assert((CB & (CB - 1)) == 0 && CB >= 2 && CB <= 32, "columns per stage: power of two (phase A reduces over CB lanes)");
    const int N = Cp->N, log2M = Cp->log2M;
    const int *pow2 = Cp->pow2, *spow2 = Cp->spow2;
    const int Q4 = N >> 2, T = blockDim.x, RB = 1 << lgRB;
    const int t = threadIdx.x, r = t / Q4, q4 = t - r * Q4;
    const int i0 = blockIdx.x << lgRB;
    const int rows_valid = min(RB, m - i0);
    const int jbeg = blockIdx.y * cols_per_split, jend = min(n, jbeg + cols_per_split);
    const int nchunks = jend > jbeg ? (jend - jbeg + CB - 1) / CB : 0;
    const int pitchE = RB + 4, pitchU = RB + 1;
    const unsigned stageBytes = (unsigned) mv_n_stage_bytes(N, RB, T, CB);
    const unsigned oUpp = CB * T * 16, oAxu = oUpp + CB * pitchU * 16, oAxd = oAxu + CB * 16, oAxs = oAxd + CB * N * 4, oExp = oAxs + CB * N * 4,
                   oSgn = oExp + CB * pitchE * 4, oAxe = oSgn + CB * pitchE * 4, oAxg = oAxe + CB * 4;
    int *code_s = (int *) (vsm + STAGES * stageBytes);   // [RB][CB + 1]
    int *rowmin_s = code_s + RB * (CB + 1), *rowtop_s = rowmin_s + RB;
    const long long lenA = A.len(), lenv = v.len();
    const int zrow = 2 * (log2M + 1);   // all-zero row of the signed power table

    unsigned mq[4];
    unsigned long long muq[4];
    {
        const int4 mm = *(const int4 *) (Cp->moduli + 4 * q4);
        mq[0] = mm.x; mq[1] = mm.y; mq[2] = mm.z; mq[3] = mm.w;
#pragma unroll
        for (int e = 0; e < 4; ++e) muq[e] = Cp->barrett[4 * q4 + e];
    }
    for (int i = t; i < RB; i += T) { rowmin_s[i] = kVecNone; rowtop_s[i] = INT_MIN; }
    // 16-byte copies of four exponents / signs need aligned rows (i0 and RB are multiples of four)
    const bool quads = (lda & 3) == 0 && ((((size_t) A.exp) | ((size_t) A.sign)) & 15) == 0;
    const bool row_ok = r < rows_valid;
    const long long ldaN = (long long) lda * N;
    const int *gdig = A.digits + (long long) i0 * N + 4 * t;   // + column * ldaN
    const unsigned smem0 = (unsigned) __cvta_generic_to_shared(vsm);

    auto cp16 = [](unsigned dst, const void *src) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src)); };
    auto cp4 = [](unsigned dst, const void *src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(dst), "l"(src)); };

    auto load = [&](int ch, int st) {
        const unsigned S = smem0 + st * stageBytes;
        const int j0 = jbeg + ch * CB;
        const int nc = min(CB, jend - j0);   // valid columns of this chunk
        if (row_ok) {
            const int *g = gdig + (long long) j0 * ldaN;
#pragma unroll
            for (int c = 0; c < CB; ++c, g += ldaN)
                if (c < nc) cp16(S + (c * T + t) * 16, g);
        }
        for (int e = t; e < (CB << lgRB); e += T) {   // upper interval bounds
            const int c = e >> lgRB, rr = e & (RB - 1);
            if (rr < rows_valid && c < nc) cp16(S + oUpp + (c * pitchU + rr) * 16, A.eval + ((long long) (j0 + c) * lda + i0 + rr + lenA));
        }
        for (int e = t; e < (CB << (lgRB - 2)); e += T) {   // exponents and signs, four rows per item
            const int c = e >> (lgRB - 2), r4 = (e & ((RB >> 2) - 1)) << 2;
            if (r4 < rows_valid && c < nc) {
                const long long idx = (long long) (j0 + c) * lda + i0 + r4;
                const unsigned de = S + oExp + (c * pitchE + r4) * 4, ds = S + oSgn + (c * pitchE + r4) * 4;
                if (quads && r4 + 3 < rows_valid) {
                    cp16(de, A.exp + idx);
                    cp16(ds, A.sign + idx);
                } else {
                    for (int k = 0; k < 4 && r4 + k < rows_valid; ++k) { cp4(de + 4 * k, A.exp + idx + k); cp4(ds + 4 * k, A.sign + idx + k); }
                }
            }
        }
        if (r < nc) {   // vector entry j0 + r: thread (r, q4) copies its four digits, q4 == 0 the scalar fields
            const long long jv = (long long) (j0 + r) * vs;
            cp16(S + oAxd + (r * N + 4 * q4) * 4, v.digits + jv * N + 4 * q4);
            if (q4 == 0) {
                cp4(S + oAxe + r * 4, v.exp + jv);
                cp4(S + oAxg + r * 4, v.sign + jv);
                cp16(S + oAxu + r * 16, v.eval + (jv + lenv));
            }
        }
    };

    acc_t acc[4] = {0, 0, 0, 0};
    int mylab = kVecNone;
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nchunks) load(s, s);
        cp_async_commit();
    }
    for (int ch = 0; ch < nchunks; ++ch) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        if (ch + STAGES - 1 < nchunks) load(ch + STAGES - 1, (ch + STAGES - 1) % STAGES);
        cp_async_commit();
        unsigned char *S = vsm + (ch % STAGES) * stageBytes;
        const int nc = min(CB, jend - (jbeg + ch * CB));
        // ---- phase A: Shoup quotients of the vector digits; exponent / sign / window of every term ----
        {
            if (r < nc) {
                const int4 x = *(const int4 *) (S + oAxd + (r * N + 4 * q4) * 4);
                int4 xs;
                xs.x = (int) shoup_pre((unsigned) x.x, mq[0], muq[0]); xs.y = (int) shoup_pre((unsigned) x.y, mq[1], muq[1]);
                xs.z = (int) shoup_pre((unsigned) x.z, mq[2], muq[2]); xs.w = (int) shoup_pre((unsigned) x.w, mq[3], muq[3]);
                *(int4 *) (S + oAxs + (r * N + 4 * q4) * 4) = xs;
            }
            const int *expA = (const int *) (S + oExp), *sgnA = (const int *) (S + oSgn);
            const Er *uppA = (const Er *) (S + oUpp), *axu = (const Er *) (S + oAxu);
            const int *axe = (const int *) (S + oAxe), *axg = (const int *) (S + oAxg);
            for (int e = t; e < (CB << lgRB); e += T) {   // CB * RB and T are multiples of 32: whole warps
                const int c = e & (CB - 1), rr = e / CB;
                int mn = kVecNone, mx = INT_MIN, code = kVecSent;
                if (rr < rows_valid && c < nc) {
                    const Er ua = uppA[c * pitchU + rr], ux = axu[c];
                    if (ua.frac != 0 && ux.frac != 0) {
                        const int ea = expA[c * pitchE + rr], ex = axe[c];
                        const int et = ea + ex;
                        long long tp = (long long) et + ua.exp + ux.exp;
                        tp = tp > (1 << 30) ? (1 << 30) : (tp < -(1 << 30) ? -(1 << 30) : tp);
                        if (abs(ea) > kVecExpLimit || abs(ex) > kVecExpLimit) tp = 1 << 30;   // keeps the 32-bit shift arithmetic exact
                        mn = et; mx = (int) tp;
                        code = et * 2 + (ABS ? 0 : ((sgnA[c * pitchE + rr] ^ axg[c]) & 1));
                    }
                }
                code_s[rr * (CB + 1) + c] = code;
#pragma unroll
                for (int o = 1; o < CB; o <<= 1) {
                    mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
                    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                }
                if (c == 0) { rowmin_s[rr] = min(rowmin_s[rr], mn); rowtop_s[rr] = max(rowtop_s[rr], mx); }
            }
        }
        __syncthreads();
        // ---- phase B: digits ----
        if (row_ok) {
            const int tm = rowmin_s[r];
            if (tm < mylab) {
                const int nl = tm - kVecSlack;
                if (mylab != kVecNone) {
                    const unsigned du = (unsigned) (mylab - nl);
                    const int d = (int) min(du, (unsigned) log2M);   // beyond the table: the row fails the window guard anyway
                    const int4 pw = __ldg((const int4 *) (pow2 + d * N) + q4);
                    const int pv[4] = {pw.x, pw.y, pw.z, pw.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) acc[e] = (acc_t) (unsigned) mulmod((int) vacc_canon<SMALL>(acc[e], mq[e], muq[e]), pv[e], (int) mq[e], muq[e]);
                }
                mylab = nl;
            }
            const int4 *dig = (const int4 *) S + t;
            const int4 *axd4 = (const int4 *) (S + oAxd) + q4, *axs4 = (const int4 *) (S + oAxs) + q4;
            const int *codes = code_s + r * (CB + 1);
#pragma unroll
            for (int c = 0; c < CB; ++c) {
                const int code = codes[c];
                const unsigned su = (unsigned) ((code >> 1) - mylab);
                const int s = (int) min(su, (unsigned) log2M);
                if constexpr (SMALL) {
                    // branch-free: the sign lives in the power table (row 2s + neg), an exact zero selects the zero row;
                    // products r * pw < 2^28 * 2^27 are summed lazily
                    const int row = code == kVecSent ? zrow : 2 * s + (code & 1);
                    const int4 d = dig[c * T], x = axd4[c * Q4], xs = axs4[c * Q4];
                    const int4 pw = __ldg((const int4 *) (spow2 + row * N) + q4);
                    acc[0] += (unsigned long long) shoup_mul_lazy((unsigned) d.x, (unsigned) x.x, (unsigned) xs.x, mq[0]) * (unsigned) pw.x;
                    acc[1] += (unsigned long long) shoup_mul_lazy((unsigned) d.y, (unsigned) x.y, (unsigned) xs.y, mq[1]) * (unsigned) pw.y;
                    acc[2] += (unsigned long long) shoup_mul_lazy((unsigned) d.z, (unsigned) x.z, (unsigned) xs.z, mq[2]) * (unsigned) pw.z;
                    acc[3] += (unsigned long long) shoup_mul_lazy((unsigned) d.w, (unsigned) x.w, (unsigned) xs.w, mq[3]) * (unsigned) pw.w;
                } else if (code != kVecSent) {
                    const int neg = code & 1;
                    const int4 d = dig[c * T], x = axd4[c * Q4], xs = axs4[c * Q4];
                    const int4 pw = __ldg((const int4 *) (pow2 + s * N) + q4);
                    vacc_add<SMALL>(acc[0], shoup_mul((unsigned) d.x, (unsigned) x.x, (unsigned) xs.x, mq[0]), neg, (unsigned) pw.x, mq[0], muq[0]);
                    vacc_add<SMALL>(acc[1], shoup_mul((unsigned) d.y, (unsigned) x.y, (unsigned) xs.y, mq[1]), neg, (unsigned) pw.y, mq[1], muq[1]);
                    vacc_add<SMALL>(acc[2], shoup_mul((unsigned) d.z, (unsigned) x.z, (unsigned) xs.z, mq[2]), neg, (unsigned) pw.z, mq[2], muq[2]);
                    vacc_add<SMALL>(acc[3], shoup_mul((unsigned) d.w, (unsigned) x.w, (unsigned) xs.w, mq[3]), neg, (unsigned) pw.w, mq[3], muq[3]);
                }
            }
            if (SMALL && (ch & 31) == 31) {   // at most 32 CB = 256 products of < 2^55 between reductions
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[e] = (acc_t) vacc_canon<SMALL>(acc[e], mq[e], muq[e]);
            }
        }
    }
    cp_async_wait<0>();
    __syncthreads();
    if (row_ok) {
        const long long o = (long long) blockIdx.y * m + i0 + r;
        int4 out;
        out.x = (int) vacc_canon<SMALL>(acc[0], mq[0], muq[0]); out.y = (int) vacc_canon<SMALL>(acc[1], mq[1], muq[1]);
        out.z = (int) vacc_canon<SMALL>(acc[2], mq[2], muq[2]); out.w = (int) vacc_canon<SMALL>(acc[3], mq[3], muq[3]);
        *((int4 *) (pd + o * N) + q4) = out;
        if (q4 == 0) { plab[o] = mylab; pmin[o] = rowmin_s[r]; ptop[o] = rowtop_s[r]; }
    }
}

// shared-memory bytes of one stage of k_mv_acc_t
__host__ __device__ inline size_t mv_t_stage_bytes(int RS, int T, int RT) {
    size_t off = 0;
    off += (size_t) RT * T * 16 * 2;   // digits of the matrix column and of the vector
    off += (size_t) RS * 16 * 2;       // upper bounds
    off += (size_t) RS * 4 * 4;        // exponents and signs
    return (off + 15) & ~(size_t) 15;
}

// ---- y = A^T x and x . y: one block = one column of M (blockIdx.x; ncols = 1, ldm = 0 for DOT) and the rows of one
//      split (blockIdx.y).  All threads of the block accumulate the same output under one block-wide label.
// pd[(split ncols + col) N + q], plab / pmin / ptop[split ncols + col].  kb: common bit length of the moduli (SMALL).
// ABS / vs as in k_mv_acc_n (mp_asum, column sums of mp_ge_norm: V = one entry "1", vs = 0).
template <bool SMALL, int RT, int STAGES, bool ABS = false>
__global__ void __launch_bounds__(256) k_mv_acc_t(const DevConsts *Cp, SoA M, long long ldm, long long nrows, int ncols, SoA V, int lgRB,
                                                  long long rows_per_split, int kb, int *pd, int *plab, int *pmin, int *ptop, int vs = 1) {
    extern __shared__ __align__(16) unsigned char vsm[];
    typedef typename VecAcc<SMALL>::type acc_t;
    __shared__ int s_min, s_top;
    __shared__ unsigned long long s_sum[kMaxN];
    const int N = Cp->N, log2M = Cp->log2M;
    const int *pow2 = Cp->pow2, *spow2 = Cp->spow2;
    const int Q4 = N >> 2, T = blockDim.x, RB = 1 << lgRB;
    const int t = threadIdx.x, rl = t / Q4, q4 = t - rl * Q4;
    const int RS = RT << lgRB;
    const int col = blockIdx.x, split = blockIdx.y;
    const long long rbeg = (long long) split * rows_per_split, rend = min(nrows, rbeg + rows_per_split);
    const int nchunks = rend > rbeg ? (int) ((rend - rbeg + RS - 1) / RS) : 0;
    const unsigned stageBytes = (unsigned) mv_t_stage_bytes(RS, T, RT);
    const unsigned oDigV = RT * T * 16, oUppM = 2 * oDigV, oUppV = oUppM + RS * 16, oExpM = oUppV + RS * 16, oExpV = oExpM + RS * 4,
                   oSgnM = oExpV + RS * 4, oSgnV = oSgnM + RS * 4;
    int *code_s = (int *) (vsm + STAGES * stageBytes);   // [RS]
    const long long lenM = M.len(), lenV = V.len();
    const long long mbase = (long long) col * ldm;
    const int bs1 = kb - 1, bs2 = kb + 1;

    unsigned mq[4], mu32[4];
    unsigned long long muq[4];
    {
        const int4 mm = *(const int4 *) (Cp->moduli + 4 * q4);
        mq[0] = mm.x; mq[1] = mm.y; mq[2] = mm.z; mq[3] = mm.w;
#pragma unroll
        for (int e = 0; e < 4; ++e) { muq[e] = Cp->barrett[4 * q4 + e]; mu32[e] = SMALL ? (unsigned) ((1ull << (2 * kb)) / mq[e]) : 0u; }
    }
    if (t == 0) { s_min = kVecNone; s_top = INT_MIN; }
    for (int i = t; i < N; i += T) s_sum[i] = 0ull;
    // rbeg and RS are multiples of four: 16-byte copies of four exponents / signs need aligned bases
    const bool quadsM = (mbase & 3) == 0 && ((((size_t) M.exp) | ((size_t) M.sign)) & 15) == 0;
    const bool quadsV = vs != 0 && ((((size_t) V.exp) | ((size_t) V.sign)) & 15) == 0;
    const unsigned smem0 = (unsigned) __cvta_generic_to_shared(vsm);
    const int *gM = M.digits + (mbase + rbeg) * N + 4 * t, *gV = V.digits + (vs ? rbeg * N + 4 * t : 4 * q4);   // + chunk * RS * N + u * 4 T

    auto cp16 = [](unsigned dst, const void *src) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src)); };
    auto cp4 = [](unsigned dst, const void *src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(dst), "l"(src)); };

    // Loader state: the chunks are requested in order, so every source pointer is a running one (this thread's first item of the NEXT chunk to
    // load) advanced by a constant per chunk.  Computing the addresses from the chunk index inside the lambda cost ~45 instructions per pair of
    // copies -- 44 % of the kernel's instructions at 16384 x 16384 / 16 moduli (ncu source view) -- against the ~40 of the arithmetic on that pair.
    // Full chunks whose eight source runs are 16-byte aligned travel as eight bulk copies (cp.async.bulk, the TMA unit) issued by ONE thread and
    // tracked by the stage's mbarrier: every array of a column's rows is one contiguous run, and the per-thread 16-byte copies below spent more
    // instructions on issuing the loads than the arithmetic on using them.  The last (partial) chunk of a split, unaligned views and the
    // single-entry vector of the ABS variants (vs = 0) keep the per-thread copies.
    __shared__ __align__(8) unsigned long long s_bar[STAGES];
    const bool bulk_ok = vs == 1 && quadsM && quadsV &&
                         ((((size_t) M.digits) | ((size_t) V.digits) | ((size_t) M.eval) | ((size_t) V.eval)) & 15) == 0;
    if (t == 0) {
#pragma unroll
        for (int sgi = 0; sgi < STAGES; ++sgi)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned) __cvta_generic_to_shared(&s_bar[sgi])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    auto bulk = [](unsigned dst, const void *src, unsigned bytes, unsigned bar) {
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
    };
    const long long chunkM = (long long) RS * N, chunkV = chunkM * vs;          // ints of digits per chunk
    const int uM = 4 * T, uV = 4 * T * vs;                                      // ints between this thread's rows u and u + 1
    const int *ldM = gM, *ldV = gV;
    const Er *ldEM = M.eval + (mbase + rbeg + lenM) + t, *ldEV = V.eval + (rbeg + t) * vs + lenV;
    const int *ldXM = M.exp + mbase + rbeg + 4 * t, *ldSM = M.sign + mbase + rbeg + 4 * t;
    const int *ldXV = V.exp + (rbeg + 4 * t) * vs, *ldSV = V.sign + (rbeg + 4 * t) * vs;
    long long ld_left = rend - rbeg;                                            // rows not yet requested
    const unsigned sdig = smem0 + t * 16;
    auto load = [&](int st) {
        const unsigned so = st * stageBytes;
        const int nr = (int) min((long long) RS, ld_left);   // valid rows of this chunk
        if (bulk_ok && nr == RS) {
            if (t == 0) {
                const unsigned bar = (unsigned) __cvta_generic_to_shared(&s_bar[st]), S = smem0 + so;
                const unsigned bd = (unsigned) RS * (unsigned) N * 4u, be = (unsigned) RS * 16u, bx = (unsigned) RS * 4u;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // the stage was last touched through the generic proxy
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(2 * bd + 2 * be + 4 * bx) : "memory");
                bulk(S, ldM - 4 * t, bd, bar);
                bulk(S + oDigV, ldV - 4 * t, bd, bar);
                bulk(S + oUppM, ldEM - t, be, bar);
                bulk(S + oUppV, ldEV - t, be, bar);
                bulk(S + oExpM, ldXM - 4 * t, bx, bar);
                bulk(S + oSgnM, ldSM - 4 * t, bx, bar);
                bulk(S + oExpV, ldXV - 4 * t, bx, bar);
                bulk(S + oSgnV, ldSV - 4 * t, bx, bar);
            }
            ldM += chunkM; ldV += chunkV; ldEM += RS; ldEV += RS * vs; ldXM += RS; ldSM += RS; ldXV += RS * vs; ldSV += RS * vs;
            ld_left -= RS;
            return;
        }
#pragma unroll
        for (int u = 0; u < RT; ++u) {
            if (rl + (u << lgRB) < nr) {                // 16-byte chunk t + T u: row rl + RB u, moduli 4 q4 ..
                cp16(sdig + so + T * u * 16, ldM + u * uM);
                cp16(sdig + so + oDigV + T * u * 16, ldV + u * uV);
            }
        }
        {
            const Er *em = ldEM, *ev = ldEV;
            for (int e = t; e < RS; e += T, em += T, ev += T * vs) {
                if (e < nr) {
                    cp16(smem0 + so + oUppM + e * 16, em);
                    cp16(smem0 + so + oUppV + e * 16, ev);
                }
            }
        }
        {
            const int *xm = ldXM, *sm = ldSM, *xv = ldXV, *sv = ldSV;
            for (int e = t; e < (RS >> 2); e += T, xm += 4 * T, sm += 4 * T, xv += 4 * T * vs, sv += 4 * T * vs) {   // exponents and signs, four rows per item
                const int e4 = 4 * e;
                if (e4 < nr) {
                    const bool full = e4 + 3 < nr;
                    const unsigned S = smem0 + so;
                    if (quadsM && full) {
                        cp16(S + oExpM + e * 16, xm);
                        cp16(S + oSgnM + e * 16, sm);
                    } else {
                        for (int k = 0; k < 4 && e4 + k < nr; ++k) { cp4(S + oExpM + e * 16 + 4 * k, xm + k); cp4(S + oSgnM + e * 16 + 4 * k, sm + k); }
                    }
                    if (quadsV && full) {
                        cp16(S + oExpV + e * 16, xv);
                        cp16(S + oSgnV + e * 16, sv);
                    } else {
                        for (int k = 0; k < 4 && e4 + k < nr; ++k) { cp4(S + oExpV + e * 16 + 4 * k, xv + k * vs); cp4(S + oSgnV + e * 16 + 4 * k, sv + k * vs); }
                    }
                }
            }
        }
        ldM += chunkM; ldV += chunkV; ldEM += RS; ldEV += RS * vs; ldXM += RS; ldSM += RS; ldXV += RS * vs; ldSV += RS * vs;
        ld_left -= RS;
    };

    acc_t acc[4] = {0, 0, 0, 0};
    int mylab = kVecNone, pending = 0;
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nchunks) load(s);
        cp_async_commit();
    }
    for (int ch = 0; ch < nchunks; ++ch) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        if (ch + STAGES - 1 < nchunks) load((ch + STAGES - 1) % STAGES);
        cp_async_commit();
        unsigned char *S = vsm + (ch % STAGES) * stageBytes;
        const int nr = (int) min((long long) RS, rend - (rbeg + (long long) ch * RS));
        if (bulk_ok && nr == RS)          // every earlier chunk of this stage was a bulk one too (only a split's last chunk can be partial)
            ptx::mbar_wait((uint64_t *) &s_bar[ch % STAGES], (unsigned) ((ch / STAGES) & 1));
        {   // phase A: one term per thread and pass; block-wide minimum exponent and window top
            const int *expM = (const int *) (S + oExpM), *sgnM = (const int *) (S + oSgnM), *expV = (const int *) (S + oExpV), *sgnV = (const int *) (S + oSgnV);
            const Er *uppM = (const Er *) (S + oUppM), *uppV = (const Er *) (S + oUppV);
            int mn = kVecNone, mx = INT_MIN;
            for (int e = t; e < RS; e += T) {
                int code = kVecSent;
                if (e < nr) {
                    const Er ua = uppM[e], ux = uppV[e];
                    if (ua.frac != 0 && ux.frac != 0) {
                        const int ea = expM[e], ex = expV[e];
                        const int et = ea + ex;
                        long long tp = (long long) et + ua.exp + ux.exp;
                        tp = tp > (1 << 30) ? (1 << 30) : (tp < -(1 << 30) ? -(1 << 30) : tp);
                        if (abs(ea) > kVecExpLimit || abs(ex) > kVecExpLimit) tp = 1 << 30;
                        mn = min(mn, et); mx = max(mx, (int) tp);
                        code = et * 2 + (ABS ? 0 : ((sgnM[e] ^ sgnV[e]) & 1));
                    }
                }
                code_s[e] = code;
            }
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) {
                mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
                mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            }
            if ((t & 31) == 0 && mn != kVecNone) { atomicMin(&s_min, mn); atomicMax(&s_top, mx); }
        }
        __syncthreads();
        {   // phase B
            const int tm = s_min;   // running minimum over everything this block has seen (block-uniform)
            if (tm < mylab) {
                const int nl = tm - kVecSlack;
                if (mylab != kVecNone) {
                    const unsigned du = (unsigned) (mylab - nl);
                    const int d = (int) min(du, (unsigned) log2M);
                    const int4 pw = __ldg((const int4 *) (pow2 + d * N) + q4);
                    const int pv[4] = {pw.x, pw.y, pw.z, pw.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) acc[e] = (acc_t) (unsigned) mulmod((int) vacc_canon<SMALL>(acc[e], mq[e], muq[e]), pv[e], (int) mq[e], muq[e]);
                    pending = 0;
                }
                mylab = nl;
            }
            const int4 *digM = (const int4 *) S + t, *digV = (const int4 *) (S + oDigV) + t;
            const int zrow = 2 * (log2M + 1);
#pragma unroll
            for (int u = 0; u < RT; ++u) {
                const int code = code_s[rl + (u << lgRB)];
                const unsigned su = (unsigned) ((code >> 1) - mylab);
                const int s = (int) min(su, (unsigned) log2M);
                if constexpr (SMALL) {
                    // sign folded into the power table, an exact zero selects the zero row;
                    // lazy products r * pw < 3 * 2^27 * 2^27 < 2^56
                    const int row = code == kVecSent ? zrow : 2 * s + (code & 1);
                    const int4 a = digM[T * u], b = digV[T * u];
                    const int4 pw = __ldg((const int4 *) (spow2 + row * N) + q4);
                    const unsigned av[4] = {(unsigned) a.x, (unsigned) a.y, (unsigned) a.z, (unsigned) a.w};
                    const unsigned bv[4] = {(unsigned) b.x, (unsigned) b.y, (unsigned) b.z, (unsigned) b.w};
                    const unsigned pv[4] = {(unsigned) pw.x, (unsigned) pw.y, (unsigned) pw.z, (unsigned) pw.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const unsigned long long p = (unsigned long long) av[e] * bv[e];
                        const unsigned q = (unsigned) (((unsigned long long) (unsigned) (p >> bs1) * mu32[e]) >> bs2);
                        acc[e] += (unsigned long long) ((unsigned) p - q * mq[e]) * pv[e];
                    }
                } else if (code != kVecSent) {
                    const int neg = code & 1;
                    const int4 a = digM[T * u], b = digV[T * u];
                    const int4 pw = __ldg((const int4 *) (pow2 + s * N) + q4);
                    const unsigned av[4] = {(unsigned) a.x, (unsigned) a.y, (unsigned) a.z, (unsigned) a.w};
                    const unsigned bv[4] = {(unsigned) b.x, (unsigned) b.y, (unsigned) b.z, (unsigned) b.w};
                    const unsigned pv[4] = {(unsigned) pw.x, (unsigned) pw.y, (unsigned) pw.z, (unsigned) pw.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) vacc_add<SMALL>(acc[e], (unsigned) mulmod((int) av[e], (int) bv[e], (int) mq[e], muq[e]), neg, pv[e], mq[e], muq[e]);
                }
            }
            pending += RT;
            if (SMALL && pending >= 120) {
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[e] = (acc_t) vacc_canon<SMALL>(acc[e], mq[e], muq[e]);
                pending = 0;
            }
        }
    }
    cp_async_wait<0>();
    __syncthreads();
    // ---- sum the accumulators of the block (one common label) ----
    if (mylab != kVecNone) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const unsigned a = vacc_canon<SMALL>(acc[e], mq[e], muq[e]);
            if (a) atomicAdd(&s_sum[4 * q4 + e], (unsigned long long) a);
        }
    }
    __syncthreads();
    const long long o = (long long) split * ncols + col;
    for (int i = t; i < N; i += T) pd[o * N + i] = reduce64(s_sum[i], Cp->moduli[i], Cp->barrett[i]);
    if (t == 0) { plab[o] = mylab; pmin[o] = s_min; ptop[o] = s_top; }
}

// Partial sums of one output: digits pd (value = digits * 2^lab), label lab, smallest term exponent mn, window top.
struct MvParts { int *pd, *lab, *mn, *top; };

// ---- many splits, few outputs (DOT): fold the partial sums of one output with a whole block ---------------------
// Block (o, g) folds the splits [g gs, (g + 1) gs) of output o into partial g of `out` (layout [g][o]); with gs >= nsplit `out` holds
// ONE partial per output.  DOT with ~1200 splits folds in two levels (37 blocks, then one) instead of one block walking all of them.
__global__ void __launch_bounds__(256) k_mv_combine(const DevConsts *Cp, int nout, int nsplit, int gs, MvParts in, MvParts out) {
    __shared__ int s_lab, s_min, s_top;
    __shared__ unsigned long long s_sum[kMaxN];
    const DevConsts &C = *Cp;
    const int N = C.N, t = threadIdx.x, o = blockIdx.x;
    const int p0 = blockIdx.y * gs, p1 = min(nsplit, p0 + gs);
    if (t == 0) { s_lab = kVecNone; s_min = kVecNone; s_top = INT_MIN; }
    for (int i = t; i < N; i += 256) s_sum[i] = 0ull;
    __syncthreads();
    int lb = kVecNone, mn = kVecNone, mx = INT_MIN;
    for (int p = p0 + t; p < p1; p += 256) {
        const long long e = (long long) p * nout + o;
        const int b = in.lab[e];
        if (b != kVecNone) { lb = min(lb, b); mn = min(mn, in.mn[e]); mx = max(mx, in.top[e]); }
    }
    if (lb != kVecNone) { atomicMin(&s_lab, lb); atomicMin(&s_min, mn); atomicMax(&s_top, mx); }
    __syncthreads();
    const int nb = s_lab;
    for (long long it = t; it < (long long) (p1 - p0) * N; it += 256) {
        const int p = p0 + (int) (it / N), q = (int) (it % N);
        const long long e = (long long) p * nout + o;
        const int b = in.lab[e];
        if (b == kVecNone) continue;
        long long dl = (long long) b - nb;
        const int d = dl > C.log2M ? C.log2M : (int) dl;
        const int v = mulmod(in.pd[e * N + q], __ldg(C.pow2 + (long long) d * N + q), C.moduli[q], C.barrett[q]);
        if (v) atomicAdd(&s_sum[q], (unsigned long long) v);
    }
    __syncthreads();
    const long long eo = (long long) blockIdx.y * nout + o;
    for (int i = t; i < N; i += 256) out.pd[eo * N + i] = reduce64(s_sum[i], C.moduli[i], C.barrett[i]);
    if (t == 0) { out.lab[eo] = nb; out.mn[eo] = s_min; out.top[eo] = s_top; }
}

// ---- combine the splits, normalise, apply --------------------------------------------------------------------
// add_to_y: y[o] = round(y[o] + S_o) (y already holds round(beta y), src/blas/gemv.cuh:199-218); otherwise r = S
// (stored to out[0] or to the AoS record rec_out; store_each: to y[o]).  nterms: length of every sum (for the magnitude bound).
template <int G, int R>
__global__ void __launch_bounds__(256) k_mv_finalize(const DevConsts *Cp, int nout, int nsplit, MvParts in, long long nterms, bool add_to_y, SoA y,
                                                     int incy, char *rec_out, int *todo, int *todo_count, bool fallback_allowed, bool store_each = false) {
    const DevConsts &C = *Cp;
    Lane<R> L;
    lane_init<G, R>(C, L);
    const long long o = ((long long) blockIdx.x * blockDim.x + threadIdx.x) / G;
    if (o >= nout) return;
    int nb = kVecNone, tm = kVecNone, top = INT_MIN;
    for (int p = 0; p < nsplit; ++p) {
        const long long e = (long long) p * nout + o;
        const int b = in.lab[e];
        if (b != kVecNone) { nb = min(nb, b); tm = min(tm, in.mn[e]); top = max(top, in.top[e]); }
    }
    Num<R> s;
    num_zero(s);
    if (nb != kVecNone) {
        // Result exponent: the smallest term exponent, or 0 if that is positive (the reference's sums start from
        // MP_ZERO, whose exponent 0 takes part in exp = min(...), src/arith/add.cuh:172).
        const int rexp = tm < 0 ? tm : 0;
        if (rexp < nb) nb = rexp;            // only when every term exponent is positive
        const int d = rexp - nb;             // digits are scaled by 2^d too much: exact division after the sign is known
        int lg = 0;
        while ((1ll << lg) < nterms) ++lg;
        // |X_l Y_l 2^(e_l - nb)| < 2^(2 log2M + 4 + top - nb): X < 2^(log2M + up.exp + 2) for both factors
        long long bound = 2ll * C.log2M + 4 + (long long) top - nb + lg;
        if (bound > (long long) C.log2M - 2) {
            if ((threadIdx.x & (G - 1)) == 0) {
                const int pos = atomicAdd(todo_count, 1);
                if (fallback_allowed) todo[pos] = (int) o;
            }
            if (fallback_allowed) return;
            bound = C.log2M - 2;
        }
        for (int p = 0; p < nsplit; ++p) {
            const long long e = (long long) p * nout + o;
            const int b = in.lab[e];
            if (b == kVecNone) continue;
            long long dl = (long long) b - nb;
            const int sh = dl > C.log2M ? C.log2M : (int) dl;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (!L.act[r]) continue;
                const int x = __ldg(in.pd + e * C.N + L.idx[r]);
                const int v = mulmod(x, __ldg(C.pow2 + (long long) sh * C.N + L.idx[r]), L.m[r], L.mu[r]);
                const int w = s.d[r] + v - L.m[r];
                s.d[r] = w < 0 ? w + L.m[r] : w;
            }
        }
        Er lo, up;
        const int sg = sign_eval_window<G, R>(C, L, s.d, (int) bound, lo, up);
        if (sg != 0) {
            const int dd = d > C.log2M ? C.log2M : d;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                int v = s.d[r];
                if (sg < 0 && v) v = L.m[r] - v;
                s.d[r] = L.act[r] ? mulmod(v, __ldg(C.inv_pow2 + (long long) dd * C.N + L.idx[r]), L.m[r], L.mu[r]) : 0;
            }
            s.sign = sg < 0 ? 1 : 0;
            s.exp = rexp;
            s.lo = lo; s.up = up;
            s.lo.exp -= dd; s.up.exp -= dd;
            round_if_needed<G, R>(C, L, s);
        } else {
            num_zero(s);
        }
    }
    if (add_to_y) {
        const long long iy = inc_index(o, nout, incy);
        Num<R> a;
        load_num<G, R>(C, L, y, iy, a);
        mp_add<G, R, true>(C, L, a, a, s);
        store_num<G, R>(C, L, y, iy, a);
    } else if (rec_out) {
        store_rec<G, R>(C, L, rec_out, 0, s);
    } else {
        store_num<G, R>(C, L, y, store_each ? inc_index(o, nout, incy) : 0, s);
    }
}

}  // namespace mpres

// ---- host launchers ----------------------------------------------------------------------------------------

// rows per block: the largest power of two RB with RB * N/4 <= 256 threads and a whole number of warps
static inline int mv_log2_rows_per_block(int N) {
    const int Q4 = N / 4;
    for (int lg = 8; lg >= 3; --lg)
        if ((Q4 << lg) <= 256 && ((Q4 << lg) % 32) == 0) return lg;
    return -1;
}
// every modulus has the same bit length kb <= 27 (the predefined n-double sets): lazy 64-bit accumulation and
// the 32-bit Barrett step apply.  Returns kb or 0.
static inline int mv_small_moduli_bits(const mpres_ctx *c) {
    int kb = 0;
    for (int i = 0; i < c->hc.N; ++i) {
        int b = 0;
        while ((1ll << b) <= c->hc.moduli[i]) ++b;
        if (kb == 0) kb = b;
        if (b != kb || b > 27 || b < 2) return 0;
    }
    return kb;
}

// partial-sum workspace: [nsplit][nout] partials, one folded partial per output, todo list [nout]
constexpr int kMvMidMax = 64;   // partials of the first folding level (DOT)
struct MvWork { MvParts parts, folded, mid; int *todo; };
static inline int mv_workspace(mpres_ctx *c, long long nout, int nsplit, MvWork *w) {
    const size_t cnt = (size_t) nout * nsplit;
    const size_t nmid = nout == 1 ? kMvMidMax : 0;
    const size_t bytes = (cnt + nout + nmid) * c->hc.N * 4 + 48 + (cnt + nout + nmid) * 12 + (size_t) nout * 4 + 64;
    void *p;
    int rc = ws_reserve(c, 2, bytes, &p);
    if (rc) return rc;
    char *b = (char *) p;
    w->parts.pd = (int *) b; b += (cnt * c->hc.N * 4 + 15) / 16 * 16;
    w->folded.pd = (int *) b; b += ((size_t) nout * c->hc.N * 4 + 15) / 16 * 16;
    w->mid.pd = (int *) b; b += (nmid * c->hc.N * 4 + 15) / 16 * 16;
    w->mid.lab = (int *) b; b += nmid * 4;
    w->mid.mn = (int *) b; b += nmid * 4;
    w->mid.top = (int *) b; b += nmid * 4;
    w->parts.lab = (int *) b; b += cnt * 4;
    w->parts.mn = (int *) b; b += cnt * 4;
    w->parts.top = (int *) b; b += cnt * 4;
    w->folded.lab = (int *) b; b += (size_t) nout * 4;
    w->folded.mn = (int *) b; b += (size_t) nout * 4;
    w->folded.top = (int *) b; b += (size_t) nout * 4;
    w->todo = (int *) b;
    return 0;
}

static inline void mv_mark(mpres_ctx *c, int i, cudaStream_t st) {
    if (c->profiling) { if (!c->ev[i]) cudaEventCreate(&c->ev[i]); cudaEventRecord(c->ev[i], st); }
}

template <bool SMALL, int CB, int STAGES>
static inline void mv_launch_n(mpres_ctx *c, dim3 grid, int T, int lgRB, SoA A, int lda, int m, int n, SoA v, int cols_per, const MvParts &o, cudaStream_t st) {
    const size_t smem = STAGES * mv_n_stage_bytes(c->hc.N, 1 << lgRB, T, CB) + mv_n_common_bytes(1 << lgRB, CB);
    cudaFuncSetAttribute(k_mv_acc_n<SMALL, CB, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    k_mv_acc_n<SMALL, CB, STAGES><<<grid, T, smem, st>>>(c->dconsts, A, lda, m, n, v, lgRB, cols_per, o.pd, o.lab, o.mn, o.top);
}
template <bool SMALL, int RT, int STAGES>
static inline void mv_launch_t(mpres_ctx *c, dim3 grid, int T, int lgRB, SoA M, long long ldm, long long nrows, int ncols, SoA V, long long rows_per,
                               int kb, const MvParts &o, cudaStream_t st) {
    const int RS = RT << lgRB;
    const size_t smem = STAGES * mv_t_stage_bytes(RS, T, RT) + (size_t) RS * 4;
    cudaFuncSetAttribute(k_mv_acc_t<SMALL, RT, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    k_mv_acc_t<SMALL, RT, STAGES><<<grid, T, smem, st>>>(c->dconsts, M, ldm, nrows, ncols, V, lgRB, rows_per, kb, o.pd, o.lab, o.mn, o.top);
}
// sums of magnitudes against the single entry `one`
template <bool SMALL>
static inline void mv_launch_t_abs(mpres_ctx *c, dim3 grid, int T, int lgRB, SoA M, long long ldm, long long nrows, int ncols, SoA one, long long rows_per,
                                   int kb, const MvParts &o, cudaStream_t st) {
    constexpr int RT = 4, STAGES = 2;
    const int RS = RT << lgRB;
    const size_t smem = STAGES * mv_t_stage_bytes(RS, T, RT) + (size_t) RS * 4;
    cudaFuncSetAttribute(k_mv_acc_t<SMALL, RT, STAGES, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    k_mv_acc_t<SMALL, RT, STAGES, true><<<grid, T, smem, st>>>(c->dconsts, M, ldm, nrows, ncols, one, lgRB, rows_per, kb, o.pd, o.lab, o.mn, o.top, 0);
}
template <bool SMALL>
static inline void mv_launch_n_abs(mpres_ctx *c, dim3 grid, int T, int lgRB, SoA A, int lda, int m, int n, SoA one, int cols_per, const MvParts &o, cudaStream_t st) {
    constexpr int CB = 8, STAGES = 2;
    const size_t smem = STAGES * mv_n_stage_bytes(c->hc.N, 1 << lgRB, T, CB) + mv_n_common_bytes(1 << lgRB, CB);
    cudaFuncSetAttribute(k_mv_acc_n<SMALL, CB, STAGES, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    k_mv_acc_n<SMALL, CB, STAGES, true><<<grid, T, smem, st>>>(c->dconsts, A, lda, m, n, one, lgRB, cols_per, o.pd, o.lab, o.mn, o.top, 0);
}

// tile configurations (A/B switch mpres_set_vec_config): columns (N) / row tiles (T) per stage x ring depth
constexpr int kMvRTmax = 4;

// y := y + op(A) v with y = round(beta y), v = round(alpha x) already formed.  *done: all launches issued
// (rows whose window guard fails are recomputed in reference order from the todo list).
inline int gemv_fast(mpres_ctx *c, bool tr, int m, int n, SoA A, int lda, SoA v, SoA y, int incy, cudaStream_t st, bool *done) {
    *done = false;
    const int N = c->hc.N;
    if (N % 4 != 0 || N > kMaxN) return 0;
    const int lgRB = mv_log2_rows_per_block(N);
    if (lgRB < 0) return 0;
    const int RB = 1 << lgRB, Q4 = N / 4, T = RB * Q4;
    const int kb = mv_small_moduli_bits(c);
    const bool allow_fb = c->mode == MPRES_MODE_AUTO;
    const int nout = tr ? n : m;
    const long long nterms = tr ? m : n;
    const int target = c->sm_count * 8;
    const int cfg = c->vec_config;
    MvWork w;
    int rc, nsplit;
    mv_mark(c, 1, st);
    if (!tr) {
        const int CB = cfg == 1 ? 4 : 8;
        const int nrb = (m + RB - 1) / RB;
        nsplit = std::max(1, std::min((target + nrb - 1) / nrb, (n + 8 * CB - 1) / (8 * CB)));
        int cols_per = (n + nsplit - 1) / nsplit;
        cols_per = (cols_per + CB - 1) / CB * CB;
        nsplit = (n + cols_per - 1) / cols_per;
        if ((rc = mv_workspace(c, nout, nsplit, &w))) return rc;
        const dim3 grid((unsigned) nrb, (unsigned) nsplit);
        if (kb) {
            if (cfg == 1) mv_launch_n<true, 4, 3>(c, grid, T, lgRB, A, lda, m, n, v, cols_per, w.parts, st);
            else if (cfg == 2) mv_launch_n<true, 8, 3>(c, grid, T, lgRB, A, lda, m, n, v, cols_per, w.parts, st);
            else mv_launch_n<true, 8, 2>(c, grid, T, lgRB, A, lda, m, n, v, cols_per, w.parts, st);
        } else {
            if (cfg == 1) mv_launch_n<false, 4, 3>(c, grid, T, lgRB, A, lda, m, n, v, cols_per, w.parts, st);
            else mv_launch_n<false, 8, 2>(c, grid, T, lgRB, A, lda, m, n, v, cols_per, w.parts, st);
        }
    } else {
        const int RT = cfg == 1 ? 2 : 4;
        const int RS = RT * RB;
        nsplit = std::max(1, std::min((target + n - 1) / n, (m + 4 * RS - 1) / (4 * RS)));
        long long rows_per = ((long long) m + nsplit - 1) / nsplit;
        rows_per = (rows_per + RS - 1) / RS * RS;
        nsplit = (int) ((m + rows_per - 1) / rows_per);
        if ((rc = mv_workspace(c, nout, nsplit, &w))) return rc;
        const dim3 grid((unsigned) n, (unsigned) nsplit);
        if (kb) {
            if (cfg == 1) mv_launch_t<true, 2, 3>(c, grid, T, lgRB, A, lda, m, n, v, rows_per, kb, w.parts, st);
            else if (cfg == 2) mv_launch_t<true, 4, 3>(c, grid, T, lgRB, A, lda, m, n, v, rows_per, kb, w.parts, st);
            else mv_launch_t<true, 4, 2>(c, grid, T, lgRB, A, lda, m, n, v, rows_per, kb, w.parts, st);
        } else {
            if (cfg == 1) mv_launch_t<false, 2, 3>(c, grid, T, lgRB, A, lda, m, n, v, rows_per, kb, w.parts, st);
            else mv_launch_t<false, 4, 2>(c, grid, T, lgRB, A, lda, m, n, v, rows_per, kb, w.parts, st);
        }
    }
    LAUNCHED(c);
    CUDA_TRY(cudaGetLastError());
    mv_mark(c, 2, st);
    MPRES_DISPATCH(N, {
        const unsigned blocks = (unsigned) (((long long) nout * G + 255) / 256);
        k_mv_finalize<G, R><<<blocks, 256, 0, st>>>(c->dconsts, nout, nsplit, w.parts, nterms, true, y, incy, nullptr, w.todo, c->d_counter, allow_fb);
        LAUNCHED(c);
        if (allow_fb) {
            k_gemv_ref_order<G, R><<<c->sm_count * 4, 128, 0, st>>>(c->dconsts, tr, m, n, A, lda, v, y, incy, w.todo, c->d_counter);
            LAUNCHED(c);
        }
    });
    CUDA_TRY(cudaGetLastError());
    mv_mark(c, 3, st);
    c->ev_valid = c->profiling;
    c->last_stage2_launches = 1;
    *done = true;
    return 0;
}

// r := x . y (unit strides) into the AoS record rec_out or out[0].  d_counter[0] != 0 afterwards means the window
// guard failed and the caller's reference-order kernels (gated on that counter) produce the result instead.
inline int dot_fast(mpres_ctx *c, int n, SoA x, int incx, SoA y, int incy, char *rec_out, SoA out, cudaStream_t st, bool *done, bool *launched) {
    *done = false;
    *launched = false;
    const int N = c->hc.N;
    if (incx != 1 || incy != 1 || N % 4 != 0 || N > kMaxN) return 0;
    const int lgRB = mv_log2_rows_per_block(N);
    if (lgRB < 0) return 0;
    const int cfg = c->vec_config;
    const int RT = cfg == 1 ? 2 : 4;
    const int RB = 1 << lgRB, Q4 = N / 4, T = RB * Q4, RS = RT * RB;
    const int kb = mv_small_moduli_bits(c);
    const bool allow_fb = c->mode == MPRES_MODE_AUTO;
    int nsplit = std::max(1, std::min(c->sm_count * 8, (n + 4 * RS - 1) / (4 * RS)));
    long long rows_per = ((long long) n + nsplit - 1) / nsplit;
    rows_per = (rows_per + RS - 1) / RS * RS;
    nsplit = (int) ((n + rows_per - 1) / rows_per);
    MvWork w;
    int rc;
    if ((rc = mv_workspace(c, 1, nsplit, &w))) return rc;
    mv_mark(c, 0, st);
    mv_mark(c, 1, st);
    const dim3 grid(1, (unsigned) nsplit);
    if (kb) {
        if (cfg == 1) mv_launch_t<true, 2, 3>(c, grid, T, lgRB, x, 0, n, 1, y, rows_per, kb, w.parts, st);
        else if (cfg == 2) mv_launch_t<true, 4, 3>(c, grid, T, lgRB, x, 0, n, 1, y, rows_per, kb, w.parts, st);
        else mv_launch_t<true, 4, 2>(c, grid, T, lgRB, x, 0, n, 1, y, rows_per, kb, w.parts, st);
    } else {
        if (cfg == 1) mv_launch_t<false, 2, 3>(c, grid, T, lgRB, x, 0, n, 1, y, rows_per, kb, w.parts, st);
        else mv_launch_t<false, 4, 2>(c, grid, T, lgRB, x, 0, n, 1, y, rows_per, kb, w.parts, st);
    }
    LAUNCHED(c);
    CUDA_TRY(cudaGetLastError());
    mv_mark(c, 2, st);
    const bool fold = nsplit > 4;
    if (fold) {
        const int gs = 32, groups = (nsplit + gs - 1) / gs;
        if (groups > 1 && groups <= kMvMidMax) {      // two levels: `groups` blocks fold 32 splits each, one block folds their partials
            k_mv_combine<<<dim3(1, (unsigned) groups), 256, 0, st>>>(c->dconsts, 1, nsplit, gs, w.parts, w.mid);
            k_mv_combine<<<1, 256, 0, st>>>(c->dconsts, 1, groups, groups, w.mid, w.folded);
            LAUNCHED(c);
        } else {
            k_mv_combine<<<1, 256, 0, st>>>(c->dconsts, 1, nsplit, nsplit, w.parts, w.folded);
        }
        LAUNCHED(c);
    }
    MPRES_DISPATCH(N, {
        k_mv_finalize<G, R><<<1, 256, 0, st>>>(c->dconsts, 1, fold ? 1 : nsplit, fold ? w.folded : w.parts, (long long) n, false, out, 1, rec_out, w.todo,
                                               c->d_counter, allow_fb);
    });
    LAUNCHED(c);
    CUDA_TRY(cudaGetLastError());
    mv_mark(c, 3, st);
    c->ev_valid = c->profiling;
    c->last_stage2_launches = 1;
    *launched = true;
    *done = !allow_fb;   // AUTO: the caller still launches the reference-order kernels, gated on d_counter[0]
    return 0;
}

// ---- sums of magnitudes (mp_asum, the row / column sums of mp_ge_norm) on the exact-window accumulators ------------------------------
namespace mpres {

// the number 1 as a one-element SoA (digits 1, exponent 0, interval evaluation of 1 / M)
__global__ void k_make_one(const DevConsts *Cp, SoA one) {
    const DevConsts &C = *Cp;
    const int t = threadIdx.x;
    if (t < C.N) one.digits[t] = 1;
    if (t == 0) {
        one.sign[0] = 0; one.exp[0] = 0;
        Er lo = C.unit_low, up = C.unit_upp;
        er_adjust(lo); er_adjust(up);
        one.eval[0] = lo; one.eval[1] = up;
    }
}

// reference-order fallback: out[o] = sum_l |X(o, l)| for the listed outputs (or all of them), one lane group each, sequential mp_add
// with the rounding of src/arith/add.cuh:197-199.  Element (o, l) at o * so + l * sl.
template <int G, int R>
__global__ void k_abs_sum_ref(const DevConsts *Cp, SoA X, long long so, long long sl, int nout, long long nterms, SoA out, int inco,
                              const int *todo, const int *todo_count) {
    const DevConsts &C = *Cp;
    Lane<R> L;
    lane_init<G, R>(C, L);
    long long grp = ((long long) blockIdx.x * blockDim.x + threadIdx.x) / G;
    const long long ngrp = (long long) gridDim.x * blockDim.x / G;
    const long long total = todo ? (long long) *todo_count : (long long) nout;
    for (; grp < total; grp += ngrp) {
        const long long o = todo ? todo[grp] : grp;
        Num<R> sum, a;
        num_zero(sum);
        for (long long l = 0; l < nterms; ++l) {
            load_num<G, R>(C, L, X, o * so + l * sl, a);
            a.sign = 0;
            mp_add<G, R, true>(C, L, sum, sum, a);
        }
        store_num<G, R>(C, L, out, inc_index(o, nout, inco), sum);
    }
}

}  // namespace mpres

// out[o] = sum over l of |X(o, l)|, o < nout, l < nterms; X(o, l) at o * so + l * sl with so == 1 (lines strided: row sums of a column-major
// matrix) or sl == 1 (lines contiguous: column sums, a vector).  One pass over X at HBM speed; outputs whose window guard fails
// (exponent spreads beyond the format) are recomputed in reference order.
inline int abs_sums_fast(mpres_ctx *c, SoA X, long long so, long long sl, int nout, long long nterms, SoA out, cudaStream_t st, bool *done) {
    *done = false;
    const int N = c->hc.N;
    if (N % 4 != 0 || N > kMaxN || nterms > 0x7fffffffll) return 0;
    const int lgRB = mv_log2_rows_per_block(N);
    if (lgRB < 0) return 0;
    const int RB = 1 << lgRB, Q4 = N / 4, T = RB * Q4;
    const int kb = mv_small_moduli_bits(c);
    SoA one;
    int rc;
    if ((rc = ws_soa(c, 19, 1, &one))) return rc;
    k_make_one<<<1, 128, 0, st>>>(c->dconsts, one);
    LAUNCHED(c);
    MvWork w;
    int nsplit;
    const int target = c->sm_count * 8;
    if (sl == 1) {
        constexpr int RT = 4;
        const int RS = RT * RB;
        nsplit = (int) std::max<long long>(1, std::min<long long>((target + nout - 1) / nout, (nterms + 4 * RS - 1) / (4 * RS)));
        long long rows_per = (nterms + nsplit - 1) / nsplit;
        rows_per = (rows_per + RS - 1) / RS * RS;
        nsplit = (int) ((nterms + rows_per - 1) / rows_per);
        if ((rc = mv_workspace(c, nout, nsplit, &w))) return rc;
        const dim3 grid((unsigned) nout, (unsigned) nsplit);
        if (kb) mv_launch_t_abs<true>(c, grid, T, lgRB, X, so, nterms, nout, one, rows_per, kb, w.parts, st);
        else mv_launch_t_abs<false>(c, grid, T, lgRB, X, so, nterms, nout, one, rows_per, kb, w.parts, st);
    } else if (so == 1) {
        constexpr int CB = 8;
        const int nrb = (nout + RB - 1) / RB;
        nsplit = (int) std::max<long long>(1, std::min<long long>((target + nrb - 1) / nrb, (nterms + 8 * CB - 1) / (8 * CB)));
        int cols_per = (int) ((nterms + nsplit - 1) / nsplit);
        cols_per = (cols_per + CB - 1) / CB * CB;
        nsplit = (int) ((nterms + cols_per - 1) / cols_per);
        if ((rc = mv_workspace(c, nout, nsplit, &w))) return rc;
        const dim3 grid((unsigned) nrb, (unsigned) nsplit);
        if (kb) mv_launch_n_abs<true>(c, grid, T, lgRB, X, (int) sl, nout, (int) nterms, one, cols_per, w.parts, st);
        else mv_launch_n_abs<false>(c, grid, T, lgRB, X, (int) sl, nout, (int) nterms, one, cols_per, w.parts, st);
    } else {
        return 0;
    }
    LAUNCHED(c);
    CUDA_TRY(cudaGetLastError());
    const bool allow_fb = c->mode == MPRES_MODE_AUTO;
    MPRES_DISPATCH(N, {
        const unsigned blocks = (unsigned) (((long long) nout * G + 255) / 256);
        k_mv_finalize<G, R><<<blocks, 256, 0, st>>>(c->dconsts, nout, nsplit, w.parts, nterms, false, out, 1, nullptr, w.todo, c->d_counter, allow_fb, true);
        LAUNCHED(c);
        if (allow_fb) {
            k_abs_sum_ref<G, R><<<c->sm_count * 4, 128, 0, st>>>(c->dconsts, X, so, sl, nout, nterms, out, 1, w.todo, c->d_counter);
            LAUNCHED(c);
        }
    });
    CUDA_TRY(cudaGetLastError());
    *done = true;
    return 0;
}
