// kernels_ref_order.cuh -- "reference-order" kernels: the reference's per-step semantics (mp_mul,
// mp_add, rounding after every operation, same summation order where it is fixed by the algorithm),
// executed residue-parallel by lane groups.  They are (a) the bit-exact parity mode against the
// reference kernels and (b) the per-element fallback of the exact-window fast path.
#pragma once

#include "mp_device.cuh"

namespace mpres {

// op(X)(row, col) of a column-major matrix with leading dimension ld
__device__ __forceinline__ long long mat_index(bool trans, long long row, long long col, long long ld) {
    return trans ? col + row * ld : row + col * ld;
}

// ---- element-wise probes (tests) -------------------------------------------------------------------
template <int G, int R>
__global__ void k_probe(const DevConsts *Cp, int op, char *r, const char *x, const char *y, const int *bits, long long n) {
    const DevConsts &C = *Cp;
    Lane<R> L;
    lane_init<G, R>(C, L);
    const long long grp = ((long long) blockIdx.x * blockDim.x + threadIdx.x) / G;
    if (grp >= n) return;
    Num<R> a, b, out;
    load_rec<G, R>(C, L, x, grp, a);
    if (op == 0) { load_rec<G, R>(C, L, y, grp, b); mp_mul<G, R, true>(C, L, out, a, b); }
    else if (op == 1) { load_rec<G, R>(C, L, y, grp, b); mp_add<G, R, true>(C, L, out, a, b); }
    else if (op == 2) { out = a; eval_compute<G, R, false>(C, L, out.d, out.lo, out.up); }
    else if (op == 3) { out = a; eval_compute<G, R, true>(C, L, out.d, out.lo, out.up); }
    else { out = a; mp_round<G, R>(C, L, out, bits[grp]); }
    store_rec<G, R>(C, L, r, grp, out);
}

// ---- GEMM stage: S = op(A) * op(B), reference order (src/blas/gemm.cuh:39-58) ----------------------
// One lane group per element of S; groups are laid out along rows so that neighbouring groups read
// neighbouring elements of a column of A.  `todo` (optional) lists the elements to compute
// (fallback of the fast path): todo[t] = row + col * m.
template <int G, int R>
__global__ void k_gemm_ref_order(const DevConsts *Cp, bool ta, bool tb, int m, int n, int k, SoA A, int lda, SoA B, int ldb,
                                 SoA S, int lds, const long long *todo, const int *todo_count) {
    const DevConsts &C = *Cp;
    Lane<R> L;
    lane_init<G, R>(C, L);
    long long grp = ((long long) blockIdx.x * blockDim.x + threadIdx.x) / G;
    const long long ngrp = (long long) gridDim.x * blockDim.x / G;
    const long long total = todo ? (long long) *todo_count : (long long) m * n;
    for (; grp < total; grp += ngrp) {
        const long long e = todo ? todo[grp] : grp;
        const int row = (int) (e % m), col = (int) (e / m);
        Num<R> sum, prod, a, b;
        num_zero(sum);
        for (int l = 0; l < k; ++l) {
            load_num<G, R>(C, L, A, mat_index(ta, row, l, lda), a);
            load_num<G, R>(C, L, B, mat_index(tb, l, col, ldb), b);
            mp_mul<G, R, true>(C, L, prod, a, b);
            mp_add<G, R, true>(C, L, sum, sum, prod);
        }
        store_num<G, R>(C, L, S, row + (long long) col * lds, sum);
    }
}

// ---- GEMM epilogue: C = alpha * S + beta * C with the reference's three roundings
//      (src/blas/gemm.cuh:142-166: K2-K4 on buffer, K2-K4 on C, K5-K6-K4) fused into one pass -------
template <int G, int R>
__global__ void k_gemm_epilogue(const DevConsts *Cp, int m, int n, SoA alpha, SoA beta, SoA S, int lds, SoA Cm, int ldc) {
    const DevConsts &C = *Cp;
    Lane<R> L;
    lane_init<G, R>(C, L);
    long long grp = ((long long) blockIdx.x * blockDim.x + threadIdx.x) / G;
    const long long ngrp = (long long) gridDim.x * blockDim.x / G;
    Num<R> al, be;
    load_num<G, R>(C, L, alpha, 0, al);
    load_num<G, R>(C, L, beta, 0, be);
    for (; grp < (long long) m * n; grp += ngrp) {
        const int row = (int) (grp % m), col = (int) (grp / m);
        Num<R> s, c, t1, t2;
        load_num<G, R>(C, L, S, row + (long long) col * lds, s);
        load_num<G, R>(C, L, Cm, row + (long long) col * ldc, c);
        mp_mul<G, R, true>(C, L, t1, s, al);
        mp_mul<G, R, true>(C, L, t2, c, be);
        mp_add<G, R, true>(C, L, c, t2, t1);
        store_num<G, R>(C, L, Cm, row + (long long) col * ldc, c);
    }
}

// ---- vector scale: r[i] = round(x[ix] * s[0]) (src/mpvector.cuh:139-216 + 669-714) -----------------
__device__ __forceinline__ long long inc_index(long long i, long long n, int inc) {
    return inc > 0 ? i * inc : (-n + i + 1) * (long long) inc;   // BLAS convention, mpvector.cuh:68-70
}
template <int G, int R>
__global__ void k_vec_scale(const DevConsts *Cp, long long n, SoA r, int incr, SoA x, int incx, SoA s) {
    const DevConsts &C = *Cp;
    Lane<R> L;
    lane_init<G, R>(C, L);
    long long grp = ((long long) blockIdx.x * blockDim.x + threadIdx.x) / G;
    const long long ngrp = (long long) gridDim.x * blockDim.x / G;
    Num<R> sc;
    load_num<G, R>(C, L, s, 0, sc);
    for (; grp < n; grp += ngrp) {
        Num<R> a, t;
        load_num<G, R>(C, L, x, inc_index(grp, n, incx), a);
        mp_mul<G, R, true>(C, L, t, a, sc);
        store_num<G, R>(C, L, r, inc_index(grp, n, incr), t);
    }
}

// y = round(round(alpha * x) + y): cuda::mp_axpy (src/blas/axpy.cuh:46-76: product into a buffer, rounding, sum, rounding), fused:
// no buffer, one pass over x and y.  The sum uses scalar mp_add semantics.
template <int G, int R>
__global__ void k_vec_axpy(const DevConsts *Cp, long long n, SoA s, SoA x, int incx, SoA y, int incy) {
    const DevConsts &C = *Cp;
    Lane<R> L;
    lane_init<G, R>(C, L);
    long long grp = ((long long) blockIdx.x * blockDim.x + threadIdx.x) / G;
    const long long ngrp = (long long) gridDim.x * blockDim.x / G;
    Num<R> sc;
    load_num<G, R>(C, L, s, 0, sc);
    for (; grp < n; grp += ngrp) {
        Num<R> a, b, t, r;
        load_num<G, R>(C, L, x, inc_index(grp, n, incx), a);
        const long long iy = inc_index(grp, n, incy);
        load_num<G, R>(C, L, y, iy, b);
        mp_mul<G, R, true>(C, L, t, a, sc);
        mp_add<G, R, true>(C, L, r, t, b);
        store_num<G, R>(C, L, y, iy, r);
    }
}

// w = round(round(beta * y) + round(alpha * x)): cuda::mp_waxpby (src/blas/waxpby.cuh:50-93), one pass
template <int G, int R>
__global__ void k_vec_waxpby(const DevConsts *Cp, long long n, SoA al, SoA x, int incx, SoA be, SoA y, int incy, SoA w, int incw) {
    const DevConsts &C = *Cp;
    Lane<R> L;
    lane_init<G, R>(C, L);
    long long grp = ((long long) blockIdx.x * blockDim.x + threadIdx.x) / G;
    const long long ngrp = (long long) gridDim.x * blockDim.x / G;
    Num<R> sa, sb;
    load_num<G, R>(C, L, al, 0, sa);
    load_num<G, R>(C, L, be, 0, sb);
    for (; grp < n; grp += ngrp) {
        Num<R> a, b, ta, tb, r;
        load_num<G, R>(C, L, x, inc_index(grp, n, incx), a);
        load_num<G, R>(C, L, y, inc_index(grp, n, incy), b);
        mp_mul<G, R, true>(C, L, ta, a, sa);
        mp_mul<G, R, true>(C, L, tb, b, sb);
        mp_add<G, R, true>(C, L, r, tb, ta);
        store_num<G, R>(C, L, w, inc_index(grp, n, incw), r);
    }
}

// C = round(round(beta * B) + round(alpha * A)), m x n column-major: cuda::mp_ge_add (src/blas/geadd.cuh:58-105) and, with C = B,
// cuda::mp_ge_acc (src/blas/geacc.cuh:57-97); one pass
template <int G, int R>
__global__ void k_ge_add(const DevConsts *Cp, int m, int n, SoA al, SoA A, int lda, SoA be, SoA B, int ldb, SoA Cm, int ldc) {
    const DevConsts &C = *Cp;
    Lane<R> L;
    lane_init<G, R>(C, L);
    long long grp = ((long long) blockIdx.x * blockDim.x + threadIdx.x) / G;
    const long long ngrp = (long long) gridDim.x * blockDim.x / G, total = (long long) m * n;
    Num<R> sa, sb;
    load_num<G, R>(C, L, al, 0, sa);
    load_num<G, R>(C, L, be, 0, sb);
    for (; grp < total; grp += ngrp) {
        const long long i = grp % m, j = grp / m;
        Num<R> a, b, ta, tb, r;
        load_num<G, R>(C, L, A, i + j * lda, a);
        load_num<G, R>(C, L, B, i + j * ldb, b);
        mp_mul<G, R, true>(C, L, ta, a, sa);
        mp_mul<G, R, true>(C, L, tb, b, sb);
        mp_add<G, R, true>(C, L, r, tb, ta);
        store_num<G, R>(C, L, Cm, i + j * ldc, r);
    }
}

// A = round(A + round(x_i * round(alpha * y_j))): cuda::mp_ger (src/blas/ger.cuh:157-206), one pass (alpha * y_j is recomputed per
// entry: a handful of instructions against the 2 (4N+40) bytes of A moved)
template <int G, int R>
__global__ void k_ger(const DevConsts *Cp, int m, int n, SoA al, SoA x, int incx, SoA y, int incy, SoA A, int lda) {
    const DevConsts &C = *Cp;
    Lane<R> L;
    lane_init<G, R>(C, L);
    long long grp = ((long long) blockIdx.x * blockDim.x + threadIdx.x) / G;
    const long long ngrp = (long long) gridDim.x * blockDim.x / G, total = (long long) m * n;
    Num<R> sa;
    load_num<G, R>(C, L, al, 0, sa);
    for (; grp < total; grp += ngrp) {
        const long long i = grp % m, j = grp / m;
        Num<R> xv, yv, ay, pr, a, r;
        load_num<G, R>(C, L, x, inc_index(i, m, incx), xv);
        load_num<G, R>(C, L, y, inc_index(j, n, incy), yv);
        load_num<G, R>(C, L, A, i + j * lda, a);
        mp_mul<G, R, true>(C, L, ay, yv, sa);
        mp_mul<G, R, true>(C, L, pr, xv, ay);
        mp_add<G, R, true>(C, L, r, a, pr);
        store_num<G, R>(C, L, A, i + j * lda, r);
    }
}

// w = round(w - round(alpha * v)): the first half of cuda::mp_axpy_dot (src/blas/axpydot.cuh:47-64; the difference is a sum with the
// sign of the product inverted, src/mpvector.cuh:512), one pass
template <int G, int R>
__global__ void k_vec_wsub(const DevConsts *Cp, long long n, SoA s, SoA v, int incv, SoA w, int incw) {
    const DevConsts &C = *Cp;
    Lane<R> L;
    lane_init<G, R>(C, L);
    long long grp = ((long long) blockIdx.x * blockDim.x + threadIdx.x) / G;
    const long long ngrp = (long long) gridDim.x * blockDim.x / G;
    Num<R> sc;
    load_num<G, R>(C, L, s, 0, sc);
    for (; grp < n; grp += ngrp) {
        Num<R> a, b, t, r;
        load_num<G, R>(C, L, v, inc_index(grp, n, incv), a);
        const long long iw = inc_index(grp, n, incw);
        load_num<G, R>(C, L, w, iw, b);
        mp_mul<G, R, true>(C, L, t, a, sc);
        t.sign ^= 1;
        mp_add<G, R, true>(C, L, r, b, t);
        store_num<G, R>(C, L, w, iw, r);
    }
}

// A = round(A D) (right side: column j times d_j) or round(D A) (left side: row i times d_i): cuda::mp_ge_diag_scale
// (src/blas/gediagscale.cuh:54-99), and, with both diagonals, A = round(round(DL A) DR): cuda::mp_ge_lr_scale
// (src/blas/gelrscale.cuh:56-93: left product, rounding, right product, rounding); one pass over A
template <int G, int R>
__global__ void k_ge_diag_scale(const DevConsts *Cp, int m, int n, SoA DL, int incdl, SoA DR, int incdr, bool left, bool right, SoA A, int lda) {
    const DevConsts &C = *Cp;
    Lane<R> L;
    lane_init<G, R>(C, L);
    long long grp = ((long long) blockIdx.x * blockDim.x + threadIdx.x) / G;
    const long long ngrp = (long long) gridDim.x * blockDim.x / G, total = (long long) m * n;
    for (; grp < total; grp += ngrp) {
        const long long i = grp % m, j = grp / m;
        Num<R> a, d, t;
        load_num<G, R>(C, L, A, i + j * lda, a);
        if (left) {
            load_num<G, R>(C, L, DL, inc_index(i, m, incdl), d);
            mp_mul<G, R, true>(C, L, t, a, d);
            a = t;
        }
        if (right) {
            load_num<G, R>(C, L, DR, inc_index(j, n, incdr), d);
            mp_mul<G, R, true>(C, L, t, a, d);
            a = t;
        }
        store_num<G, R>(C, L, A, i + j * lda, a);
    }
}

// Givens rotation, x = round(round(c x) + round(s y)), y = round(round(c y) - round(s x)): cuda::mp_rot (src/blas/rot.cuh:49-100:
// s x and s y into the buffers, mp_scal by c, sum / difference (the difference is a sum with the sign of s x inverted,
// src/mpvector.cuh:512), final rounding); unit increments, one pass, no buffers
template <int G, int R>
__global__ void k_vec_rot(const DevConsts *Cp, long long n, SoA x, SoA y, SoA c, SoA s) {
    const DevConsts &C = *Cp;
    Lane<R> L;
    lane_init<G, R>(C, L);
    long long grp = ((long long) blockIdx.x * blockDim.x + threadIdx.x) / G;
    const long long ngrp = (long long) gridDim.x * blockDim.x / G;
    Num<R> cc, ss;
    load_num<G, R>(C, L, c, 0, cc);
    load_num<G, R>(C, L, s, 0, ss);
    for (; grp < n; grp += ngrp) {
        Num<R> a, b, sx, sy, cx, cy, r;
        load_num<G, R>(C, L, x, grp, a);
        load_num<G, R>(C, L, y, grp, b);
        mp_mul<G, R, true>(C, L, sx, a, ss);
        mp_mul<G, R, true>(C, L, sy, b, ss);
        mp_mul<G, R, true>(C, L, cx, a, cc);
        mp_mul<G, R, true>(C, L, cy, b, cc);
        mp_add<G, R, true>(C, L, r, cx, sy);
        store_num<G, R>(C, L, x, grp, r);
        sx.sign ^= 1;
        mp_add<G, R, true>(C, L, r, cy, sx);
        store_num<G, R>(C, L, y, grp, r);
    }
}

// r = round(x + y) or round(x - y), strided: the last two steps of cuda::mp_rot with general increments
template <int G, int R>
__global__ void k_vec_addsub(const DevConsts *Cp, long long n, SoA x, int incx, SoA y, int incy, bool sub) {
    const DevConsts &C = *Cp;
    Lane<R> L;
    lane_init<G, R>(C, L);
    long long grp = ((long long) blockIdx.x * blockDim.x + threadIdx.x) / G;
    const long long ngrp = (long long) gridDim.x * blockDim.x / G;
    for (; grp < n; grp += ngrp) {
        Num<R> a, b, r;
        const long long ix = inc_index(grp, n, incx);
        load_num<G, R>(C, L, x, ix, a);
        load_num<G, R>(C, L, y, inc_index(grp, n, incy), b);
        if (sub) b.sign ^= 1;
        mp_add<G, R, true>(C, L, r, a, b);
        store_num<G, R>(C, L, x, ix, r);
    }
}

// ---- GEMV, reference order: y[o] = y[o] + sum_q op(A)(o, q) * ax[q]  (src/blas/gemv.cuh:199-218) ----
// y already holds round(beta * y) and ax = round(alpha * x).  One group per output element.
template <int G, int R>
__global__ void k_gemv_ref_order(const DevConsts *Cp, bool trans, int m, int n, SoA A, int lda, SoA ax, SoA y, int incy,
                                 const int *todo, const int *todo_count) {
    const DevConsts &C = *Cp;
    Lane<R> L;
    lane_init<G, R>(C, L);
    const int leny = trans ? n : m, lenx = trans ? m : n;
    long long it = ((long long) blockIdx.x * blockDim.x + threadIdx.x) / G;
    const long long ngrp = (long long) gridDim.x * blockDim.x / G;
    const long long total = todo ? (long long) *todo_count : (long long) leny;   // todo: outputs the fast path handed back
    for (; it < total; it += ngrp) {
        const long long grp = todo ? (long long) todo[it] : it;
        Num<R> sum, prod, a, b;
        num_zero(sum);
        for (int q = 0; q < lenx; ++q) {
            load_num<G, R>(C, L, A, trans ? q + grp * lda : grp + (long long) q * lda, a);
            load_num<G, R>(C, L, ax, q, b);
            mp_mul<G, R, true>(C, L, prod, a, b);
            mp_add<G, R, true>(C, L, sum, sum, prod);
        }
        const long long iy = inc_index(grp, leny, incy);
        load_num<G, R>(C, L, y, iy, a);
        mp_add<G, R, true>(C, L, a, a, sum);
        store_num<G, R>(C, L, y, iy, a);
    }
}

// ---- DOT, reference order: per-group serial mul/add over a strided slice, partials as AoS records
//      (structure of src/mpreduct.cuh:38-74 with groups in place of threads) --------------------------
template <int G, int R>
__global__ void k_dot_partial(const DevConsts *Cp, long long n, SoA x, int incx, SoA y, int incy, char *partials, const int *gate) {
    if (gate && *gate == 0) return;   // the fast path produced the result
    const DevConsts &C = *Cp;
    Lane<R> L;
    lane_init<G, R>(C, L);
    const long long grp = ((long long) blockIdx.x * blockDim.x + threadIdx.x) / G;
    const long long ngrp = (long long) gridDim.x * blockDim.x / G;
    Num<R> sum, prod, a, b;
    num_zero(sum);
    for (long long i = grp; i < n; i += ngrp) {
        load_num<G, R>(C, L, x, inc_index(i, n, incx), a);
        if (y.digits) {
            load_num<G, R>(C, L, y, inc_index(i, n, incy), b);
            mp_mul<G, R, true>(C, L, prod, a, b);
        } else {                                        // no second operand: sum of magnitudes (mp_asum, src/mpreduct.cuh:120-149)
            prod = a;
            prod.sign = 0;
        }
        mp_add<G, R, true>(C, L, sum, sum, prod);
    }
    store_rec<G, R>(C, L, partials, grp, sum);
}

// Sum `count` AoS records in index order with one group, result to SoA out[out_idx] or AoS record.
template <int G, int R>
__global__ void k_reduce_records(const DevConsts *Cp, const char *recs, long long count, SoA out, long long out_idx, char *out_rec) {
    const DevConsts &C = *Cp;
    Lane<R> L;
    lane_init<G, R>(C, L);
    if (threadIdx.x >= G || blockIdx.x > 0) return;
    Num<R> sum, a;
    num_zero(sum);
    for (long long i = 0; i < count; ++i) {
        load_rec<G, R>(C, L, recs, i, a);
        mp_add<G, R, true>(C, L, sum, sum, a);
    }
    if (out_rec) store_rec<G, R>(C, L, out_rec, 0, sum);
    else store_num<G, R>(C, L, out, out_idx, sum);
}

// Pairwise tree over AoS records held by the groups of ONE block: rec[g] += rec[g + stride].
template <int G, int R>
__global__ void k_tree_records(const DevConsts *Cp, char *recs, long long count, SoA out, long long out_idx, char *out_rec, const int *gate) {
    if (gate && *gate == 0) return;
    const DevConsts &C = *Cp;
    Lane<R> L;
    lane_init<G, R>(C, L);
    const int grp = threadIdx.x / G, ngrp = blockDim.x / G;
    // fold the tail onto the first ngrp records, then tree
    Num<R> sum, a;
    if (grp < count) {
        load_rec<G, R>(C, L, recs, grp, sum);
        for (long long i = grp + ngrp; i < count; i += ngrp) {
            load_rec<G, R>(C, L, recs, i, a);
            mp_add<G, R, true>(C, L, sum, sum, a);
        }
        store_rec<G, R>(C, L, recs, grp, sum);
    }
    __syncthreads();
    long long live = count < ngrp ? count : ngrp;
    int p = 1;
    while (p < live) p <<= 1;
    for (int s = p >> 1; s >= 1; s >>= 1) {
        if (grp < s && grp + s < live) {
            load_rec<G, R>(C, L, recs, grp, sum);
            load_rec<G, R>(C, L, recs, grp + s, a);
            mp_add<G, R, true>(C, L, sum, sum, a);
            store_rec<G, R>(C, L, recs, grp, sum);
        }
        __syncthreads();
    }
    if (grp == 0) {
        load_rec<G, R>(C, L, recs, 0, sum);
        if (count <= 0) num_zero(sum);
        if (out_rec) store_rec<G, R>(C, L, out_rec, 0, sum);
        else store_num<G, R>(C, L, out, out_idx, sum);
    }
}

// ---- device-side conversion from binary significands (replaces mp_set_mpfr, assign.cuh:86-127) -----
// One group per element.  limbs: little-endian 32-bit words.
template <int G, int R>
__global__ void k_set_binary(const DevConsts *Cp, SoA dst, long long offset, const int *sign, const int *exp,
                             const uint32_t *limbs, int nlimbs, long long count) {
    const DevConsts &C = *Cp;
    Lane<R> L;
    lane_init<G, R>(C, L);
    long long grp = ((long long) blockIdx.x * blockDim.x + threadIdx.x) / G;
    const long long ngrp = (long long) gridDim.x * blockDim.x / G;
    for (; grp < count; grp += ngrp) {
        const uint32_t *w = limbs + grp * nlimbs;
        int top = nlimbs;
        while (top > 0 && w[top - 1] == 0) --top;
        Num<R> x;
        num_zero(x);
        if (top > 0) {
            int tz = 0;
            while (w[tz >> 5] == 0) tz += 32;
            tz += __ffs(w[tz >> 5]) - 1;
            const int ws = tz >> 5, bs = tz & 31;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                unsigned long long acc = 0;
                for (int l = top - 1 - ws; l >= 0; --l) {
                    uint32_t lo = w[l + ws], hi = (l + ws + 1 < top) ? w[l + ws + 1] : 0u;
                    uint32_t v = bs ? ((lo >> bs) | (hi << (32 - bs))) : lo;
                    acc = ((acc << 32) | v) % (unsigned long long) (unsigned) L.m[r];
                }
                x.d[r] = L.act[r] ? (int) acc : 0;
            }
            x.sign = sign[grp] ? 1 : 0;
            x.exp = exp[grp] + tz;
            eval_compute<G, R, false>(C, L, x.d, x.lo, x.up);
        }
        store_num<G, R>(C, L, dst, offset + grp, x);
    }
}

}  // namespace mpres
