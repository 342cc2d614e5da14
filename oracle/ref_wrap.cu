/*
 * oracle/ref_wrap.cu -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * A thin extern "C" wrapper that compiles the UNMODIFIED reference headers where they lie under
 * /root/reference/src and exposes (a) the reference's host mp_float_t arithmetic, (b) its constants,
 * and (c) its CUDA kernels (v1 mp_gemm / mp_gemv / mp_dot, v2 mp_gemm) so tests and bench.py can run
 * them next to our implementation on identical inputs.  Built by oracle/Makefile into
 * oracle/_ref/libmpres_ref_N<k>.so, one binary per moduli set because the reference fixes the
 * precision at compile time (src/params.h).  The moduli set is chosen by force-including one of
 * /root/reference/src/params/32-bit-n-double-moduli/params.*.h, whose include guard (MPRES_PARAMS_H)
 * turns the later #include "params.h" into a no-op -- no reference source is copied or modified.
 *
 * No reference code is reproduced here; this file only calls it.
 */
#include <cstdint>
#include <cstdlib>
#include <cfloat>
#include <cstring>
#include <cmath>
#include <cstdio>
#include <string>
#include <sstream>
#include <vector>
#include <omp.h>

#include "arith/assign.cuh"
#include "arith/mul.cuh"
#include "arith/add.cuh"
#include "mparray.cuh"
#include "blas/gemm.cuh"
#include "blas/gemv.cuh"
#include "blas/dot.cuh"
#include "blas/asum.cuh"
#include "blas/norm.cuh"
#include "blas/genorm.cuh"
#include "mpcollection.cuh"
#include "sparse/mpmtx/spmv_mpmtx_csr2st.cuh"
#include "sparse/mpmtx/spmv_mpmtx_ell2st.cuh"
// the reference's own MPFR GEMM loop (tests/blas/v2/gemm/test_mpfr_gemm.cuh:29-60), included where it lies (-I $(REF)/tests)
#include "blas/v2/gemm/test_mpfr_gemm.cuh"

namespace v2 {
// the v2 header defines cuda::mp_gemm as a __global__ with a different signature; include it in the
// same TU (overloads coexist)
}
#include "blas/v2/gemm_v2.cuh"

static bool g_host_ready = false;

extern "C" {

/* ---- lifecycle / constants ------------------------------------------------------------------- */

/* Runs rns_const_init + mp_const_init (rns.cuh:324, arith_utils.cuh:44). Without a GPU the
 * cudaMemcpyToSymbol calls fail silently, which is fine for host-only use. */
void ref_init(void) {
    if (g_host_ready) return;
    rns_const_init();
    mp_const_init();
    cudaGetLastError();
    g_host_ready = true;
}

/* Re-upload constants to the current device (call after a GPU became current). */
void ref_gpu_init(void) {
    g_host_ready = false;
    ref_init();
    cudaDeviceSynchronize();
}

int ref_moduli_size(void) { return RNS_MODULI_SIZE; }
int ref_moduli_product_log2(void) { return RNS_MODULI_PRODUCT_LOG2; }
int ref_sizeof_mp_float(void) { return (int) sizeof(mp_float_t); }
int ref_mp_precision(void) { return MP_PRECISION; }
int ref_mp_h(void) { return MP_H; }
int ref_mp_j(void) { return MP_J; }
int ref_block_size_for_residues(void) { return BLOCK_SIZE_FOR_RESIDUES; }

void ref_get_moduli(int *out) { for (int i = 0; i < RNS_MODULI_SIZE; i++) out[i] = RNS_MODULI[i]; }
void ref_get_part_inverse(int *out) { memcpy(out, RNS_PART_MODULI_PRODUCT_INVERSE, sizeof(int) * RNS_MODULI_SIZE); }
void ref_get_pow2(int *out) { memcpy(out, RNS_POW2, sizeof(RNS_POW2)); }
void ref_get_m_pow2_residues(int *out) { memcpy(out, RNS_MODULI_PRODUCT_POW2_RESIDUES, sizeof(RNS_MODULI_PRODUCT_POW2_RESIDUES)); }
void ref_get_mi_pow2_residues(int *out) { memcpy(out, RNS_PART_MODULI_PRODUCT_POW2_RESIDUES, sizeof(RNS_PART_MODULI_PRODUCT_POW2_RESIDUES)); }
void ref_get_pow2_inverse(int *out) { memcpy(out, RNS_POW2_INVERSE, sizeof(RNS_POW2_INVERSE)); }
void ref_get_mrc_mult_inv(int *out) { memcpy(out, MRC_MULT_INV, sizeof(MRC_MULT_INV)); }
void ref_get_recip(double *rd, double *ru) {
    memcpy(rd, RNS_MODULI_RECIP_RD, sizeof(double) * RNS_MODULI_SIZE);
    memcpy(ru, RNS_MODULI_RECIP_RU, sizeof(double) * RNS_MODULI_SIZE);
}
/* out = {accuracy, unit.low.frac, unit.upp.frac, inv.low.frac, inv.upp.frac}; iout = {ref_factor,
 * unit.low.exp, unit.upp.exp, inv.low.exp, inv.upp.exp} */
void ref_get_eval_consts(double *out, long *iout) {
    out[0] = RNS_EVAL_ACCURACY;
    out[1] = RNS_EVAL_UNIT.low.frac; out[2] = RNS_EVAL_UNIT.upp.frac;
    out[3] = RNS_EVAL_INV_UNIT.low.frac; out[4] = RNS_EVAL_INV_UNIT.upp.frac;
    iout[0] = RNS_EVAL_REF_FACTOR;
    iout[1] = RNS_EVAL_UNIT.low.exp; iout[2] = RNS_EVAL_UNIT.upp.exp;
    iout[3] = RNS_EVAL_INV_UNIT.low.exp; iout[4] = RNS_EVAL_INV_UNIT.upp.exp;
}
void ref_get_mp_min(mp_float_t *out) { *out = MP_MIN; }

/* ---- conversions ------------------------------------------------------------------------------ */

/* value = (-1)^sign * mant * 2^exp with mant given as little-endian bytes; goes through an mpfr_t of
 * `prec` bits and the reference's own mp_set_mpfr (assign.cuh:86-127). */
void ref_set_from_int(mp_float_t *out, int sign, const unsigned char *mant_le, int nbytes, long exp, int prec) {
    mpz_t z; mpz_init(z);
    mpz_import(z, nbytes, -1, 1, 0, 0, mant_le);
    if (sign) mpz_neg(z, z);
    mpfr_t f; mpfr_init2(f, prec < 2 ? 2 : prec);
    mpfr_set_z_2exp(f, z, exp, MPFR_RNDN);
    mp_set_mpfr(out, f);
    mpfr_clear(f); mpz_clear(z);
}

void ref_set_d(mp_float_t *out, double x) { mp_set_d(out, x); }

/* CRT back to binary (rns.cuh:289-299): little-endian bytes of the significand; returns byte count */
int ref_get_mantissa(const mp_float_t *x, unsigned char *out_le, int cap) {
    mpz_t z; mpz_init(z);
    rns_to_binary(z, (int *) x->digits);
    size_t cnt = 0;
    memset(out_le, 0, cap);
    size_t need = (mpz_sizeinbase(z, 2) + 7) / 8;
    if ((int) need > cap) { mpz_clear(z); return -1; }
    mpz_export(out_le, &cnt, -1, 1, 0, 0, z);
    mpz_clear(z);
    return (int) cnt;
}

/* ---- host scalar arithmetic (the reference's CPU path) ----------------------------- */

void ref_host_mul(mp_float_t *r, const mp_float_t *x, const mp_float_t *y) { mp_mul(r, *x, *y); }
void ref_host_add(mp_float_t *r, const mp_float_t *x, const mp_float_t *y) { mp_add(r, *x, *y); }
void ref_host_round(mp_float_t *x, int bits) { mp_round(x, bits); }
void ref_host_eval(er_float_t *low, er_float_t *upp, const int *digits) { rns_eval_compute(low, upp, (int *) digits); }
void ref_host_eval_fast(er_float_t *low, er_float_t *upp, const int *digits) { rns_eval_compute_fast(low, upp, (int *) digits); }
void ref_host_scale2pow(int *res, const int *x, unsigned int D) { rns_scale2pow(res, (int *) x, D); }
int ref_host_mrc_compare(const int *x, const int *y) { return mrc_compare_rns((int *) x, (int *) y); }

void ref_host_mul_vec(mp_float_t *r, const mp_float_t *x, const mp_float_t *y, long n) {
    for (long i = 0; i < n; i++) mp_mul(&r[i], x[i], y[i]);
}
void ref_host_add_vec(mp_float_t *r, const mp_float_t *x, const mp_float_t *y, long n) {
    for (long i = 0; i < n; i++) mp_add(&r[i], x[i], y[i]);
}

/* r = sum_i x[i]*y[i], sequential mp_mul + mp_add (BASELINE config 1) */
void ref_host_dot(mp_float_t *r, const mp_float_t *x, const mp_float_t *y, long n) {
    mp_float_t acc = MP_ZERO, t;
    for (long i = 0; i < n; i++) {
        mp_mul(&t, x[i], y[i]);
        mp_add(&acc, acc, t);
    }
    *r = acc;
}

/* Threaded variant used as the CPU baseline: contiguous chunks per thread, partials added in
 * thread order.  Returns the number of threads used. */
int ref_host_dot_omp(mp_float_t *r, const mp_float_t *x, const mp_float_t *y, long n) {
    int nt = omp_get_max_threads();
    std::vector<mp_float_t> part(nt, MP_ZERO);
    #pragma omp parallel num_threads(nt)
    {
        int t = omp_get_thread_num();
        long lo = n * t / nt, hi = n * (t + 1) / nt;
        mp_float_t acc = MP_ZERO, p;
        for (long i = lo; i < hi; i++) { mp_mul(&p, x[i], y[i]); mp_add(&acc, acc, p); }
        part[t] = acc;
    }
    mp_float_t acc = MP_ZERO;
    for (int t = 0; t < nt; t++) mp_add(&acc, acc, part[t]);
    *r = acc;
    return nt;
}

/* Rows [row0,row1) of C = alpha*(A*B) + beta*C with the v1 mp_gemm semantics (gemm.cuh:39-58,
 * 142-166) evaluated with host scalar arithmetic; column-major, OpenMP over (row, col).
 * Returns the number of threads used. */
int ref_host_gemm_rows(int row0, int row1, int n, int k, const mp_float_t *alpha, const mp_float_t *A, int lda,
                       const mp_float_t *B, int ldb, const mp_float_t *beta, mp_float_t *C, int ldc) {
    int nt = omp_get_max_threads();
    #pragma omp parallel for collapse(2) schedule(static)
    for (int j = 0; j < n; j++) {
        for (int i = row0; i < row1; i++) {
            mp_float_t sum = MP_ZERO, mul;
            for (int l = 0; l < k; l++) {
                mp_mul(&mul, A[(size_t) lda * l + i], B[(size_t) ldb * j + l]);
                mp_add(&sum, sum, mul);
            }
            mp_float_t t1, t2;
            mp_mul(&t1, sum, *alpha);
            mp_mul(&t2, C[(size_t) ldc * j + i], *beta);
            mp_add(&C[(size_t) ldc * j + i], t2, t1);
        }
    }
    return nt;
}

/* y = alpha*A*x + beta*y rows [row0,row1) with v1 mp_gemv semantics on the host */
int ref_host_gemv_rows(int row0, int row1, int n, const mp_float_t *alpha, const mp_float_t *A, int lda,
                       const mp_float_t *x, const mp_float_t *beta, mp_float_t *y) {
    int nt = omp_get_max_threads();
    std::vector<mp_float_t> ax(n);
    for (int j = 0; j < n; j++) mp_mul(&ax[j], x[j], *alpha);
    #pragma omp parallel for schedule(static)
    for (int i = row0; i < row1; i++) {
        mp_float_t sum = MP_ZERO, mul, yb;
        for (int j = 0; j < n; j++) {
            mp_mul(&mul, A[(size_t) lda * j + i], ax[j]);
            mp_add(&sum, sum, mul);
        }
        mp_mul(&yb, y[i], *beta);
        mp_add(&y[i], yb, sum);
    }
    return nt;
}

/* ---- MPFR baselines (tests/blas/v2/gemm/test_mpfr_gemm.cuh:29-60 style) ------------------------ */

static void to_mpfr(mpfr_t out, const mp_float_t *x) { mp_get_mpfr(out, *x); }

/* dot in MPFR at `prec` bits, RNDN; result written as decimal string "%.*Re" */
void ref_mpfr_dot(const mp_float_t *x, const mp_float_t *y, long n, int prec, char *out, int cap) {
    mpfr_t acc, a, b, t;
    mpfr_init2(acc, prec); mpfr_init2(t, prec);
    mpfr_init2(a, 4 * MP_PRECISION + 64); mpfr_init2(b, 4 * MP_PRECISION + 64);
    mpfr_set_ui(acc, 0, MPFR_RNDN);
    for (long i = 0; i < n; i++) {
        to_mpfr(a, &x[i]); to_mpfr(b, &y[i]);
        mpfr_mul(t, a, b, MPFR_RNDN);
        mpfr_add(acc, acc, t, MPFR_RNDN);
    }
    mpfr_sprintf(out, "%.*Re", cap > 400 ? 300 : 60, acc);
    mpfr_clear(acc); mpfr_clear(a); mpfr_clear(b); mpfr_clear(t);
}

/* Timed MPFR loops on pre-converted operands (conversion excluded), OpenMP over chunks; returns secs */
double ref_mpfr_dot_timed(const mp_float_t *x, const mp_float_t *y, long n, int prec, int *threads) {
    mpfr_t *a = new mpfr_t[n], *b = new mpfr_t[n];
    #pragma omp parallel for
    for (long i = 0; i < n; i++) {
        mpfr_init2(a[i], prec); mpfr_init2(b[i], prec);
        to_mpfr(a[i], &x[i]); to_mpfr(b[i], &y[i]);
    }
    int nt = omp_get_max_threads();
    *threads = nt;
    double t0 = omp_get_wtime();
    #pragma omp parallel num_threads(nt)
    {
        int t = omp_get_thread_num();
        long lo = n * t / nt, hi = n * (t + 1) / nt;
        mpfr_t acc, p; mpfr_init2(acc, prec); mpfr_init2(p, prec);
        mpfr_set_ui(acc, 0, MPFR_RNDN);
        for (long i = lo; i < hi; i++) { mpfr_mul(p, a[i], b[i], MPFR_RNDN); mpfr_add(acc, acc, p, MPFR_RNDN); }
        mpfr_clear(acc); mpfr_clear(p);
    }
    double t1 = omp_get_wtime();
    for (long i = 0; i < n; i++) { mpfr_clear(a[i]); mpfr_clear(b[i]); }
    delete[] a; delete[] b;
    return t1 - t0;
}


/* MPFR@prec GEMM baseline: the reference's own mpfr_gemm (tests/blas/v2/gemm/test_mpfr_gemm.cuh:29-60, OpenMP over rows)
 * on an m x n block of C with the full inner dimension k; operands pre-converted (conversion excluded, as the reference's test does).
 * Column-major, lda = m, ldb = k, ldc = m.  Returns seconds; C (host AoS) is left untouched. */
double ref_mpfr_gemm_timed(int m, int n, int k, const mp_float_t *alpha, const mp_float_t *A, const mp_float_t *B, const mp_float_t *beta,
                           const mp_float_t *C, int prec, int *threads) {
    const size_t na = (size_t) m * k, nb = (size_t) k * n, nc = (size_t) m * n;
    mpfr_t *a = new mpfr_t[na], *b = new mpfr_t[nb], *c = new mpfr_t[nc];
    mpfr_t al, be;
    mpfr_init2(al, prec); mpfr_init2(be, prec);
    to_mpfr(al, alpha); to_mpfr(be, beta);
    #pragma omp parallel for
    for (long i = 0; i < (long) na; i++) { mpfr_init2(a[i], prec); to_mpfr(a[i], &A[i]); }
    #pragma omp parallel for
    for (long i = 0; i < (long) nb; i++) { mpfr_init2(b[i], prec); to_mpfr(b[i], &B[i]); }
    #pragma omp parallel for
    for (long i = 0; i < (long) nc; i++) { mpfr_init2(c[i], prec); to_mpfr(c[i], &C[i]); }
    *threads = omp_get_max_threads();
    double t0 = omp_get_wtime();
    mpfr_gemm(mblas_no_trans, mblas_no_trans, m, n, k, al, a, m, b, k, be, c, m);
    double t1 = omp_get_wtime();
    for (size_t i = 0; i < na; i++) mpfr_clear(a[i]);
    for (size_t i = 0; i < nb; i++) mpfr_clear(b[i]);
    for (size_t i = 0; i < nc; i++) mpfr_clear(c[i]);
    delete[] a; delete[] b; delete[] c;
    mpfr_clear(al); mpfr_clear(be);
    return t1 - t0;
}

/* decimal strings ("%.*Re", `digits` significant digits) of records, through the reference's mp_get_mpfr: for accuracy checks */
void ref_to_string(const mp_float_t *x, int prec, int digits, char *out, int cap) {
    mpfr_t f; mpfr_init2(f, prec);
    to_mpfr(f, x);
    (void) cap; mpfr_sprintf(out, "%.*Re", digits, f);
    mpfr_clear(f);
}

/* ---- reference CUDA kernels (run on the GPU box only) ----------------------------------------- */

static void upload(mp_array_t &d, const mp_float_t *h, size_t n) {
    cuda::mp_array_init(d, n);
    std::vector<int> dig(n * RNS_MODULI_SIZE), sg(n), ex(n);
    std::vector<er_float_t> ev(2 * n);
    for (size_t i = 0; i < n; i++) {
        memcpy(&dig[i * RNS_MODULI_SIZE], h[i].digits, sizeof(int) * RNS_MODULI_SIZE);
        sg[i] = h[i].sign; ex[i] = h[i].exp; ev[i] = h[i].eval[0]; ev[n + i] = h[i].eval[1];
    }
    cudaMemcpy(d.digits, dig.data(), dig.size() * sizeof(int), cudaMemcpyHostToDevice);
    cudaMemcpy(d.sign, sg.data(), n * sizeof(int), cudaMemcpyHostToDevice);
    cudaMemcpy(d.exp, ex.data(), n * sizeof(int), cudaMemcpyHostToDevice);
    cudaMemcpy(d.eval, ev.data(), 2 * n * sizeof(er_float_t), cudaMemcpyHostToDevice);
}

static void download(mp_float_t *h, mp_array_t &d, size_t n) {
    std::vector<int> dig(n * RNS_MODULI_SIZE), sg(n), ex(n);
    std::vector<er_float_t> ev(2 * n);
    cudaMemcpy(dig.data(), d.digits, dig.size() * sizeof(int), cudaMemcpyDeviceToHost);
    cudaMemcpy(sg.data(), d.sign, n * sizeof(int), cudaMemcpyDeviceToHost);
    cudaMemcpy(ex.data(), d.exp, n * sizeof(int), cudaMemcpyDeviceToHost);
    cudaMemcpy(ev.data(), d.eval, 2 * n * sizeof(er_float_t), cudaMemcpyDeviceToHost);
    for (size_t i = 0; i < n; i++) {
        memcpy(h[i].digits, &dig[i * RNS_MODULI_SIZE], sizeof(int) * RNS_MODULI_SIZE);
        h[i].sign = sg[i]; h[i].exp = ex[i]; h[i].eval[0] = ev[i]; h[i].eval[1] = ev[n + i];
    }
}

/* v1 mp_gemm with the launch configuration of tests/blas/performance/test_gemm_performance.cu:53-57.
 * C (host AoS, ldc == m) is overwritten with the result; if AB_out != NULL it receives the raw output
 * of matrix_multiply_notrans_kernel (before the alpha/beta epilogue).  Returns kernel time in ms
 * (CUDA events around the mp_gemm call, `repeat` calls averaged; C restored between repeats). */
float ref_gpu_gemm(int m, int n, int k, const mp_float_t *alpha, const mp_float_t *A, const mp_float_t *B,
                   const mp_float_t *beta, mp_float_t *C, mp_float_t *AB_out, int repeat) {
    mp_array_t dA, dB, dC, dBuf, dAlpha, dBeta;
    upload(dA, A, (size_t) m * k); upload(dB, B, (size_t) k * n); upload(dC, C, (size_t) m * n);
    upload(dAlpha, alpha, 1); upload(dBeta, beta, 1);
    cuda::mp_array_init(dBuf, (size_t) m * n);
    if (AB_out) {
        dim3 block(16, 16), grid((n + 15) / 16, (m + 15) / 16);
        cuda::matrix_multiply_notrans_kernel<<<grid, block>>>(m, n, k, dAlpha, dA, m, dB, k, dBuf, m);
        cudaDeviceSynchronize();
        download(AB_out, dBuf, (size_t) m * n);
    }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float total = 0;
    if (repeat < 1) repeat = 1;
    for (int r = 0; r < repeat; r++) {
        if (r > 0) { cuda::mp_array_clear(dC); upload(dC, C, (size_t) m * n); }
        cudaEventRecord(e0);
        cuda::mp_gemm<32, 1, 128, 64, 16>(mblas_no_trans, mblas_no_trans, m, n, k, dAlpha, dA, m, dB, k, dBeta, dC, m, dBuf);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); total += ms;
    }
    download(C, dC, (size_t) m * n);
    cuda::mp_array_clear(dA); cuda::mp_array_clear(dB); cuda::mp_array_clear(dC); cuda::mp_array_clear(dBuf);
    cuda::mp_array_clear(dAlpha); cuda::mp_array_clear(dBeta);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return total / repeat;
}

/* v2 single-kernel mp_gemm (blas/v2/gemm_v2.cuh:55-81), 16x16 blocks; AoS in/out. */
float ref_gpu_gemm_v2(int transa, int transb, int m, int n, int k, const mp_float_t *alpha, const mp_float_t *A, int lda,
                      const mp_float_t *B, int ldb, const mp_float_t *beta, mp_float_t *C, int ldc,
                      size_t sizeA, size_t sizeB, size_t sizeC) {
    mp_float_t *dA, *dB, *dC, *dAl, *dBe;
    cudaMalloc(&dA, sizeA * sizeof(mp_float_t)); cudaMalloc(&dB, sizeB * sizeof(mp_float_t)); cudaMalloc(&dC, sizeC * sizeof(mp_float_t));
    cudaMalloc(&dAl, sizeof(mp_float_t)); cudaMalloc(&dBe, sizeof(mp_float_t));
    cudaMemcpy(dAl, alpha, sizeof(mp_float_t), cudaMemcpyHostToDevice);
    cudaMemcpy(dBe, beta, sizeof(mp_float_t), cudaMemcpyHostToDevice);
    cudaMemcpy(dA, A, sizeA * sizeof(mp_float_t), cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B, sizeB * sizeof(mp_float_t), cudaMemcpyHostToDevice);
    cudaMemcpy(dC, C, sizeC * sizeof(mp_float_t), cudaMemcpyHostToDevice);
    dim3 block(16, 16), grid((n + 15) / 16, (m + 15) / 16);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    cuda::mp_gemm<<<grid, block>>>((enum mblas_trans_type) transa, (enum mblas_trans_type) transb, m, n, k, dAl, dA, lda, dB, ldb, dBe, dC, ldc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(C, dC, sizeC * sizeof(mp_float_t), cudaMemcpyDeviceToHost);
    cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dAl); cudaFree(dBe);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return ms;
}

/* v1 mp_gemv, launch configuration of test_gemv_performance.cu:51-54; inc = 1. trans: 111 / 112. */
float ref_gpu_gemv(int trans, int m, int n, const mp_float_t *alpha, const mp_float_t *A, const mp_float_t *x,
                   const mp_float_t *beta, mp_float_t *y, int repeat) {
    int lenx = trans == mblas_no_trans ? n : m, leny = trans == mblas_no_trans ? m : n;
    mp_array_t dA, dx, dy, dAlpha, dBeta, dB1, dB2;
    upload(dA, A, (size_t) m * n); upload(dx, x, lenx); upload(dy, y, leny);
    upload(dAlpha, alpha, 1); upload(dBeta, beta, 1);
    cuda::mp_array_init(dB1, lenx); cuda::mp_array_init(dB2, (size_t) m * n);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float total = 0;
    if (repeat < 1) repeat = 1;
    for (int r = 0; r < repeat; r++) {
        if (r > 0) { cuda::mp_array_clear(dy); upload(dy, y, leny); }
        cudaEventRecord(e0);
        cuda::mp_gemv<256, 128, 256, 32>((enum mblas_trans_type) trans, m, n, dAlpha, dA, m, dx, 1, dBeta, dy, 1, dB1, dB2);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); total += ms;
    }
    download(y, dy, leny);
    cuda::mp_array_clear(dA); cuda::mp_array_clear(dx); cuda::mp_array_clear(dy); cuda::mp_array_clear(dAlpha);
    cuda::mp_array_clear(dBeta); cuda::mp_array_clear(dB1); cuda::mp_array_clear(dB2);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return total / repeat;
}

/* v1 mp_dot; cfg = 0: performance launch config (test_dot_performance.cu:41-45),
 * cfg = 1: accuracy-test config (test_dot_accuracy.cu:28-33). inc = 1. */
float ref_gpu_dot(int n, const mp_float_t *x, const mp_float_t *y, mp_float_t *r, int cfg, int repeat) {
    mp_array_t dx, dy, dr, dbuf;
    upload(dx, x, n); upload(dy, y, n);
    cuda::mp_array_init(dr, 1); cuda::mp_array_init(dbuf, n);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float total = 0;
    if (repeat < 1) repeat = 1;
    for (int q = 0; q < repeat; q++) {
        cudaEventRecord(e0);
        if (cfg == 0) cuda::mp_dot<512, 128, 8192, 256, 64>(n, dx, 1, dy, 1, dr, dbuf);
        else cuda::mp_dot<512, 128, 32768, 64, 64>(n, dx, 1, dy, 1, dr, dbuf);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); total += ms;
    }
    download(r, dr, 1);
    cuda::mp_array_clear(dx); cuda::mp_array_clear(dy); cuda::mp_array_clear(dr); cuda::mp_array_clear(dbuf);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return total / repeat;
}


/* v1 mp_asum / mp_norm (src/blas/asum.cuh:41, norm.cuh:43) with the launch configuration <grid, block> = <256, 64>
 * (tests/blas/performance/test_asum_performance.cu:42-43) or, cfg = 1, <128, 32> (test_norm_performance.cu:49-50).
 * kind: 0 asum, 171 one-norm, 175 inf-norm.  incx > 0; x holds 1 + (n - 1) * incx records. */
float ref_gpu_asum_norm(int kind, int n, const mp_float_t *x, int incx, mp_float_t *r, int cfg) {
    mp_array_t dx, dr;
    upload(dx, x, (size_t) 1 + (size_t) (n - 1) * incx); cuda::mp_array_init(dr, 1);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    if (kind == 0) { if (cfg == 0) cuda::mp_asum<256, 64>(n, dx, incx, dr); else cuda::mp_asum<128, 32>(n, dx, incx, dr); }
    else { if (cfg == 0) cuda::mp_norm<256, 64>((enum mblas_norm_type) kind, n, dx, incx, dr); else cuda::mp_norm<128, 32>((enum mblas_norm_type) kind, n, dx, incx, dr); }
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    download(r, dr, 1);
    cuda::mp_array_clear(dx); cuda::mp_array_clear(dr);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return ms;
}

/* v1 mp_ge_norm (src/blas/genorm.cuh:142), <gridDim1, blockDim1> = <128, 32>; A is lda x n column-major */
float ref_gpu_ge_norm(int kind, int m, int n, const mp_float_t *A, int lda, mp_float_t *r) {
    mp_array_t dA, dr, dbuf;
    upload(dA, A, (size_t) lda * n); cuda::mp_array_init(dr, 1); cuda::mp_array_init(dbuf, (size_t) (m > n ? m : n));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    cuda::mp_ge_norm<128, 32>((enum mblas_norm_type) kind, m, n, dA, lda, dr, dbuf);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    download(r, dr, 1);
    cuda::mp_array_clear(dA); cuda::mp_array_clear(dr); cuda::mp_array_clear(dbuf);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return ms;
}

/* two-stage SpMV over mp_collection_t (src/sparse/mpmtx/spmv_mpmtx_csr2st.cuh:106, spmv_mpmtx_ell2st.cuh:119), launch configuration
 * <32, 32, 32, 32>.  fmt 0: CSR (ptr = irp[m + 1], idx = ja[nnz], as[nnz]); fmt 1: ELLPACK (idx = ja[m * maxnzr] column-major,
 * as[m * maxnzr], nnz := maxnzr).  x: n records, y: m records out. */
float ref_gpu_spmv_2st(int fmt, int m, int n, int nnz, const int *ptr, const int *idx, const mp_float_t *as, const mp_float_t *x, mp_float_t *y) {
    const size_t cnt = fmt == 0 ? (size_t) nnz : (size_t) m * nnz;
    mp_collection_t das, dbuf;
    mp_array_t dx, dy;
    cuda::mp_collection_init(das, cnt); cuda::mp_collection_init(dbuf, cnt);
    /* the reference rounds EVERY slot of the buffer (mp_vector_round_kernel over m * maxnzr entries), the ELLPACK padding included, which no kernel wrote:
     * leftovers of earlier allocations there can keep its scaling loop busy for minutes.  A fresh process hands out zeroed memory; this makes it so always. */
    cudaMemset(dbuf.digits, 0, sizeof(int) * RNS_MODULI_SIZE * cnt); cudaMemset(dbuf.sign, 0, sizeof(int) * cnt); cudaMemset(dbuf.exp, 0, sizeof(int) * cnt);
    cudaMemset(dbuf.eval, 0, sizeof(er_float_t) * 2 * cnt);
    cuda::mp_collection_host2device(das, (mp_float_ptr) as, cnt);
    upload(dx, x, n); cuda::mp_array_init(dy, m);
    int *dptr = nullptr, *didx = nullptr;
    if (fmt == 0) { cudaMalloc(&dptr, sizeof(int) * (m + 1)); cudaMemcpy(dptr, ptr, sizeof(int) * (m + 1), cudaMemcpyHostToDevice); }
    cudaMalloc(&didx, sizeof(int) * cnt); cudaMemcpy(didx, idx, sizeof(int) * cnt, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    if (fmt == 0) cuda::mp_spmv_mpmtx_csr2st<32, 32, 32, 32>(m, n, nnz, dptr, didx, das, dx, dy, dbuf);
    else cuda::mp_spmv_mpmtx_ell2st<32, 32, 32, 32>(m, n, nnz, didx, das, dx, dy, dbuf);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    download(y, dy, m);
    cuda::mp_collection_clear(das); cuda::mp_collection_clear(dbuf); cuda::mp_array_clear(dx); cuda::mp_array_clear(dy);
    if (dptr) cudaFree(dptr);
    cudaFree(didx);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return ms;
}

} // extern "C"

/* element-wise probes of the device scalar routines: one thread per element */
__global__ void k_probe_mul(mp_float_t *r, const mp_float_t *x, const mp_float_t *y, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { mp_float_t t; cuda::mp_mul(&t, x[i], y[i]); r[i] = t; }
}
__global__ void k_probe_add(mp_float_t *r, const mp_float_t *x, const mp_float_t *y, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { mp_float_t t; cuda::mp_add(&t, x[i], y[i]); r[i] = t; }
}
__global__ void k_probe_eval(mp_float_t *r, const mp_float_t *x, int n, int fast) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        mp_float_t t = x[i];
        if (fast) cuda::rns_eval_compute_fast(&t.eval[0], &t.eval[1], t.digits);
        else cuda::rns_eval_compute(&t.eval[0], &t.eval[1], t.digits);
        r[i] = t;
    }
}
__global__ void k_probe_round(mp_float_t *r, const mp_float_t *x, const int *bits, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { mp_float_t t = x[i]; cuda::mp_round(&t, bits[i]); r[i] = t; }
}

extern "C" {

/* op: 0 mul, 1 add, 2 eval (full), 3 eval_fast, 4 round by bits[i] */
void ref_gpu_probe(int op, mp_float_t *r, const mp_float_t *x, const mp_float_t *y, const int *bits, int n) {
    mp_float_t *dr, *dx, *dy = nullptr; int *db = nullptr;
    cudaMalloc(&dr, n * sizeof(mp_float_t)); cudaMalloc(&dx, n * sizeof(mp_float_t));
    cudaMemcpy(dx, x, n * sizeof(mp_float_t), cudaMemcpyHostToDevice);
    if (y) { cudaMalloc(&dy, n * sizeof(mp_float_t)); cudaMemcpy(dy, y, n * sizeof(mp_float_t), cudaMemcpyHostToDevice); }
    if (bits) { cudaMalloc(&db, n * sizeof(int)); cudaMemcpy(db, bits, n * sizeof(int), cudaMemcpyHostToDevice); }
    int bs = 64, gs = (n + bs - 1) / bs;
    if (op == 0) k_probe_mul<<<gs, bs>>>(dr, dx, dy, n);
    else if (op == 1) k_probe_add<<<gs, bs>>>(dr, dx, dy, n);
    else if (op == 2) k_probe_eval<<<gs, bs>>>(dr, dx, n, 0);
    else if (op == 3) k_probe_eval<<<gs, bs>>>(dr, dx, n, 1);
    else if (op == 4) k_probe_round<<<gs, bs>>>(dr, dx, db, n);
    cudaDeviceSynchronize();
    cudaMemcpy(r, dr, n * sizeof(mp_float_t), cudaMemcpyDeviceToHost);
    cudaFree(dr); cudaFree(dx); if (dy) cudaFree(dy); if (db) cudaFree(db);
}

const char *ref_last_cuda_error(void) { return cudaGetErrorString(cudaGetLastError()); }

} // extern "C"
