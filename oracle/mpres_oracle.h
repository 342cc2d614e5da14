/*
 * oracle/mpres_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the multiple-precision arithmetic on the mp_gemm / mp_gemv / mp_dot
 * path of MPRES-BLAS.  It is the checker for tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg; nothing under mpres-blas_b200/ may include, link or call it.
 *
 * Two flavours, because the reference itself has two:
 *   ORC_HOST   -- the reference's CPU functions: directed rounding EMULATED by widening
 *                 (dinterval.cuh:55-165), double-reciprocal modular multiply (modular.cuh:90-95),
 *                 host rank rule in rns_scale2pow (rns.cuh:1045).
 *   ORC_DEVICE -- the reference's cuda:: functions: IEEE directed rounding (__dadd_rd ...), exact
 *                 64-bit %, device rank rule (rns.cuh:1157).
 * Records use the reference's AoS mp_float_t layout (types.cuh:69-74) for a run-time N:
 *   int digits[N]; int sign; int exp; struct { double frac; long exp; } eval[2];   (4N + 40 bytes)
 */
#ifndef MPRES_ORACLE_H
#define MPRES_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_HOST 0
#define ORC_DEVICE 1
#define ORC_MAX_N 128

typedef struct { double frac; long exp; } orc_er_t;

typedef struct orc_ctx orc_ctx;

orc_ctx *orc_create(int N, int log2M, int flavor, int mp_h, int mp_j, int ref_factor, double accuracy,
                    const orc_er_t *unit_low, const orc_er_t *unit_upp, const orc_er_t *inv_low,
                    const orc_er_t *inv_upp, const int *moduli, const int *part_inverse, const int *pow2,
                    const int *m_pow2, const int *mi_pow2, const int *pow2_inv, const int *mrc_inv,
                    const double *recip_rd, const double *recip_ru);
void orc_destroy(orc_ctx *c);
int orc_record_size(const orc_ctx *c);

/* scalar ops on AoS records (void* = one record) */
void orc_mp_mul(const orc_ctx *c, void *r, const void *x, const void *y);
void orc_mp_add(const orc_ctx *c, void *r, const void *x, const void *y);
void orc_mp_round(const orc_ctx *c, void *x, int bits);
void orc_eval_compute(const orc_ctx *c, orc_er_t *low, orc_er_t *upp, const int *digits);
void orc_eval_compute_fast(const orc_ctx *c, orc_er_t *low, orc_er_t *upp, const int *digits);
void orc_scale2pow(const orc_ctx *c, int *result, const int *x, unsigned int D);
int orc_mrc_compare(const orc_ctx *c, const int *x, const int *y);
/* significand given as nlimbs little-endian 32-bit limbs; trims trailing zero bits like
 * mp_set_mpfr (assign.cuh:95-111) and evaluates with the full rns_eval_compute */
void orc_mp_set(const orc_ctx *c, void *r, int sign, const uint32_t *limbs, int nlimbs, int exp);

/* bulk synthetic records for CPU-baseline timing: uniform(-1,1) magnitudes, `bits`-bit significands
 * (top 53 bits from a double product as in tests/tsthelper.cuh:63-65, low bits uniform), OpenMP */
void orc_random_fill(const orc_ctx *c, void *recs, long n, int bits, uint64_t seed);

/* vector / BLAS-level restatements (reference-order summation) */
void orc_mul_vec(const orc_ctx *c, void *r, const void *x, const void *y, long n);
void orc_add_vec(const orc_ctx *c, void *r, const void *x, const void *y, long n);
void orc_dot_seq(const orc_ctx *c, void *r, const void *x, const void *y, long n);
int orc_dot_omp(const orc_ctx *c, void *r, const void *x, const void *y, long n);
/* v1 mp_dot structure (dot.cuh:84-107, mpreduct.cuh:38-112): element-wise rounded products, then the
 * two-pass tree sum with `grid` blocks of `block` threads */
void orc_dot_v1(const orc_ctx *c, void *r, const void *x, const void *y, long n, int grid, int block);
/* v1 mp_gemm semantics, rows [row0,row1) (gemm.cuh:39-58 + 142-166); returns threads used.
 * AB_out (may be NULL, ld = m) receives the k-loop result before the alpha/beta epilogue. */
int orc_gemm_rows(const orc_ctx *c, int row0, int row1, int m, int n, int k, const void *alpha, const void *A,
                  int lda, const void *B, int ldb, const void *beta, void *C, int ldc, void *AB_out);
/* v1 mp_gemv semantics (gemv.cuh:150-268), inc = 1; trans = 111 (N) or 112 (T); block = blockDim3 */
int orc_gemv(const orc_ctx *c, int trans, int m, int n, const void *alpha, const void *A, int lda,
             const void *x, const void *beta, void *y, int block);

#ifdef __cplusplus
}
#endif
#endif
