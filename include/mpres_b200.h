/*
 * mpres_b200.h -- C-ABI of the B200-native mp_gemm / mp_gemv / mp_dot path.
 *
 * This is the drop-in boundary for the multiple-precision GEMM path of MPRES-BLAS (kisupov/mpres-blas
 * v1.7.0).  Every entry point names the reference interface it replaces (paths relative to the
 * reference root).  Plain pointers and sizes only; all mp_array_t / mp_collection_t members are DEVICE
 * pointers exactly as in the reference (src/mparray.cuh:35-54), scalars (alpha, beta, r) are
 * length-1 device arrays (tests/blas/performance/test_gemm_performance.cu:156-157).
 *
 * Return value of every function: 0 = ok, < 0 = invalid argument (the reference returns silently,
 * src/blas/gemm.cuh:75-96), > 0 = cudaError_t.  There is no CPU fallback: if the CUDA library cannot
 * run, calls fail.
 */
#ifndef MPRES_B200_H
#define MPRES_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- data types, byte-identical to src/types.cuh ------------------------------------------------- */

/* src/types.cuh:46-49 */
typedef struct {
    double frac;
    long exp;
} mpres_er_float_t;

/* CUDA's int4 without the cuda headers (16-byte aligned in the reference; only ever used through a
 * device pointer here) */
typedef struct { int x, y, z, w; } mpres_int4;

/* src/types.cuh:85-92 -- SoA array with scratch `buf` and device-resident length `len` */
typedef struct {
    int *digits;             /* [len * N], element-major: all residues of x0, all residues of x1, ... */
    int *sign;               /* [len] */
    int *exp;                /* [len] */
    mpres_er_float_t *eval;  /* [2 * len]: lower bounds of all elements, then upper bounds */
    mpres_int4 *buf;         /* [len] scratch of the reference's two-kernel additions; unused by this library */
    int *len;                /* device scalar: ALLOCATED length (offset of the upper bounds) */
} mpres_array_t;

/* src/types.cuh:99-104 -- SoA array without buf/len; the length travels as an argument */
typedef struct {
    int *digits;
    int *sign;
    int *exp;
    mpres_er_float_t *eval;
} mpres_collection_t;

/* src/blas/mblas_enum.cuh:25-29 */
enum { MPRES_NO_TRANS = 111, MPRES_TRANS = 112, MPRES_CONJ_TRANS = 113 };
enum { MPRES_LEFT_SIDE = 141, MPRES_RIGHT_SIDE = 142 };   /* mblas_side_type, src/blas/mblas_enum.cuh:37-40 */

/* Stage-2 strategy.  AUTO = exact-window fast path wherever its guard holds, reference-order
 * fallback per element otherwise.  REFERENCE_ORDER = the k-loop of src/blas/gemm.cuh:46-49 step by
 * step (mp_mul, mp_add, rounding after each), residue-parallel: bit-identical to the reference
 * kernels including the interval evaluations.  FAST = fast path only (elements whose guard fails
 * are reported through mpres_last_fallback_count). */
enum { MPRES_MODE_AUTO = 0, MPRES_MODE_REFERENCE_ORDER = 1, MPRES_MODE_FAST = 2 };
enum { MPRES_STAGE2_UMMA = 0, MPRES_STAGE2_UMMA_UNSTACKED = 1, MPRES_STAGE2_MMA_SYNC = 2, MPRES_STAGE2_SMALL = 3, MPRES_STAGE2_SMALL_TILED = 4, MPRES_STAGE2_SMALL_K64 = 5, MPRES_STAGE2_SMALL_T128 = 6 };

typedef struct mpres_ctx mpres_ctx;
typedef void *mpres_stream_t; /* cudaStream_t */

/* ---- lifecycle: replaces rns_const_init() + mp_const_init() (src/rns.cuh:324-442,
 *      src/arith/arith_utils.cuh:44-85), which need GMP/MPFR and one precision per binary -------- */

/* moduli_size selects one of the predefined sets of src/params/32-bit-n-double-moduli/
 * (8, 16, 24, 32, 40, 48, 56, 64 or 128 moduli); constants are uploaded to `device`. */
int mpres_init(mpres_ctx **ctx, int moduli_size, int device);
/* any pairwise-coprime odd moduli < 2^31 (what overwriting src/params.h does in the reference); 2 <= moduli_size <= 128 and even
 * (an odd count would pad the reference's mp_float_t to 4N + 44 bytes: -15), more than 128 moduli: -11 */
int mpres_init_moduli(mpres_ctx **ctx, const int *moduli, int moduli_size, int device);
int mpres_finalize(mpres_ctx *ctx);

/* RNS_MODULI_SIZE, RNS_MODULI_PRODUCT_LOG2 (src/params.h:31-36), MP_PRECISION, MP_H, MP_J
 * (src/arith/arith_utils.cuh:33-35) */
int mpres_moduli_size(const mpres_ctx *ctx);
int mpres_moduli_product_log2(const mpres_ctx *ctx);
int mpres_precision(const mpres_ctx *ctx);
int mpres_mp_h(const mpres_ctx *ctx);
int mpres_mp_j(const mpres_ctx *ctx);
int mpres_device(const mpres_ctx *ctx);
/* sizeof(mp_float_t) for this moduli set = 4N + 40 (src/types.cuh:69-74) */
size_t mpres_sizeof_mp_float(const mpres_ctx *ctx);
/* Host copies of the constant tables, for tests: which = 0 RNS_MODULI, 1 RNS_PART_MODULI_PRODUCT_INVERSE,
 * 2 RNS_POW2, 3 RNS_MODULI_PRODUCT_POW2_RESIDUES, 4 RNS_PART_MODULI_PRODUCT_POW2_RESIDUES,
 * 5 RNS_POW2_INVERSE, 6 MRC_MULT_INV (int tables); 7 RNS_MODULI_RECIP_RD, 8 RNS_MODULI_RECIP_RU,
 * 9 {RNS_EVAL_ACCURACY, UNIT.low.frac, UNIT.upp.frac, INV_UNIT.low.frac, INV_UNIT.upp.frac},
 * 10 {RNS_EVAL_REF_FACTOR, UNIT.low.exp, UNIT.upp.exp, INV_UNIT.low.exp, INV_UNIT.upp.exp} as doubles.
 * Returns the number of bytes written (<= cap) or < 0. */
long mpres_get_constant(const mpres_ctx *ctx, int which, void *out, size_t cap);

int mpres_set_mode(mpres_ctx *ctx, int mode);
int mpres_get_mode(const mpres_ctx *ctx);
/* Stage-2 kernel of the fast path.  SMALL (default) = the exact sums are accumulated modulo one-byte moduli
 * (256, 251, 243, ...): one tcgen05.mma kind::i8 GEMM per modulus, inputs converted through their binary
 * representation, results returned to the moduli of the number format by a CRT base extension; chosen per call
 * when the sums fit the small base (about 360 bits), otherwise the call runs as UMMA; a persistent kernel (one
 * CTA per SM, 256 x 256 tiles, 128-byte operand rows), SMALL_T128 = 128 x 256 tiles with a double-buffered
 * accumulator, SMALL_K64 = the latter with 64-byte operand rows,
 * SMALL_TILED = one tile per CTA.  UMMA = tcgen05.mma
 * kind::i8 over four byte limbs of the format's own moduli with TMA-fed limb tiles, UMMA_UNSTACKED = the same
 * kernel issuing one MMA per limb pair, MMA_SYNC = the legacy warp-level int8 MMA kernel.  All of them produce
 * identical residues; the switch exists for A/B measurement. */
int mpres_set_stage2_kernel(mpres_ctx *ctx, int kind);
/* Small-modulus base the last fast-path call used: *moduli = how many one-byte moduli (0: the call ran on the
 * format's own moduli), *input_moduli = how many residues of each operand entry its conversion read.  Synchronises. */
int mpres_last_small_base(mpres_ctx *ctx, int *moduli, int *input_moduli);
/* 1 when the last fast-path mp_gemm rebuilt its exact sums in binary and rounded them once (full-precision inputs: the sums exceed
 * the number format but not the one-byte base; results then agree with the reference within its error model, src/arith/mul.cuh:108-110
 * rounds every product instead), 0 when it ran in the bit-exact window, < 0 on error. */
int mpres_last_binary_rounding(const mpres_ctx *ctx);
/* Workspace budget.  The fast mp_gemm keeps its planes (one-byte residues of A', B' and of the sums, shift planes, lists) in a per-context
 * pool that grows on demand: about (P + 2) k (m + n) + (P + 4 N + 10) m n bytes for P one-byte moduli, against the m x n scratch matrix the
 * reference needs (src/blas/gemm.cuh:98-139).  mpres_set_workspace_limit caps the pool (0 = no cap): a call whose reservation would
 * exceed the cap -- or that the device cannot satisfy -- is served in reference order instead (same results as MPRES_MODE_REFERENCE_ORDER,
 * one m x n scratch matrix) and counted by mpres_workspace_fallbacks; a sharded call returns cudaErrorMemoryAllocation (2) instead,
 * because its peers wait for this rank's planes.  mpres_workspace_bytes = bytes the pool holds now. */
int mpres_set_workspace_limit(mpres_ctx *ctx, size_t bytes);
long mpres_workspace_fallbacks(const mpres_ctx *ctx);
size_t mpres_workspace_bytes(const mpres_ctx *ctx);
/* the index-th one-byte modulus (0 when out of range or when the small base is unavailable for this moduli set) */
int mpres_small_modulus(const mpres_ctx *ctx, int index);
/* Test probe: copy `bytes` at `offset` of internal workspace `slot` to the host after synchronising the last stream
 * (slots of the fast mp_gemm path: 3/4 limb planes of A/B, 5 residue planes of the sums, 8/9 one-byte planes of A/B,
 * 10 one-byte planes of the sums).  Returns the bytes copied or < 0. */
long mpres_debug_read_workspace(mpres_ctx *ctx, int slot, size_t offset, void *host, size_t bytes);
/* Stage-3 kernel of the fast path: 0 = entry-per-thread normalisation with a residue-parallel list
 * kernel for the entries that need refinement / rounding / sign resolution (default), 1 = the
 * residue-parallel tile kernel for every entry, 2 = like 0 but with the generic 64-bit modular products
 * instead of the 32-bit Barrett step used when every modulus has the same bit length <= 27, 3 = like 0 but,
 * on the small-modulus path, with the base extension fused into the normalisation kernel (no residue planes
 * in memory; measured slightly slower on B200 because of its register / shared-memory footprint), 4 = like 0 but
 * reading the residues of the sums straight from global memory instead of staging them in shared memory (cp.async).
 * Identical results (including interval evaluations). */
int mpres_set_stage3_kernel(mpres_ctx *ctx, int kind);
/* Reduced-base fast path (default on): stages 1 and 2 run on the first n' moduli only, n' the smallest
 * multiple of four whose product exceeds four times the largest exact sum (from the per-row / per-column
 * magnitude windows); the remaining residues follow by mixed-radix base extension.  Results are identical
 * to the full-base path.  mpres_last_base_size returns n' of the last call (synchronises). */
int mpres_set_reduced_base(mpres_ctx *ctx, int on);
long mpres_last_base_size(mpres_ctx *ctx);
/* Stage-1 kernels: 0 = vectorised alignment (four residues per work item) and the (min,+) exponent product from
 * candidate lists (default), 1 = one-residue-per-thread alignment kernel and the dense (min,+) kernel, 2 = vectorised
 * alignment and the dense (min,+) kernel, 3 = like 0 but the small-modulus alignment computes its residues on the tensor
 * cores (mma.sync) instead of with byte dot products (dp4a); measured slower on B200.  Identical results. */
int mpres_set_stage1_kernel(mpres_ctx *ctx, int kind);
/* entries of the last fast-path call whose (min,+) value the candidate lists did not cover (recomputed densely; synchronises) */
long mpres_last_minplus_dense_count(mpres_ctx *ctx);
/* Tile configuration of the single-pass mp_gemv / mp_dot kernels: 0 = 8 columns (4 row tiles) per stage, 2-deep
 * ring (default); 1 = 4 columns (2 row tiles), 3-deep; 2 = 8 columns (4 row tiles), 3-deep.  Identical results. */
int mpres_set_vec_config(mpres_ctx *ctx, int cfg);
/* number of result elements of the last fast-path call that the entry-per-thread normalisation kernel
 * handed to the residue-parallel list kernel (diagnostic; synchronises) */
long mpres_last_slow_count(mpres_ctx *ctx);
/* number of result elements the last AUTO/FAST call routed to the reference-order fallback
 * (synchronises the stream of that call) */
long mpres_last_fallback_count(mpres_ctx *ctx);
/* kernels launched by this library since init */
long mpres_launch_count(const mpres_ctx *ctx);
/* Optional per-stage timing of the fast mp_gemm path with CUDA events on the caller's stream (what
 * the reference's tests do around whole calls with tests/timers.cuh:57-76).  ms[0] = stage 1
 * alignment, ms[1] = stage 2 per-modulus multiply-accumulate, ms[2] = stage 3 normalisation + epilogue. */
int mpres_set_profiling(mpres_ctx *ctx, int on);
int mpres_last_stage_ms(mpres_ctx *ctx, float *ms, int *stage2_launches);
/* with profiling on: "kernel=ms;kernel=ms;..." of the last fast-path mp_gemm call, CUDA events on the caller's stream after every kernel
 * (group) of the call; returns the bytes written or < 0.  Synchronises. */
long mpres_last_kernel_ms(mpres_ctx *ctx, char *out, size_t cap);

/* ---- containers: replace cuda::mp_array_init / clear / host2device / device2host
 *      (src/mparray.cuh:35,59,76,125) and the mp_collection_* twins (src/mpcollection.cuh:35,54,69,117).
 *      Host side is the reference's AoS mp_float_t[] (4N+40 bytes per element); copies are bulk. ---- */
int mpres_array_init(mpres_ctx *ctx, mpres_array_t *arr, size_t size);
int mpres_array_clear(mpres_ctx *ctx, mpres_array_t *arr);
int mpres_array_host2device(mpres_ctx *ctx, mpres_array_t *dst, const void *host_mp_float, size_t size);
int mpres_array_device2host(mpres_ctx *ctx, void *host_mp_float, const mpres_array_t *src, size_t size);
int mpres_collection_init(mpres_ctx *ctx, mpres_collection_t *arr, size_t size);
int mpres_collection_clear(mpres_ctx *ctx, mpres_collection_t *arr);
int mpres_collection_host2device(mpres_ctx *ctx, mpres_collection_t *dst, const void *host_mp_float, size_t size);
int mpres_collection_device2host(mpres_ctx *ctx, void *host_mp_float, const mpres_collection_t *src, size_t size);

/* Device-side conversion (replaces the host string round trip of mp_set_mpfr / mp_set_d,
 * src/arith/assign.cuh:54-127): element i of dst (from `offset`) := (-1)^sign[i] * L_i * 2^exp[i] with
 * L_i the little-endian nlimbs x 32-bit integer at limbs[i * nlimbs]; trailing zero bits are trimmed
 * and the interval evaluation computed like mp_set_mpfr does.  All pointers are device pointers. */
int mpres_array_set_binary(mpres_ctx *ctx, mpres_array_t *dst, size_t offset, const int *sign, const int *exp,
                           const uint32_t *limbs, int nlimbs, size_t count, mpres_stream_t stream);

/* ---- BLAS entry points ---------------------------------------------------------------------------- */

/* cuda::mp_gemm<blockDim1x, blockDim1y, gridDim2x, gridDim2y, blockDim3> (src/blas/gemm.cuh:69-70):
 * C = alpha * op(A) * op(B) + beta * C, column-major.  The launch-shape template parameters of the
 * reference have no equivalent here.  `buffer` (m*n scratch in the reference) may be NULL.  All four
 * transpose combinations are accepted (the reference prints and returns for anything but N/N,
 * gemm.cuh:114-125; semantics of src/blas/v2/gemm_v2.cuh:64-71). */
int mpres_gemm(mpres_ctx *ctx, int transa, int transb, int m, int n, int k, const mpres_array_t *alpha,
               const mpres_array_t *A, int lda, const mpres_array_t *B, int ldb, const mpres_array_t *beta,
               mpres_array_t *C, int ldc, mpres_array_t *buffer, mpres_stream_t stream);

/* cuda::mp_gemv<gridDim1, blockDim1, gridDim2, blockDim3> (src/blas/gemv.cuh:150-152):
 * y = alpha * op(A) * x + beta * y.  buffer1 / buffer2 may be NULL (no m*n intermediate is built). */
int mpres_gemv(mpres_ctx *ctx, int trans, int m, int n, const mpres_array_t *alpha, const mpres_array_t *A, int lda,
               const mpres_array_t *x, int incx, const mpres_array_t *beta, mpres_array_t *y, int incy,
               mpres_array_t *buffer1, mpres_array_t *buffer2, mpres_stream_t stream);

/* cuda::mp_dot<gridDim1, blockDim1, gridDim2, gridDim3, blockDim3> (src/blas/dot.cuh:84-85):
 * r[0] = sum x_i * y_i.  `buffer` (n scratch in the reference) may be NULL. */
int mpres_dot(mpres_ctx *ctx, int n, const mpres_array_t *x, int incx, const mpres_array_t *y, int incy,
              mpres_array_t *r, mpres_array_t *buffer, mpres_stream_t stream);

/* cuda::mp_scal<gridDim1, blockDim1, gridDim2> (src/blas/scal.cuh:45-62): x = round(alpha * x); n <= 0 or incx <= 0 return silently. */
int mpres_scal(mpres_ctx *ctx, int n, const mpres_array_t *alpha, mpres_array_t *x, int incx, mpres_stream_t stream);

/* cuda::mp_axpy<gridDim1, blockDim1, gridDim2> (src/blas/axpy.cuh:46-76): y = round(round(alpha * x) + y), one fused pass.
 * `buffer` (n scratch elements in the reference) may be NULL. */
int mpres_axpy(mpres_ctx *ctx, int n, const mpres_array_t *alpha, const mpres_array_t *x, int incx, mpres_array_t *y, int incy,
               mpres_array_t *buffer, mpres_stream_t stream);

/* cuda::mp_waxpby<gridDim1, blockDim1, gridDim2> (src/blas/waxpby.cuh:50-93): w = round(round(beta * y) + round(alpha * x)). */
int mpres_waxpby(mpres_ctx *ctx, int n, const mpres_array_t *alpha, const mpres_array_t *x, int incx, const mpres_array_t *beta,
                 const mpres_array_t *y, int incy, mpres_array_t *w, int incw, mpres_array_t *buffer, mpres_stream_t stream);

/* cuda::mp_ge_add<...> (src/blas/geadd.cuh:58-105): C = alpha * A + beta * B, and cuda::mp_ge_acc<...> (src/blas/geacc.cuh:57-97):
 * B = alpha * A + beta * B; m x n, column-major, non-transposed; products and sum rounded as in the reference. */
int mpres_ge_add(mpres_ctx *ctx, int m, int n, const mpres_array_t *alpha, const mpres_array_t *A, int lda, const mpres_array_t *beta,
                 const mpres_array_t *B, int ldb, mpres_array_t *C, int ldc, mpres_array_t *buffer, mpres_stream_t stream);
int mpres_ge_acc(mpres_ctx *ctx, int m, int n, const mpres_array_t *alpha, const mpres_array_t *A, int lda, const mpres_array_t *beta,
                 mpres_array_t *B, int ldb, mpres_array_t *buffer, mpres_stream_t stream);

/* cuda::mp_ger<...> (src/blas/ger.cuh:157-206): A = alpha * x * y^T + A (rank-1 update; alpha * y is rounded first, as in the reference). */
int mpres_ger(mpres_ctx *ctx, int m, int n, const mpres_array_t *alpha, const mpres_array_t *x, int incx, const mpres_array_t *y, int incy,
              mpres_array_t *A, int lda, mpres_array_t *buffer1, mpres_array_t *buffer2, mpres_stream_t stream);

/* cuda::mp_ge_diag_scale<...> (src/blas/gediagscale.cuh:54-99): A = round(A D) (MPRES_RIGHT_SIDE, D of n elements) or round(D A)
 * (MPRES_LEFT_SIDE, D of m elements), D a diagonal matrix stored as a vector with increment incd (negative increments in the BLAS
 * convention). Where the reference returns silently on incd == 0 or lda < max(1, m), this returns -3 and writes nothing. */
int mpres_ge_diag_scale(mpres_ctx *ctx, int side, int m, int n, const mpres_array_t *D, int incd, mpres_array_t *A, int lda, mpres_stream_t stream);

/* cuda::mp_ge_lr_scale<...> (src/blas/gelrscale.cuh:56-93): A = round(round(DL A) DR). */
int mpres_ge_lr_scale(mpres_ctx *ctx, int m, int n, const mpres_array_t *DL, int incdl, const mpres_array_t *DR, int incdr, mpres_array_t *A, int lda,
                      mpres_stream_t stream);

/* cuda::mp_rot<gridDim1, blockDim1, gridDim2> (src/blas/rot.cuh:49-100): x = round(round(c x) + round(s y)), y = round(round(c y) - round(s x)).
 * Unit increments: one fused pass, the buffers (n elements each in the reference) are not touched and may be NULL.  Other increments
 * follow the reference's kernel sequence step for step (its mp_scal(n, c, x, 1) calls included) and need both buffers. */
int mpres_rot(mpres_ctx *ctx, int n, mpres_array_t *x, int incx, mpres_array_t *y, int incy, const mpres_array_t *c, const mpres_array_t *s,
              mpres_array_t *buffer1, mpres_array_t *buffer2, mpres_stream_t stream);

/* cuda::mp_axpy_dot<gridDim1, blockDim1, gridDim2, gridDim3, blockDim3> (src/blas/axpydot.cuh:33-68): w = round(w - round(alpha v)) in one
 * fused pass, then r[0] = u^T w through mpres_dot (same result contract as mpres_dot). `buffer` may be NULL. */
int mpres_axpy_dot(mpres_ctx *ctx, int n, const mpres_array_t *alpha, mpres_array_t *w, int incw, const mpres_array_t *v, int incv,
                   const mpres_array_t *u, int incu, mpres_array_t *r, mpres_array_t *buffer, mpres_stream_t stream);

/* mp_gemm over HOST operands: the reference caller's mp_array_host2device x 3 + cuda::mp_gemm + mp_array_device2host sequence
 * (tests/blas/test_gemm.cu) as ONE call that pipelines the PCIe transfers with the compute by column panels of B and C (both bus
 * directions busy at once; the device arrays live in the context's workspace).  alpha, beta: one mp_float_t record each; A, B, Cin,
 * Cout: column-major AoS mp_float_t[] with the leading dimensions of mpres_gemm (Cout may alias Cin; rows m..ldc-1 of Cout receive
 * Cin's records).  Pinned host memory (cudaHostAlloc / cudaHostRegister) is needed for the overlap, pageable memory works unpipelined.
 * panels = 0 picks the panel count (<= 8, whole multiples of 256 columns); transposed B runs as one panel.  Synchronous: the result
 * is in Cout on return.  The fallback / base counters afterwards describe the last panel. */
int mpres_gemm_host(mpres_ctx *ctx, int transa, int transb, int m, int n, int k, const void *alpha, const void *A, int lda, const void *B, int ldb,
                    const void *beta, const void *Cin, void *Cout, int ldc, int panels);

/* mpres_gemm_host with B already resident on the device (an mp_array_t: e.g. each rank of a row-sharded multi-GPU GEMM uploads 1/N of B
 * and gathers the rest over NVLink instead of pulling all of B through its own PCIe link).  B must be complete before the call (the
 * transfers run on the library's own streams: synchronise the stream that produced B first).  Panels also apply to transposed B. */
/* Residues per entry of A / B the last mpres_gemm_host[_bdev] call moved to the device: the host cores cut the operand records down to what the
 * fast path reads (first n_in residues, sign, exponent, upper interval bound: 4 n_in + 24 of 4N + 40 bytes) before they cross the PCIe link; the
 * count is guessed from a sample, checked against every record while packing and against the call's own choice on the device, and anything that does
 * not fit goes up in full.  0: full records (reference-order mode, formats without the one-byte base, MPRES_HOST_LEAN=0, or a fallback). */
int mpres_last_host_upload_residues(const mpres_ctx *ctx);
int mpres_gemm_host_bdev(mpres_ctx *ctx, int transa, int transb, int m, int n, int k, const void *alpha, const void *A, int lda, const mpres_array_t *B,
                         int ldb, const void *beta, const void *Cin, void *Cout, int ldc, int panels);

/* ---- row-sharded mp_gemm over the GPUs of one NVLink domain (BASELINE config 3; the reference has no multi-GPU code) --------------
 * One rank per GPU (a process each, or threads of one process).  Rank r owns rows [m r / G, m (r + 1) / G) of A and C as compact
 * arrays; every rank holds B.  Inside a call each rank converts ONE column block of B (n / G columns) into the one-byte planes of the
 * small-modulus stage 2 and copies that package into every peer's receive buffer over NVLink (copy engines, peer-mapped memory: CUDA
 * IPC between processes), while its tensor kernel already multiplies the panels that have arrived -- the alignment of B, which does
 * not shrink with G when every rank converts all of B, is done once per node.  Results are identical to mpres_gemm on the row block.
 *   mpres_shard_create   allocates the rank's receive buffer for calls with this n and k <= k_max (n % world == 0)
 *   mpres_shard_export   writes mpres_shard_handle_size() bytes to publish to the other ranks (all-gather them in rank order)
 *   mpres_shard_connect  maps the peers' buffers from the gathered handles (world x handle size bytes)
 *   mpres_gemm_sharded   collective: every rank calls it with its row block (m_local rows) and its complete copy of B; the calls of
 *                        all ranks must be made in the same order.  MPRES_PUSH_STREAMS (environment) = copy streams per rank (2). */
typedef struct mpres_shard mpres_shard;
size_t mpres_shard_handle_size(void);
int mpres_shard_create(mpres_ctx *ctx, int rank, int world, int n, int k_max, mpres_shard **out);
int mpres_shard_export(mpres_shard *s, void *handle_out);
int mpres_shard_connect(mpres_shard *s, const void *handles);
int mpres_shard_destroy(mpres_shard *s);
int mpres_gemm_sharded(mpres_shard *s, int transa, int transb, int m_local, int n, int k, const mpres_array_t *alpha, const mpres_array_t *A, int lda,
                       const mpres_array_t *B, int ldb, const mpres_array_t *beta, mpres_array_t *C, int ldc, mpres_stream_t stream);

/* ---- sums of magnitudes and norms -----------------------------------------------------------------------
 * mpres_asum     r[0] = |x_0| + ... + |x_(n-1)|                                   cuda::mp_asum<gridDim1, blockDim1>(n, x, incx, r), src/blas/asum.cuh:41
 * mpres_norm     norm = MPRES_ONE_NORM: the same; MPRES_INF_NORM: max |x_i|        cuda::mp_norm<...>(norm, n, x, incx, r), src/blas/norm.cuh:43
 * mpres_ge_norm  one norm (largest column sum) / infinity norm (largest row sum)  cuda::mp_ge_norm<...>(norm, m, n, A, lda, r, buffer), src/blas/genorm.cuh:142
 * n <= 0 or incx <= 0 return without touching r, as the reference does.  The sums run on the exact-window accumulators of mp_dot (one
 * pass over the data, one rounding); whenever the reference's own additions do not round (its p/4-bit benchmark inputs) the digits,
 * sign and exponent are the reference's.  The maximum is the element itself with its sign cleared (src/mpreduct.cuh:247-250); among
 * entries of equal magnitude any one of them may be returned.  buffer (n or m elements) may be NULL. */
#define MPRES_ONE_NORM 171   /* mblas_one_norm, src/blas/mblas_enum.cuh:41-44 */
#define MPRES_INF_NORM 175   /* mblas_inf_norm */
int mpres_asum(mpres_ctx *ctx, int n, const mpres_array_t *x, int incx, mpres_array_t *r, mpres_stream_t stream);
int mpres_norm(mpres_ctx *ctx, int norm, int n, const mpres_array_t *x, int incx, mpres_array_t *r, mpres_stream_t stream);
int mpres_ge_norm(mpres_ctx *ctx, int norm, int m, int n, const mpres_array_t *A, int lda, mpres_array_t *r, mpres_array_t *buffer, mpres_stream_t stream);

/* ---- sparse matrix-vector product with multiple-precision entries ------------------------------------------
 * y = A x, A in CSR (irp[m + 1], ja[nnz], as[nnz]) or ELLPACK (column-major m x maxnzr arrays ja / as, padding marked by ja < 0), the
 * entries of A in an mp_collection_t:  cuda::mp_spmv_mpmtx_csr2st<...>(m, n, nnz, irp, ja, as, x, y, buffer), src/sparse/mpmtx/spmv_mpmtx_csr2st.cuh:106;
 * cuda::mp_spmv_mpmtx_ell2st<...>(m, n, maxnzr, ja, as, x, y, buffer), src/sparse/mpmtx/spmv_mpmtx_ell2st.cuh:119.  One pass: a lane group
 * per row forms round(a x_j) and adds it to the running sum in the reference's order -- the same mp_mul / mp_add sequence, so the same
 * records; `buffer` (the reference's nnz-element scratch) is accepted and unused, may be NULL.  irp, ja: device pointers. */
int mpres_spmv_csr2st(mpres_ctx *ctx, int m, int n, int nnz, const int *irp, const int *ja, const mpres_collection_t *as, const mpres_array_t *x,
                      mpres_array_t *y, mpres_collection_t *buffer, mpres_stream_t stream);
int mpres_spmv_ell2st(mpres_ctx *ctx, int m, int n, int maxnzr, const int *ja, const mpres_collection_t *as, const mpres_array_t *x, mpres_array_t *y,
                      mpres_collection_t *buffer, mpres_stream_t stream);

/* ---- conversions either side of the path ---------------------------------------------------------------------
 * mpres_array_set_d: dst[offset + i] = src[i] exactly (sign, exponent, odd significand reduced modulo every m_q: digits, sign and exponent
 * as mp_set_d, src/arith/assign.cuh:54-81; interval evaluation by the device rns_eval_compute).  mpres_array_get_d: dst[i] = the double
 * nearest to src[offset + i] (ties to even; mp_get_d, src/arith/assign.cuh:154-180, rounds the exact value the same way through MPFR).
 * src / dst doubles are DEVICE pointers. */
int mpres_array_set_d(mpres_ctx *ctx, mpres_array_t *dst, size_t offset, const double *src, size_t n, mpres_stream_t stream);
/* r[0] = x[0] / y[0] rounded to nearest (ties to even) at MP_PRECISION bits, on the device.  The reference's cuda::mp_div copies both numbers
 * to the host and divides with MPFR_RNDN at MP_PRECISION bits (src/arith/div.cuh:33-66): the same value.  A zero divisor leaves r untouched. */
int mpres_div(mpres_ctx *ctx, mpres_array_t *r, const mpres_array_t *x, const mpres_array_t *y, mpres_stream_t stream);
/* Conjugate gradients for A x = b, A symmetric positive definite in CSR with double entries (mp_cg_csr, src/sparse/solver/cg_csr.cuh:53; with
 * M != NULL the diagonally preconditioned iteration mp_pcg_csr, pcg_csr.cuh:57, M = n doubles).  The iteration is the reference's, operation for
 * operation, in this library's kernels: the matrix is converted once to multiple precision (exactly) and multiplied by the two-stage SpMV,
 * mpres_dot, mpres_div, fused a x + y.  irp, ja, as, M: device pointers; b, x: n elements (x: initial guess in, solution out); tol: relative
 * residual; *iters: iterations done; resvec: maxit + 1 doubles on the host (relative residual history) or NULL.  Synchronises. */
int mpres_cg_csr(mpres_ctx *ctx, int n, int nnz, const int *irp, const int *ja, const double *as, const mpres_array_t *b, double tol, int maxit,
                 const double *M, mpres_array_t *x, int *iters, double *resvec, mpres_stream_t stream);
int mpres_array_get_d(mpres_ctx *ctx, double *dst, const mpres_array_t *src, size_t offset, size_t n, mpres_stream_t stream);

/* The same three operations over mp_collection_t operands with explicit allocated lengths (the
 * reference only uses mp_collection_t in its sparse kernels, src/sparse/mpmtx/*.cuh; north_star asks
 * for the dense path over both containers). */
int mpres_gemm_coll(mpres_ctx *ctx, int transa, int transb, int m, int n, int k,
                    const mpres_collection_t *alpha, const mpres_collection_t *A, int lda, size_t lenA,
                    const mpres_collection_t *B, int ldb, size_t lenB, const mpres_collection_t *beta,
                    mpres_collection_t *C, int ldc, size_t lenC, mpres_stream_t stream);
int mpres_gemv_coll(mpres_ctx *ctx, int trans, int m, int n, const mpres_collection_t *alpha,
                    const mpres_collection_t *A, int lda, size_t lenA, const mpres_collection_t *x, int incx, size_t lenx,
                    const mpres_collection_t *beta, mpres_collection_t *y, int incy, size_t leny, mpres_stream_t stream);
int mpres_dot_coll(mpres_ctx *ctx, int n, const mpres_collection_t *x, int incx, size_t lenx,
                   const mpres_collection_t *y, int incy, size_t leny, mpres_collection_t *r, mpres_stream_t stream);

/* Partial DOT for segment sharding across GPUs: reduces x[0..n), y[0..n) to ONE packed
 * mp_float_t (4N+40 bytes, AoS) at `partial` (device); mpres_reduce_partials sums `count` packed
 * partials in index order (mp_add, src/arith/add.cuh:190-200) into r[0] -- run it on every rank after
 * an all-gather of the partial bytes. */
int mpres_dot_partial(mpres_ctx *ctx, int n, const mpres_array_t *x, int incx, const mpres_array_t *y, int incy,
                      void *partial, mpres_stream_t stream);
int mpres_reduce_partials(mpres_ctx *ctx, const void *partials, int count, mpres_array_t *r, mpres_stream_t stream);

/* Element-wise probes of the scalar device routines, for parity tests against the reference's
 * cuda::mp_mul / mp_add / rns_eval_compute[_fast] / mp_round (src/arith/mul.cuh:99-111,
 * add.cuh:190-200, rns.cuh:797-933, arith_utils.cuh:184-193).  op: 0 mul, 1 add, 2 eval, 3 eval_fast,
 * 4 round by bits[i].  x, y, r: device AoS mp_float_t arrays; bits: device ints. */
int mpres_probe(mpres_ctx *ctx, int op, void *r, const void *x, const void *y, const int *bits, size_t n,
                mpres_stream_t stream);

const char *mpres_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MPRES_B200_H */
