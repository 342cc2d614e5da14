/*
 * mpres_compat.cuh -- source-level shim: the reference's own names and signatures for the
 * mp_gemm / mp_gemv / mp_dot path, forwarding to the C-ABI of libmpres_b200.so.
 *
 * A translation unit of the reference that did
 *     #include "blas/gemm.cuh"   #include "blas/gemv.cuh"   #include "blas/dot.cuh"   #include "mparray.cuh"
 * includes this header instead (after its params.h, which supplies RNS_MODULI_SIZE / RNS_MODULI_VALUES) and links
 * -lmpres_b200.  Everything else in the caller stays as it is: rns_const_init(); mp_const_init();
 * cuda::mp_array_init(...); cuda::mp_array_host2device(...); cuda::mp_gemm<...>(...).
 *
 *   reference                                                   here
 *   rns_const_init()/mp_const_init()  src/rns.cuh:324, arith_utils.cuh:44   -> mpres_init_moduli (compiled-in moduli)
 *   cuda::mp_array_init/clear/host2device/device2host  src/mparray.cuh:35-165 -> mpres_array_*
 *   cuda::mp_gemm<bx,by,g2x,g2y,b3>   src/blas/gemm.cuh:69-70               -> mpres_gemm
 *   cuda::mp_gemv<g1,b1,g2,b3>        src/blas/gemv.cuh:150-152             -> mpres_gemv
 *   cuda::mp_dot<g1,b1,g2,g3,b3>      src/blas/dot.cuh:84-85                -> mpres_dot
 * The launch-shape template parameters are accepted and ignored.  Like the reference, the functions return void
 * and report nothing on bad arguments; the C-ABI status is available through mpres_compat_last_status().
 */
#ifndef MPRES_COMPAT_CUH
#define MPRES_COMPAT_CUH

#include <cstddef>
#include <cstdio>

#include "mpres_b200.h"

#ifndef RNS_MODULI_SIZE
#error "include the reference's params.h (RNS_MODULI_SIZE, RNS_MODULI_VALUES) before mpres_compat.cuh"
#endif

/* the reference's type names (src/types.cuh) on top of the layout-identical C-ABI structs */
typedef mpres_er_float_t er_float_t;
typedef er_float_t *er_float_ptr;
typedef struct {
    int digits[RNS_MODULI_SIZE];
    int sign;
    int exp;
    er_float_t eval[2];
} mp_float_t;
typedef mp_float_t *mp_float_ptr;
typedef mpres_array_t mp_array_t;
typedef mpres_collection_t mp_collection_t;

enum mblas_trans_type { mblas_no_trans = 111, mblas_trans = 112, mblas_conj_trans = 113 }; /* src/blas/mblas_enum.cuh:25-29 */
enum mblas_side_type { mblas_left_side = 141, mblas_right_side = 142 };                    /* src/blas/mblas_enum.cuh:37-40 */
enum mblas_norm_type { mblas_one_norm = 171, mblas_inf_norm = 175 };                       /* src/blas/mblas_enum.cuh:41-44 */

namespace mpres_compat {
inline mpres_ctx *&ctx() { static mpres_ctx *c = nullptr; return c; }
inline int &status() { static int s = 0; return s; }
}  // namespace mpres_compat

inline int mpres_compat_last_status() { return mpres_compat::status(); }

/* src/rns.cuh:324-442: all RNS constants for the compiled-in moduli set, uploaded to the current device */
inline void rns_const_init() {
    static const int moduli[RNS_MODULI_SIZE] = RNS_MODULI_VALUES;
    int dev = 0;
#ifdef __CUDACC__
    cudaGetDevice(&dev);
#endif
    if (mpres_compat::ctx()) mpres_finalize(mpres_compat::ctx());
    mpres_compat::ctx() = nullptr;
    mpres_compat::status() = mpres_init_moduli(&mpres_compat::ctx(), moduli, RNS_MODULI_SIZE, dev);
    if (mpres_compat::status() != 0) fprintf(stderr, "mpres_b200: rns_const_init failed (%d)\n", mpres_compat::status());
}
/* src/arith/arith_utils.cuh:44-85: MP_PRECISION, MP_H, MP_J are part of the same context */
inline void mp_const_init() {}
#define MP_PRECISION (mpres_precision(mpres_compat::ctx()))
#define MP_H (mpres_mp_h(mpres_compat::ctx()))
#define MP_J (mpres_mp_j(mpres_compat::ctx()))

namespace cuda {

inline void mp_array_init(mp_array_t &dev_dest, size_t size) { mpres_compat::status() = mpres_array_init(mpres_compat::ctx(), &dev_dest, size); }
inline void mp_array_clear(mp_array_t &dev_dest) { mpres_compat::status() = mpres_array_clear(mpres_compat::ctx(), &dev_dest); }
inline void mp_array_host2device(mp_array_t &dev_dest, mp_float_ptr host_src, size_t size) {
    mpres_compat::status() = mpres_array_host2device(mpres_compat::ctx(), &dev_dest, host_src, size);
}
inline void mp_array_device2host(mp_float_ptr host_dest, mp_array_t &dev_src, size_t size) {
    mpres_compat::status() = mpres_array_device2host(mpres_compat::ctx(), host_dest, &dev_src, size);
}
inline void mp_collection_init(mp_collection_t &dev_dest, size_t size) { mpres_compat::status() = mpres_collection_init(mpres_compat::ctx(), &dev_dest, size); }
inline void mp_collection_clear(mp_collection_t &dev_dest) { mpres_compat::status() = mpres_collection_clear(mpres_compat::ctx(), &dev_dest); }
inline void mp_collection_host2device(mp_collection_t &dev_dest, mp_float_ptr host_src, size_t size) {
    mpres_compat::status() = mpres_collection_host2device(mpres_compat::ctx(), &dev_dest, host_src, size);
}
inline void mp_collection_device2host(mp_float_ptr host_dest, mp_collection_t &dev_src, size_t size) {
    mpres_compat::status() = mpres_collection_device2host(mpres_compat::ctx(), host_dest, &dev_src, size);
}

template <int blockDim1x, int blockDim1y, int gridDim2x, int gridDim2y, int blockDim3>
void mp_gemm(enum mblas_trans_type transa, enum mblas_trans_type transb, const int m, const int n, const int k, mp_array_t &alpha,
             mp_array_t &A, const int lda, mp_array_t &B, const int ldb, mp_array_t &beta, mp_array_t &C, const int ldc, mp_array_t &buffer) {
    mpres_compat::status() = mpres_gemm(mpres_compat::ctx(), transa, transb, m, n, k, &alpha, &A, lda, &B, ldb, &beta, &C, ldc, &buffer, nullptr);
}

template <int gridDim1, int blockDim1, int gridDim2, int blockDim3>
void mp_gemv(enum mblas_trans_type trans, const int m, const int n, mp_array_t &alpha, mp_array_t &A, const int lda, mp_array_t &x,
             const int incx, mp_array_t &beta, mp_array_t &y, const int incy, mp_array_t &buffer1, mp_array_t &buffer2) {
    mpres_compat::status() = mpres_gemv(mpres_compat::ctx(), trans, m, n, &alpha, &A, lda, &x, incx, &beta, &y, incy, &buffer1, &buffer2, nullptr);
}

template <int gridDim1, int blockDim1, int gridDim2, int gridDim3, int blockDim3>
void mp_dot(const int n, mp_array_t &x, const int incx, mp_array_t &y, const int incy, mp_array_t &r, mp_array_t &buffer) {
    mpres_compat::status() = mpres_dot(mpres_compat::ctx(), n, &x, incx, &y, incy, &r, &buffer, nullptr);
}

/* src/blas/scal.cuh:45-46, src/blas/axpy.cuh:46 */
template <int gridDim1, int blockDim1, int gridDim2>
void mp_scal(const int n, mp_array_t &alpha, mp_array_t &x, const int incx) {
    mpres_compat::status() = mpres_scal(mpres_compat::ctx(), n, &alpha, &x, incx, nullptr);
}
template <int gridDim1, int blockDim1, int gridDim2>
void mp_axpy(const int n, mp_array_t &alpha, mp_array_t &x, const int incx, mp_array_t &y, const int incy, mp_array_t &buffer) {
    mpres_compat::status() = mpres_axpy(mpres_compat::ctx(), n, &alpha, &x, incx, &y, incy, &buffer, nullptr);
}

/* src/blas/waxpby.cuh:50, geadd.cuh:58, geacc.cuh:57, ger.cuh:157 */
template <int gridDim1, int blockDim1, int gridDim2>
void mp_waxpby(const int n, mp_array_t &alpha, mp_array_t &x, const int incx, mp_array_t &beta, mp_array_t &y, const int incy, mp_array_t &w,
               const int incw, mp_array_t &buffer) {
    mpres_compat::status() = mpres_waxpby(mpres_compat::ctx(), n, &alpha, &x, incx, &beta, &y, incy, &w, incw, &buffer, nullptr);
}
template <int blockDim1x, int blockDim1y, int gridDim2x, int gridDim2y>
void mp_ge_add(const int m, const int n, mp_array_t &alpha, mp_array_t &A, const int lda, mp_array_t &beta, mp_array_t &B, const int ldb, mp_array_t &C,
               const int ldc, mp_array_t &buffer) {
    mpres_compat::status() = mpres_ge_add(mpres_compat::ctx(), m, n, &alpha, &A, lda, &beta, &B, ldb, &C, ldc, &buffer, nullptr);
}
template <int blockDim1x, int blockDim1y, int gridDim2x, int gridDim2y>
void mp_ge_acc(const int m, const int n, mp_array_t &alpha, mp_array_t &A, const int lda, mp_array_t &beta, mp_array_t &B, const int ldb, mp_array_t &buffer) {
    mpres_compat::status() = mpres_ge_acc(mpres_compat::ctx(), m, n, &alpha, &A, lda, &beta, &B, ldb, &buffer, nullptr);
}
template <int blockDim1x, int blockDim1y, int gridDim2x, int gridDim2y>
void mp_ger(const int m, const int n, mp_array_t &alpha, mp_array_t &x, const int incx, mp_array_t &y, const int incy, mp_array_t &A, const int lda,
            mp_array_t &buffer1, mp_array_t &buffer2) {
    mpres_compat::status() = mpres_ger(mpres_compat::ctx(), m, n, &alpha, &x, incx, &y, incy, &A, lda, &buffer1, &buffer2, nullptr);
}

/* src/blas/gediagscale.cuh:53-54, gelrscale.cuh:55-56, rot.cuh:48-49 */
template <int gridDim1, int blockDim1, int gridDim2>
void mp_ge_diag_scale(enum mblas_side_type side, const int m, const int n, mp_array_t &D, const int incd, mp_array_t &A, const int lda) {
    mpres_compat::status() = mpres_ge_diag_scale(mpres_compat::ctx(), side, m, n, &D, incd, &A, lda, nullptr);
}
template <int gridDim1, int blockDim1, int gridDim2>
void mp_ge_lr_scale(const int m, const int n, mp_array_t &DL, const int incdl, mp_array_t &DR, const int incdr, mp_array_t &A, const int lda) {
    mpres_compat::status() = mpres_ge_lr_scale(mpres_compat::ctx(), m, n, &DL, incdl, &DR, incdr, &A, lda, nullptr);
}
template <int gridDim1, int blockDim1, int gridDim2>
void mp_rot(const int n, mp_array_t &x, const int incx, mp_array_t &y, const int incy, mp_array_t &c, mp_array_t &s, mp_array_t &buffer1,
            mp_array_t &buffer2) {
    mpres_compat::status() = mpres_rot(mpres_compat::ctx(), n, &x, incx, &y, incy, &c, &s, &buffer1, &buffer2, nullptr);
}


/* src/blas/axpydot.cuh:32-33 */
template <int gridDim1, int blockDim1, int gridDim2, int gridDim3, int blockDim3>
void mp_axpy_dot(const int n, mp_array_t &alpha, mp_array_t &w, const int incw, mp_array_t &v, const int incv, mp_array_t &u, const int incu,
                 mp_array_t &r, mp_array_t &buffer) {
    mpres_compat::status() = mpres_axpy_dot(mpres_compat::ctx(), n, &alpha, &w, incw, &v, incv, &u, incu, &r, &buffer, nullptr);
}

/* src/blas/asum.cuh:40-41, norm.cuh:42-43, genorm.cuh:141-142 */
template <int gridDim1, int blockDim1>
void mp_asum(const int n, mp_array_t &x, const int incx, mp_array_t &r) {
    mpres_compat::status() = mpres_asum(mpres_compat::ctx(), n, &x, incx, &r, nullptr);
}
template <int gridDim1, int blockDim1>
void mp_norm(enum mblas_norm_type norm, const int n, mp_array_t &x, const int incx, mp_array_t &r) {
    mpres_compat::status() = mpres_norm(mpres_compat::ctx(), norm, n, &x, incx, &r, nullptr);
}
template <int gridDim1, int blockDim1>
void mp_ge_norm(enum mblas_norm_type norm, const int m, const int n, mp_array_t &A, const int lda, mp_array_t &r, mp_array_t &buffer) {
    mpres_compat::status() = mpres_ge_norm(mpres_compat::ctx(), norm, m, n, &A, lda, &r, &buffer, nullptr);
}

/* src/sparse/mpmtx/spmv_mpmtx_csr2st.cuh:105-106, spmv_mpmtx_ell2st.cuh:118-119 */
template <int gridDim1, int blockDim1, int gridDim2, int blockDim3>
void mp_spmv_mpmtx_csr2st(const int m, const int n, const int nnz, const int *irp, const int *ja, mp_collection_t &as, mp_array_t &x, mp_array_t &y,
                          mp_collection_t &buffer) {
    mpres_compat::status() = mpres_spmv_csr2st(mpres_compat::ctx(), m, n, nnz, irp, ja, &as, &x, &y, &buffer, nullptr);
}
template <int gridDim1, int blockDim1, int gridDim2, int blockDim3>
void mp_spmv_mpmtx_ell2st(const int m, const int n, const int maxnzr, const int *ja, mp_collection_t &as, mp_array_t &x, mp_array_t &y, mp_collection_t &buffer) {
    mpres_compat::status() = mpres_spmv_ell2st(mpres_compat::ctx(), m, n, maxnzr, ja, &as, &x, &y, &buffer, nullptr);
}

}  // namespace cuda

#endif /* MPRES_COMPAT_CUH */
