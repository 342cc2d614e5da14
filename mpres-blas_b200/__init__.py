"""mpres-blas_b200 -- host-side mirror of the reference's operator interface for the
mp_gemm / mp_gemv / mp_dot path, over the C-ABI of libmpres_b200.so (include/mpres_b200.h).

The names follow the reference (cuda::mp_array_init, mp_array_host2device, mp_gemm, mp_gemv, mp_dot:
src/mparray.cuh, src/blas/{gemm,gemv,dot}.cuh).  Host data is the reference's AoS mp_float_t[] as a
numpy structured array (record_dtype).  There is NO CPU implementation behind these calls: if the
CUDA library is missing or no GPU is present they raise.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MPRES_B200_LIB") or os.path.join(_HERE, "libmpres_b200.so")   # the override serves A/B builds of the same library

mblas_no_trans, mblas_trans, mblas_conj_trans = 111, 112, 113  # src/blas/mblas_enum.cuh:25-29
MODE_AUTO, MODE_REFERENCE_ORDER, MODE_FAST = 0, 1, 2
STAGE2_UMMA, STAGE2_UMMA_UNSTACKED, STAGE2_MMA_SYNC, STAGE2_SMALL, STAGE2_SMALL_TILED, STAGE2_SMALL_K64, STAGE2_SMALL_T128 = 0, 1, 2, 3, 4, 5, 6

_lib = None


class MpresError(RuntimeError):
    pass


class mp_array_t(ctypes.Structure):  # src/types.cuh:85-92
    _fields_ = [("digits", ctypes.c_void_p), ("sign", ctypes.c_void_p), ("exp", ctypes.c_void_p),
                ("eval", ctypes.c_void_p), ("buf", ctypes.c_void_p), ("len", ctypes.c_void_p)]


class mp_collection_t(ctypes.Structure):  # src/types.cuh:99-104
    _fields_ = [("digits", ctypes.c_void_p), ("sign", ctypes.c_void_p), ("exp", ctypes.c_void_p),
                ("eval", ctypes.c_void_p)]


EXPORTS = [
    "mpres_init", "mpres_init_moduli", "mpres_finalize", "mpres_moduli_size", "mpres_moduli_product_log2",
    "mpres_precision", "mpres_mp_h", "mpres_mp_j", "mpres_device", "mpres_sizeof_mp_float", "mpres_get_constant",
    "mpres_set_mode", "mpres_get_mode", "mpres_set_stage2_kernel", "mpres_set_stage3_kernel", "mpres_set_stage1_kernel", "mpres_set_reduced_base", "mpres_last_base_size", "mpres_last_slow_count", "mpres_last_fallback_count", "mpres_launch_count",
    "mpres_set_profiling", "mpres_last_stage_ms", "mpres_set_vec_config", "mpres_last_small_base", "mpres_last_binary_rounding", "mpres_last_host_upload_residues", "mpres_set_workspace_limit", "mpres_workspace_fallbacks", "mpres_workspace_bytes", "mpres_small_modulus", "mpres_debug_read_workspace", "mpres_last_minplus_dense_count",
    "mpres_array_init", "mpres_array_clear", "mpres_array_host2device", "mpres_array_device2host",
    "mpres_collection_init", "mpres_collection_clear", "mpres_collection_host2device", "mpres_collection_device2host",
    "mpres_array_set_binary", "mpres_gemm", "mpres_gemv", "mpres_dot", "mpres_scal", "mpres_axpy", "mpres_waxpby", "mpres_ge_add", "mpres_ge_acc", "mpres_ger", "mpres_ge_diag_scale", "mpres_ge_lr_scale", "mpres_rot", "mpres_axpy_dot", "mpres_gemm_host", "mpres_gemm_host_bdev", "mpres_gemm_coll", "mpres_gemv_coll",
    "mpres_dot_coll", "mpres_dot_partial", "mpres_reduce_partials", "mpres_probe", "mpres_version",
    "mpres_last_kernel_ms", "mpres_shard_handle_size", "mpres_shard_create", "mpres_shard_export", "mpres_shard_connect", "mpres_shard_destroy",
    "mpres_gemm_sharded", "mpres_asum", "mpres_norm", "mpres_ge_norm", "mpres_spmv_csr2st", "mpres_spmv_ell2st", "mpres_array_set_d", "mpres_array_get_d", "mpres_div", "mpres_cg_csr",
]


def load_library():
    """dlopen libmpres_b200.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MpresError("libmpres_b200.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                         "there is no CPU fallback")
    lib = ctypes.CDLL(LIB_PATH)
    lib.mpres_version.restype = ctypes.c_char_p
    lib.mpres_sizeof_mp_float.restype = ctypes.c_size_t
    lib.mpres_shard_handle_size.restype = ctypes.c_size_t
    for f in ("mpres_last_kernel_ms", "mpres_get_constant", "mpres_last_fallback_count", "mpres_launch_count", "mpres_last_slow_count", "mpres_last_base_size", "mpres_debug_read_workspace", "mpres_last_minplus_dense_count", "mpres_workspace_fallbacks"):
        getattr(lib, f).restype = ctypes.c_long
    lib.mpres_workspace_bytes.restype = ctypes.c_size_t
    lib.mpres_set_workspace_limit.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
    _lib = lib
    return lib


def record_dtype(N):
    """numpy view of the reference's AoS mp_float_t (src/types.cuh:69-74)."""
    return np.dtype([("digits", np.int32, (N,)), ("sign", np.int32), ("exp", np.int32),
                     ("eval", [("frac", np.float64), ("exp", np.int64)], (2,))])


def _check(rc, what):
    if rc != 0:
        kind = "invalid argument" if rc < 0 else "CUDA error"
        raise MpresError("%s failed: %s %d" % (what, kind, rc))


def _vp(x):
    return ctypes.c_void_p(x) if x is not None else None


class Context:
    """rns_const_init() + mp_const_init() for one moduli set on one device (src/rns.cuh:324,
    src/arith/arith_utils.cuh:44)."""

    def __init__(self, moduli_size=8, device=0, moduli=None):
        self.lib = load_library()
        self.h = ctypes.c_void_p()
        if moduli is None:
            _check(self.lib.mpres_init(ctypes.byref(self.h), int(moduli_size), int(device)), "mpres_init")
        else:
            arr = (ctypes.c_int * len(moduli))(*moduli)
            _check(self.lib.mpres_init_moduli(ctypes.byref(self.h), arr, len(moduli), int(device)), "mpres_init_moduli")
        self.N = self.lib.mpres_moduli_size(self.h)
        self.dtype = record_dtype(self.N)
        self.device = device

    def close(self):
        if self.h:
            self.lib.mpres_finalize(self.h)
            self.h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    precision = property(lambda s: s.lib.mpres_precision(s.h))
    mp_h = property(lambda s: s.lib.mpres_mp_h(s.h))
    mp_j = property(lambda s: s.lib.mpres_mp_j(s.h))
    log2M = property(lambda s: s.lib.mpres_moduli_product_log2(s.h))
    launch_count = property(lambda s: s.lib.mpres_launch_count(s.h))

    def set_mode(self, mode):
        _check(self.lib.mpres_set_mode(self.h, mode), "mpres_set_mode")

    def set_stage2_kernel(self, kind):
        _check(self.lib.mpres_set_stage2_kernel(self.h, kind), "mpres_set_stage2_kernel")

    def set_reduced_base(self, on):
        _check(self.lib.mpres_set_reduced_base(self.h, 1 if on else 0), "mpres_set_reduced_base")

    def last_base_size(self):
        return self.lib.mpres_last_base_size(self.h)

    def last_small_base(self):
        """(one-byte moduli, residues read per operand entry) of the last fast-path call; (0, 0) if it ran on the format's moduli"""
        a, b = ctypes.c_int(), ctypes.c_int()
        _check(self.lib.mpres_last_small_base(self.h, ctypes.byref(a), ctypes.byref(b)), "mpres_last_small_base")
        return a.value, b.value

    def last_binary_rounding(self):
        """True when the last fast mp_gemm rounded its exact sums once in binary (full-precision inputs)"""
        return self.lib.mpres_last_binary_rounding(self.h) == 1

    def set_workspace_limit(self, nbytes):
        """cap the workspace pool of the fast mp_gemm (0: no cap); calls that would exceed it run in reference order instead"""
        _check(self.lib.mpres_set_workspace_limit(self.h, int(nbytes)), "mpres_set_workspace_limit")

    def workspace_fallbacks(self):
        return self.lib.mpres_workspace_fallbacks(self.h)

    def workspace_bytes(self):
        return self.lib.mpres_workspace_bytes(self.h)

    def last_host_upload_residues(self):
        """residues per entry of A / B the last mp_gemm_host call uploaded (0: full records)"""
        return self.lib.mpres_last_host_upload_residues(self.h)

    def small_moduli(self, count):
        return [self.lib.mpres_small_modulus(self.h, i) for i in range(count)]

    def debug_read_workspace(self, slot, offset, nbytes):
        out = np.zeros(nbytes, dtype=np.uint8)
        n = self.lib.mpres_debug_read_workspace(self.h, int(slot), ctypes.c_size_t(offset), out.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(nbytes))
        if n < 0:
            raise MpresError("mpres_debug_read_workspace(%d) -> %d" % (slot, n))
        return out

    def set_stage1_kernel(self, kind):
        _check(self.lib.mpres_set_stage1_kernel(self.h, kind), "mpres_set_stage1_kernel")

    def last_minplus_dense_count(self):
        return self.lib.mpres_last_minplus_dense_count(self.h)

    def last_slow_count(self):
        return self.lib.mpres_last_slow_count(self.h)

    def set_stage3_kernel(self, kind):
        _check(self.lib.mpres_set_stage3_kernel(self.h, kind), "mpres_set_stage3_kernel")

    def set_vec_config(self, cfg):
        _check(self.lib.mpres_set_vec_config(self.h, cfg), "mpres_set_vec_config")

    def set_profiling(self, on):
        _check(self.lib.mpres_set_profiling(self.h, 1 if on else 0), "mpres_set_profiling")

    def last_stage_ms(self):
        ms = (ctypes.c_float * 3)()
        n = ctypes.c_int()
        _check(self.lib.mpres_last_stage_ms(self.h, ms, ctypes.byref(n)), "mpres_last_stage_ms")
        return [ms[0], ms[1], ms[2]], n.value

    def last_fallback_count(self):
        return self.lib.mpres_last_fallback_count(self.h)

    def last_kernel_ms(self):
        """[(kernel name, ms), ...] of the last fast-path mp_gemm call (profiling on), in launch order"""
        buf = ctypes.create_string_buffer(4096)
        n = self.lib.mpres_last_kernel_ms(self.h, buf, ctypes.c_size_t(4096))
        if n < 0:
            return []
        return [(kv.split("=")[0], float(kv.split("=")[1])) for kv in buf.raw[:n].decode().split(";") if "=" in kv]

    def constant(self, which, dtype, count):
        out = np.zeros(count, dtype=dtype)
        n = self.lib.mpres_get_constant(self.h, which, out.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(out.nbytes))
        if n < 0:
            raise MpresError("mpres_get_constant(%d) -> %d" % (which, n))
        return out[: n // out.itemsize]

    # --- containers (src/mparray.cuh:35-165) ---
    def mp_array_init(self, size):
        a = MpArray(self, size)
        return a

    def mp_array_from_host(self, recs):
        recs = np.ascontiguousarray(recs, dtype=self.dtype).reshape(-1)
        a = MpArray(self, recs.size)
        a.host2device(recs)
        return a

    def mp_collection_from_host(self, recs):
        recs = np.ascontiguousarray(recs, dtype=self.dtype).reshape(-1)
        a = MpCollection(self, recs.size)
        a.host2device(recs)
        return a


class MpArray:
    """mp_array_t in device memory (src/types.cuh:85-92)."""

    def __init__(self, ctx, size):
        self.ctx, self.size = ctx, int(size)
        self.s = mp_array_t()
        _check(ctx.lib.mpres_array_init(ctx.h, ctypes.byref(self.s), ctypes.c_size_t(self.size)), "mpres_array_init")

    def host2device(self, recs):
        recs = np.ascontiguousarray(recs, dtype=self.ctx.dtype).reshape(-1)
        _check(self.ctx.lib.mpres_array_host2device(self.ctx.h, ctypes.byref(self.s), recs.ctypes.data_as(ctypes.c_void_p),
                                                    ctypes.c_size_t(recs.size)), "mpres_array_host2device")

    def device2host(self, size=None):
        n = self.size if size is None else int(size)
        out = np.zeros(n, dtype=self.ctx.dtype)
        _check(self.ctx.lib.mpres_array_device2host(self.ctx.h, out.ctypes.data_as(ctypes.c_void_p), ctypes.byref(self.s),
                                                    ctypes.c_size_t(n)), "mpres_array_device2host")
        return out

    def clear(self):
        if self.s.digits:
            self.ctx.lib.mpres_array_clear(self.ctx.h, ctypes.byref(self.s))

    def __del__(self):
        try:
            if self.ctx.h:
                self.clear()
        except Exception:
            pass


class MpCollection:
    """mp_collection_t in device memory (src/types.cuh:99-104)."""

    def __init__(self, ctx, size):
        self.ctx, self.size = ctx, int(size)
        self.s = mp_collection_t()
        _check(ctx.lib.mpres_collection_init(ctx.h, ctypes.byref(self.s), ctypes.c_size_t(self.size)), "mpres_collection_init")

    def host2device(self, recs):
        recs = np.ascontiguousarray(recs, dtype=self.ctx.dtype).reshape(-1)
        _check(self.ctx.lib.mpres_collection_host2device(self.ctx.h, ctypes.byref(self.s), recs.ctypes.data_as(ctypes.c_void_p),
                                                         ctypes.c_size_t(recs.size)), "mpres_collection_host2device")

    def device2host(self):
        out = np.zeros(self.size, dtype=self.ctx.dtype)
        _check(self.ctx.lib.mpres_collection_device2host(self.ctx.h, out.ctypes.data_as(ctypes.c_void_p), ctypes.byref(self.s),
                                                         ctypes.c_size_t(self.size)), "mpres_collection_device2host")
        return out

    def clear(self):
        if self.s.digits:
            self.ctx.lib.mpres_collection_clear(self.ctx.h, ctypes.byref(self.s))

    def __del__(self):
        try:
            if self.ctx.h:
                self.clear()
        except Exception:
            pass


def _ref(a):
    return ctypes.byref(a.s) if a is not None else None


def mp_gemm(ctx, transa, transb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, buffer=None, stream=0):
    """cuda::mp_gemm (src/blas/gemm.cuh:69-70): C = alpha*op(A)*op(B) + beta*C, column-major."""
    if isinstance(A, MpCollection):
        _check(ctx.lib.mpres_gemm_coll(ctx.h, transa, transb, m, n, k, _ref(alpha), _ref(A), lda, ctypes.c_size_t(A.size),
                                       _ref(B), ldb, ctypes.c_size_t(B.size), _ref(beta), _ref(C), ldc, ctypes.c_size_t(C.size),
                                       _vp(stream)), "mpres_gemm_coll")
        return
    _check(ctx.lib.mpres_gemm(ctx.h, transa, transb, m, n, k, _ref(alpha), _ref(A), lda, _ref(B), ldb, _ref(beta), _ref(C), ldc,
                              _ref(buffer), _vp(stream)), "mpres_gemm")


def mp_gemv(ctx, trans, m, n, alpha, A, lda, x, incx, beta, y, incy, buffer1=None, buffer2=None, stream=0):
    """cuda::mp_gemv (src/blas/gemv.cuh:150-152): y = alpha*op(A)*x + beta*y."""
    if isinstance(A, MpCollection):
        _check(ctx.lib.mpres_gemv_coll(ctx.h, trans, m, n, _ref(alpha), _ref(A), lda, ctypes.c_size_t(A.size), _ref(x), incx,
                                       ctypes.c_size_t(x.size), _ref(beta), _ref(y), incy, ctypes.c_size_t(y.size), _vp(stream)),
               "mpres_gemv_coll")
        return
    _check(ctx.lib.mpres_gemv(ctx.h, trans, m, n, _ref(alpha), _ref(A), lda, _ref(x), incx, _ref(beta), _ref(y), incy,
                              _ref(buffer1), _ref(buffer2), _vp(stream)), "mpres_gemv")


def mp_dot(ctx, n, x, incx, y, incy, r, buffer=None, stream=0):
    """cuda::mp_dot (src/blas/dot.cuh:84-85): r[0] = sum x_i*y_i."""
    if isinstance(x, MpCollection):
        _check(ctx.lib.mpres_dot_coll(ctx.h, n, _ref(x), incx, ctypes.c_size_t(x.size), _ref(y), incy, ctypes.c_size_t(y.size),
                                      _ref(r), _vp(stream)), "mpres_dot_coll")
        return
    _check(ctx.lib.mpres_dot(ctx.h, n, _ref(x), incx, _ref(y), incy, _ref(r), _ref(buffer), _vp(stream)), "mpres_dot")


mblas_one_norm, mblas_inf_norm = 171, 175   # src/blas/mblas_enum.cuh:41-44


def mp_asum(ctx, n, x, incx, r, stream=0):
    """cuda::mp_asum (src/blas/asum.cuh:41): r[0] = sum |x_i|."""
    _check(ctx.lib.mpres_asum(ctx.h, n, _ref(x), incx, _ref(r), _vp(stream)), "mpres_asum")


def mp_norm(ctx, norm, n, x, incx, r, stream=0):
    """cuda::mp_norm (src/blas/norm.cuh:43): one norm (sum of magnitudes) or infinity norm (largest magnitude) of a vector."""
    _check(ctx.lib.mpres_norm(ctx.h, norm, n, _ref(x), incx, _ref(r), _vp(stream)), "mpres_norm")


def mp_ge_norm(ctx, norm, m, n, A, lda, r, buffer=None, stream=0):
    """cuda::mp_ge_norm (src/blas/genorm.cuh:142): one norm (largest column sum) or infinity norm (largest row sum) of a matrix."""
    _check(ctx.lib.mpres_ge_norm(ctx.h, norm, m, n, _ref(A), lda, _ref(r), _ref(buffer), _vp(stream)), "mpres_ge_norm")


def _dev_ptr(t):
    """device address of a torch CUDA tensor (or a raw integer address)"""
    return ctypes.c_void_p(t if isinstance(t, int) else t.data_ptr())


def mp_spmv_mpmtx_csr2st(ctx, m, n, nnz, irp, ja, As, x, y, buffer=None, stream=0):
    """cuda::mp_spmv_mpmtx_csr2st (src/sparse/mpmtx/spmv_mpmtx_csr2st.cuh:106): y = A x, A in CSR with mp_collection_t entries; irp, ja: int32 device tensors."""
    _check(ctx.lib.mpres_spmv_csr2st(ctx.h, m, n, nnz, _dev_ptr(irp), _dev_ptr(ja), _ref(As), _ref(x), _ref(y), _ref(buffer), _vp(stream)), "mpres_spmv_csr2st")


def mp_spmv_mpmtx_ell2st(ctx, m, n, maxnzr, ja, As, x, y, buffer=None, stream=0):
    """cuda::mp_spmv_mpmtx_ell2st (src/sparse/mpmtx/spmv_mpmtx_ell2st.cuh:119): y = A x, A in ELLPACK (column-major m x maxnzr, ja < 0 = padding)."""
    _check(ctx.lib.mpres_spmv_ell2st(ctx.h, m, n, maxnzr, _dev_ptr(ja), _ref(As), _ref(x), _ref(y), _ref(buffer), _vp(stream)), "mpres_spmv_ell2st")


def mp_array_set_d(ctx, dst, offset, src, n, stream=0):
    """dst[offset + i] = src[i] (mp_set_d, src/arith/assign.cuh:54-81); src: float64 device tensor."""
    _check(ctx.lib.mpres_array_set_d(ctx.h, _ref(dst), ctypes.c_size_t(offset), _dev_ptr(src), ctypes.c_size_t(n), _vp(stream)), "mpres_array_set_d")


def mp_array_get_d(ctx, dst, src, offset, n, stream=0):
    """dst[i] = nearest double of src[offset + i] (mp_get_d, src/arith/assign.cuh:154-180); dst: float64 device tensor."""
    _check(ctx.lib.mpres_array_get_d(ctx.h, _dev_ptr(dst), _ref(src), ctypes.c_size_t(offset), ctypes.c_size_t(n), _vp(stream)), "mpres_array_get_d")


def mp_div(ctx, r, x, y, stream=0):
    """cuda::mp_div (src/arith/div.cuh:56-65): r[0] = x[0] / y[0], rounded to nearest at the working precision."""
    _check(ctx.lib.mpres_div(ctx.h, _ref(r), _ref(x), _ref(y), _vp(stream)), "mpres_div")


def mp_cg_csr(ctx, n, nnz, irp, ja, vals, b, tol, maxit, x, M=None, stream=0):
    """mp_cg_csr / mp_pcg_csr (src/sparse/solver/cg_csr.cuh:53, pcg_csr.cuh:57): returns (iterations, relative residual history)."""
    it = ctypes.c_int(0)
    res = (ctypes.c_double * (maxit + 1))()
    ctx.lib.mpres_cg_csr.argtypes = None
    _check(ctx.lib.mpres_cg_csr(ctx.h, n, nnz, _dev_ptr(irp), _dev_ptr(ja), _dev_ptr(vals), _ref(b), ctypes.c_double(tol), maxit,
                                _dev_ptr(M) if M is not None else None, _ref(x), ctypes.byref(it), res, _vp(stream)), "mpres_cg_csr")
    return it.value, list(res)[: it.value + 1]


def mp_scal(ctx, n, alpha, x, incx, stream=0):
    """cuda::mp_scal (src/blas/scal.cuh:45-46): x = alpha*x."""
    _check(ctx.lib.mpres_scal(ctx.h, n, _ref(alpha), _ref(x), incx, _vp(stream)), "mpres_scal")


def mp_axpy(ctx, n, alpha, x, incx, y, incy, buffer=None, stream=0):
    """cuda::mp_axpy (src/blas/axpy.cuh:46): y = alpha*x + y."""
    _check(ctx.lib.mpres_axpy(ctx.h, n, _ref(alpha), _ref(x), incx, _ref(y), incy, _ref(buffer), _vp(stream)), "mpres_axpy")


def mp_waxpby(ctx, n, alpha, x, incx, beta, y, incy, w, incw, buffer=None, stream=0):
    """cuda::mp_waxpby (src/blas/waxpby.cuh:50): w = alpha*x + beta*y."""
    _check(ctx.lib.mpres_waxpby(ctx.h, n, _ref(alpha), _ref(x), incx, _ref(beta), _ref(y), incy, _ref(w), incw, _ref(buffer), _vp(stream)), "mpres_waxpby")


def mp_ge_add(ctx, m, n, alpha, A, lda, beta, B, ldb, C, ldc, buffer=None, stream=0):
    """cuda::mp_ge_add (src/blas/geadd.cuh:58): C = alpha*A + beta*B."""
    _check(ctx.lib.mpres_ge_add(ctx.h, m, n, _ref(alpha), _ref(A), lda, _ref(beta), _ref(B), ldb, _ref(C), ldc, _ref(buffer), _vp(stream)), "mpres_ge_add")


def mp_ge_acc(ctx, m, n, alpha, A, lda, beta, B, ldb, buffer=None, stream=0):
    """cuda::mp_ge_acc (src/blas/geacc.cuh:57): B = alpha*A + beta*B."""
    _check(ctx.lib.mpres_ge_acc(ctx.h, m, n, _ref(alpha), _ref(A), lda, _ref(beta), _ref(B), ldb, _ref(buffer), _vp(stream)), "mpres_ge_acc")


def mp_ger(ctx, m, n, alpha, x, incx, y, incy, A, lda, buffer1=None, buffer2=None, stream=0):
    """cuda::mp_ger (src/blas/ger.cuh:157): A = alpha*x*y^T + A."""
    _check(ctx.lib.mpres_ger(ctx.h, m, n, _ref(alpha), _ref(x), incx, _ref(y), incy, _ref(A), lda, _ref(buffer1), _ref(buffer2), _vp(stream)), "mpres_ger")


LEFT_SIDE, RIGHT_SIDE = 141, 142     # mblas_side_type, src/blas/mblas_enum.cuh:37-40


def mp_ge_diag_scale(ctx, side, m, n, D, incd, A, lda, stream=0):
    """cuda::mp_ge_diag_scale (src/blas/gediagscale.cuh:54): A = A*D (RIGHT_SIDE) or D*A (LEFT_SIDE), D diagonal, stored as a vector."""
    _check(ctx.lib.mpres_ge_diag_scale(ctx.h, side, m, n, _ref(D), incd, _ref(A), lda, _vp(stream)), "mpres_ge_diag_scale")


def mp_ge_lr_scale(ctx, m, n, DL, incdl, DR, incdr, A, lda, stream=0):
    """cuda::mp_ge_lr_scale (src/blas/gelrscale.cuh:56): A = DL*A*DR."""
    _check(ctx.lib.mpres_ge_lr_scale(ctx.h, m, n, _ref(DL), incdl, _ref(DR), incdr, _ref(A), lda, _vp(stream)), "mpres_ge_lr_scale")


def mp_rot(ctx, n, x, incx, y, incy, c, s, buffer1=None, buffer2=None, stream=0):
    """cuda::mp_rot (src/blas/rot.cuh:49): x = c*x + s*y, y = c*y - s*x."""
    _check(ctx.lib.mpres_rot(ctx.h, n, _ref(x), incx, _ref(y), incy, _ref(c), _ref(s), _ref(buffer1), _ref(buffer2), _vp(stream)), "mpres_rot")


def mp_axpy_dot(ctx, n, alpha, w, incw, v, incv, u, incu, r, buffer=None, stream=0):
    """cuda::mp_axpy_dot (src/blas/axpydot.cuh:33): w = w - alpha*v, r[0] = u^T w."""
    _check(ctx.lib.mpres_axpy_dot(ctx.h, n, _ref(alpha), _ref(w), incw, _ref(v), incv, _ref(u), incu, _ref(r), _ref(buffer), _vp(stream)), "mpres_axpy_dot")


class Shard:
    """Communicator of the row-sharded mp_gemm (mpres_shard_*): this rank's receive buffer and the peer mappings.
    `exchange(handle_bytes) -> list of every rank's handle bytes in rank order` is the caller's all-gather (torch.distributed,
    MPI, or a plain list inside one process)."""

    def __init__(self, ctx, rank, world, n, k_max):
        self.ctx, self.rank, self.world = ctx, rank, world
        self.h = ctypes.c_void_p()
        _check(ctx.lib.mpres_shard_create(ctx.h, rank, world, n, k_max, ctypes.byref(self.h)), "mpres_shard_create")

    def export(self):
        size = self.ctx.lib.mpres_shard_handle_size()
        buf = ctypes.create_string_buffer(size)
        _check(self.ctx.lib.mpres_shard_export(self.h, buf), "mpres_shard_export")
        return buf.raw

    def connect(self, handles):
        blob = b"".join(handles)
        assert len(blob) == self.world * self.ctx.lib.mpres_shard_handle_size()
        _check(self.ctx.lib.mpres_shard_connect(self.h, blob), "mpres_shard_connect")

    def gemm(self, transa, transb, m_local, n, k, alpha, A, lda, B, ldb, beta, C, ldc, stream=0):
        """collective: C_r = alpha * op(A_r) * op(B) + beta * C_r on this rank's row block; B is the rank's complete copy"""
        _check(self.ctx.lib.mpres_gemm_sharded(self.h, transa, transb, m_local, n, k, _ref(alpha), _ref(A), lda, _ref(B), ldb, _ref(beta), _ref(C), ldc,
                                               _vp(stream)), "mpres_gemm_sharded")

    def close(self):
        if self.h:
            self.ctx.lib.mpres_shard_destroy(self.h)
            self.h = ctypes.c_void_p()

    def __del__(self):
        try:
            if self.ctx.h:
                self.close()
        except Exception:
            pass


def _host_ptr(a):
    """address of a host buffer: an int, a numpy array of mp_float_t records, or anything with data_ptr() (a pinned torch tensor)"""
    if isinstance(a, int):
        return ctypes.c_void_p(a)
    if hasattr(a, "data_ptr"):
        return ctypes.c_void_p(a.data_ptr())
    return ctypes.c_void_p(a.ctypes.data)


def mp_gemm_host(ctx, transa, transb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, out=None, panels=0):
    """mpres_gemm_host: cuda::mp_gemm over HOST mp_float_t[] operands, transfers pipelined with the compute by column panels;
    the result goes to `out` (default: C, in place)."""
    _check(ctx.lib.mpres_gemm_host(ctx.h, transa, transb, m, n, k, _host_ptr(alpha), _host_ptr(A), lda, _host_ptr(B), ldb, _host_ptr(beta),
                                   _host_ptr(C), _host_ptr(C if out is None else out), ldc, panels), "mpres_gemm_host")


def mp_gemm_host_bdev(ctx, transa, transb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, out=None, panels=0):
    """mpres_gemm_host_bdev: like mp_gemm_host, with B a device-resident mp_array_t (complete before the call)."""
    _check(ctx.lib.mpres_gemm_host_bdev(ctx.h, transa, transb, m, n, k, _host_ptr(alpha), _host_ptr(A), lda, _ref(B), ldb, _host_ptr(beta),
                                        _host_ptr(C), _host_ptr(C if out is None else out), ldc, panels), "mpres_gemm_host_bdev")


def synchronize(ctx):
    """cudaDeviceSynchronize through the library's fallback-counter read (it syncs the last stream)."""
    return ctx.last_fallback_count()
