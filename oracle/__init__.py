"""oracle -- TEST INFRASTRUCTURE ONLY.

CPU checker for the mp_gemm / mp_gemv / mp_dot path: `Oracle` wraps liboracle.so (the plain-C
restatement in mpres_oracle.c, constants from constants.py); `RefLib` wraps oracle/_ref/
libmpres_ref_N<k>.so, i.e. the UNMODIFIED reference compiled from /root/reference (host functions and
CUDA kernels).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package; the product (mpres-blas_b200/) never does.
"""
import ctypes
import json
import os
import subprocess
import time

import numpy as np

from . import constants

_HERE = os.path.dirname(os.path.abspath(__file__))
HOST, DEVICE = 0, 1


def moduli_sets():
    with open(os.path.join(_HERE, "moduli_sets.json")) as f:
        return {int(k): v for k, v in json.load(f).items()}


def record_dtype(N):
    """numpy view of the reference's AoS mp_float_t (types.cuh:69-74) for N moduli."""
    return np.dtype([("digits", np.int32, (N,)), ("sign", np.int32), ("exp", np.int32),
                     ("eval", [("frac", np.float64), ("exp", np.int64)], (2,))])


def build_oracle(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = [os.path.join(_HERE, f) for f in ("mpres_oracle.c", "mpres_oracle.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "oracle"], stdout=subprocess.DEVNULL)
    return so


class _Er(ctypes.Structure):
    _fields_ = [("frac", ctypes.c_double), ("exp", ctypes.c_long)]


def _ip(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class Oracle:
    """The C restatement for one moduli set and one flavour (HOST or DEVICE semantics)."""

    def __init__(self, N, flavor=DEVICE):
        self.N, self.flavor = N, flavor
        self.lib = ctypes.CDLL(build_oracle())
        L = self.lib
        L.orc_create.restype = ctypes.c_void_p
        self.c = constants.compute(moduli_sets()[N])
        c = self.c
        self.dtype = record_dtype(N)
        assert self.dtype.itemsize == 4 * N + 40
        i32 = lambda x: np.ascontiguousarray(np.array(x, dtype=np.int64).astype(np.int32))
        f64 = lambda x: np.ascontiguousarray(np.array(x, dtype=np.float64))
        self._keep = [i32(c["moduli"]), i32(c["part_inverse"]), i32(c["pow2"]), i32(c["m_pow2_residues"]),
                      i32(c["mi_pow2_residues"]), i32(c["pow2_inverse"]), i32(c["mrc_mult_inv"]),
                      f64(c["recip_rd"]), f64(c["recip_ru"])]
        er = lambda t: _Er(t[0], t[1])
        k = self._keep
        self.ctx = ctypes.c_void_p(L.orc_create(
            N, c["log2M"], flavor, c["mp_h"], c["mp_j"], c["eval_ref_factor"], ctypes.c_double(c["eval_accuracy"]),
            ctypes.byref(er(c["eval_unit_low"])), ctypes.byref(er(c["eval_unit_upp"])),
            ctypes.byref(er(c["eval_inv_unit_low"])), ctypes.byref(er(c["eval_inv_unit_upp"])),
            _ip(k[0]), _ip(k[1]), _ip(k[2]), _ip(k[3]), _ip(k[4]), _ip(k[5]), _ip(k[6]), _ip(k[7]), _ip(k[8])))
        assert self.ctx.value
        L.orc_dot_omp.restype = ctypes.c_int
        L.orc_gemm_rows.restype = ctypes.c_int
        L.orc_gemv.restype = ctypes.c_int
        L.orc_mrc_compare.restype = ctypes.c_int

    def __del__(self):
        try:
            self.lib.orc_destroy(self.ctx)
        except Exception:
            pass

    @property
    def precision(self):
        return self.c["mp_precision"]

    def empty(self, shape):
        return np.zeros(shape, dtype=self.dtype)

    def set_ints(self, signs, mants, exps):
        """records for (-1)^s * mant * 2^exp (trailing zero bits trimmed, eval computed)"""
        n = len(mants)
        out = self.empty(n)
        for i in range(n):
            mant = int(mants[i])
            nl = max(1, (mant.bit_length() + 31) // 32)
            limbs = np.array([(mant >> (32 * j)) & 0xFFFFFFFF for j in range(nl)], dtype=np.uint32)
            self.lib.orc_mp_set(self.ctx, _ip(out[i:i + 1]), int(signs[i]), _ip(limbs), nl, int(exps[i]))
        return out

    def random_records(self, count, bits, seed):
        """bulk synthetic records (C, OpenMP) for the CPU-baseline legs"""
        out = self.empty(count)
        self.lib.orc_random_fill(self.ctx, _ip(out), ctypes.c_long(count), int(bits), ctypes.c_uint64(seed))
        return out

    def _bin(self, fn, x, y):
        x = np.ascontiguousarray(x)
        y = np.ascontiguousarray(y)
        r = np.zeros_like(x)
        getattr(self.lib, fn)(self.ctx, _ip(r), _ip(x), _ip(y), ctypes.c_long(x.size))
        return r

    def mul(self, x, y):
        return self._bin("orc_mul_vec", x, y)

    def add(self, x, y):
        return self._bin("orc_add_vec", x, y)

    def round(self, x, bits):
        r = np.ascontiguousarray(x).copy()
        for i in range(r.size):
            self.lib.orc_mp_round(self.ctx, _ip(r[i:i + 1]), int(bits[i]))
        return r

    def eval(self, x, fast=False):
        r = np.ascontiguousarray(x).copy()
        fn = self.lib.orc_eval_compute_fast if fast else self.lib.orc_eval_compute
        for i in range(r.size):
            ev = np.zeros(2, dtype=[("frac", np.float64), ("exp", np.int64)])
            fn(self.ctx, _ip(ev[0:1]), _ip(ev[1:2]), _ip(np.ascontiguousarray(r["digits"][i])))
            r["eval"][i] = ev
        return r

    def scale2pow(self, digits, D):
        d = np.ascontiguousarray(digits, dtype=np.int32)
        r = np.zeros_like(d)
        self.lib.orc_scale2pow(self.ctx, _ip(r), _ip(d), ctypes.c_uint(D))
        return r

    def dot_seq(self, x, y):
        r = self.empty(1)
        self.lib.orc_dot_seq(self.ctx, _ip(r), _ip(np.ascontiguousarray(x)), _ip(np.ascontiguousarray(y)), ctypes.c_long(x.size))
        return r[0]

    def dot_omp(self, x, y):
        r = self.empty(1)
        nt = self.lib.orc_dot_omp(self.ctx, _ip(r), _ip(np.ascontiguousarray(x)), _ip(np.ascontiguousarray(y)), ctypes.c_long(x.size))
        return r[0], nt

    def dot_v1(self, x, y, grid=256, block=64):
        r = self.empty(1)
        self.lib.orc_dot_v1(self.ctx, _ip(r), _ip(np.ascontiguousarray(x)), _ip(np.ascontiguousarray(y)),
                            ctypes.c_long(x.size), grid, block)
        return r[0]

    def gemm(self, m, n, k, alpha, A, B, beta, C, rows=None, want_ab=False):
        """A: (k, lda=m) column-major flattened as array of m*k records; returns (C_out, AB or None)"""
        A = np.ascontiguousarray(A).reshape(-1)
        B = np.ascontiguousarray(B).reshape(-1)
        Cw = np.ascontiguousarray(C).reshape(-1).copy()
        ab = self.empty(m * n) if want_ab else None
        r0, r1 = rows if rows else (0, m)
        self.lib.orc_gemm_rows(self.ctx, r0, r1, m, n, k, _ip(np.ascontiguousarray(alpha).reshape(-1)), _ip(A), m, _ip(B), k,
                               _ip(np.ascontiguousarray(beta).reshape(-1)), _ip(Cw), m, _ip(ab) if want_ab else None)
        return Cw, ab

    def gemv(self, trans, m, n, alpha, A, x, beta, y, block=32):
        yw = np.ascontiguousarray(y).reshape(-1).copy()
        self.lib.orc_gemv(self.ctx, trans, m, n, _ip(np.ascontiguousarray(alpha).reshape(-1)), _ip(np.ascontiguousarray(A).reshape(-1)), m,
                          _ip(np.ascontiguousarray(x).reshape(-1)), _ip(np.ascontiguousarray(beta).reshape(-1)), _ip(yw), block)
        return yw

    # exact value helpers (Python integers; small inputs only)
    def to_int(self, rec):
        """CRT: the integer significand of one record (rns.cuh:289-299)"""
        M = self.c["M"]
        mods = self.c["moduli"]
        x = 0
        for i, m in enumerate(mods):
            Mi = M // m
            x += Mi * ((int(rec["digits"][i]) * self.c["part_inverse"][i]) % m)
        return x % M

    def to_fraction(self, rec):
        from fractions import Fraction
        v = Fraction(self.to_int(rec)) * Fraction(2) ** int(rec["exp"])
        return -v if int(rec["sign"]) else v


def ref_lib_path(N):
    return os.path.join(_HERE, "_ref", "libmpres_ref_N%d.so" % N)


def have_ref(N):
    return os.path.exists(ref_lib_path(N))


class RefLib:
    """The unmodified reference (host functions + CUDA kernels) for one compiled-in moduli set."""

    def __init__(self, N, gpu=False):
        self.N = N
        self.lib = ctypes.CDLL(ref_lib_path(N))
        L = self.lib
        assert L.ref_moduli_size() == N
        (L.ref_gpu_init if gpu else L.ref_init)()
        self.dtype = record_dtype(N)
        assert L.ref_sizeof_mp_float() == self.dtype.itemsize
        for f in ("ref_gpu_gemm", "ref_gpu_gemm_v2", "ref_gpu_gemv", "ref_gpu_dot"):
            getattr(L, f).restype = ctypes.c_float
        L.ref_mpfr_dot_timed.restype = ctypes.c_double

    def empty(self, shape):
        return np.zeros(shape, dtype=self.dtype)

    def set_ints(self, signs, mants, exps, prec):
        n = len(mants)
        out = self.empty(n)
        for i in range(n):
            mant = int(mants[i])
            nb = max(1, (mant.bit_length() + 7) // 8)
            b = mant.to_bytes(nb, "little")
            self.lib.ref_set_from_int(_ip(out[i:i + 1]), int(signs[i]), b, nb, ctypes.c_long(int(exps[i])), int(prec))
        return out

    def to_int(self, rec):
        buf = ctypes.create_string_buffer(4096)
        n = self.lib.ref_get_mantissa(_ip(np.ascontiguousarray(rec).reshape(1)), buf, 4096)
        return int.from_bytes(buf.raw[:n], "little")

    def _bin(self, fn, x, y):
        x = np.ascontiguousarray(x)
        y = np.ascontiguousarray(y)
        r = np.zeros_like(x)
        getattr(self.lib, fn)(_ip(r), _ip(x), _ip(y), ctypes.c_long(x.size))
        return r

    def host_mul(self, x, y):
        return self._bin("ref_host_mul_vec", x, y)

    def host_add(self, x, y):
        return self._bin("ref_host_add_vec", x, y)

    def host_round(self, x, bits):
        r = np.ascontiguousarray(x).copy()
        for i in range(r.size):
            self.lib.ref_host_round(_ip(r[i:i + 1]), int(bits[i]))
        return r

    def host_eval(self, x, fast=False):
        r = np.ascontiguousarray(x).copy()
        fn = self.lib.ref_host_eval_fast if fast else self.lib.ref_host_eval
        for i in range(r.size):
            ev = np.zeros(2, dtype=[("frac", np.float64), ("exp", np.int64)])
            fn(_ip(ev[0:1]), _ip(ev[1:2]), _ip(np.ascontiguousarray(r["digits"][i])))
            r["eval"][i] = ev
        return r

    def host_dot(self, x, y):
        r = self.empty(1)
        self.lib.ref_host_dot(_ip(r), _ip(np.ascontiguousarray(x)), _ip(np.ascontiguousarray(y)), ctypes.c_long(x.size))
        return r[0]

    def host_dot_omp(self, x, y):
        r = self.empty(1)
        self.lib.ref_host_dot_omp.restype = ctypes.c_int
        nt = self.lib.ref_host_dot_omp(_ip(r), _ip(np.ascontiguousarray(x)), _ip(np.ascontiguousarray(y)), ctypes.c_long(x.size))
        return r[0], nt

    def host_gemm(self, m, n, k, alpha, A, B, beta, C, rows=None):
        Cw = np.ascontiguousarray(C).reshape(-1).copy()
        r0, r1 = rows if rows else (0, m)
        self.lib.ref_host_gemm_rows.restype = ctypes.c_int
        nt = self.lib.ref_host_gemm_rows(r0, r1, n, k, _ip(np.ascontiguousarray(alpha).reshape(-1)), _ip(np.ascontiguousarray(A).reshape(-1)), m,
                                         _ip(np.ascontiguousarray(B).reshape(-1)), k, _ip(np.ascontiguousarray(beta).reshape(-1)), _ip(Cw), m)
        return Cw, nt

    def host_gemv(self, m, n, alpha, A, x, beta, y, rows=None):
        yw = np.ascontiguousarray(y).reshape(-1).copy()
        r0, r1 = rows if rows else (0, m)
        self.lib.ref_host_gemv_rows.restype = ctypes.c_int
        nt = self.lib.ref_host_gemv_rows(r0, r1, n, _ip(np.ascontiguousarray(alpha).reshape(-1)), _ip(np.ascontiguousarray(A).reshape(-1)), m,
                                         _ip(np.ascontiguousarray(x).reshape(-1)), _ip(np.ascontiguousarray(beta).reshape(-1)), _ip(yw))
        return yw, nt

    def mpfr_gemm_timed(self, m, n, k, alpha, A, B, beta, C, prec):
        """seconds and threads of the reference's MPFR GEMM loop (tests/blas/v2/gemm/test_mpfr_gemm.cuh:29-60) on an m x n x k problem"""
        self.lib.ref_mpfr_gemm_timed.restype = ctypes.c_double
        nt = ctypes.c_int()
        secs = self.lib.ref_mpfr_gemm_timed(m, n, k, _ip(np.ascontiguousarray(alpha).reshape(-1)), _ip(np.ascontiguousarray(A).reshape(-1)),
                                            _ip(np.ascontiguousarray(B).reshape(-1)), _ip(np.ascontiguousarray(beta).reshape(-1)),
                                            _ip(np.ascontiguousarray(C).reshape(-1)), int(prec), ctypes.byref(nt))
        return secs, nt.value

    def to_string(self, rec, prec, digits=40):
        buf = ctypes.create_string_buffer(4096)
        self.lib.ref_to_string(_ip(np.ascontiguousarray(rec).reshape(1)), int(prec), int(digits), buf, 4096)
        return buf.value.decode()

    # CUDA kernels of the reference (GPU box only)
    def gpu_asum_norm(self, kind, x, n, incx=1, cfg=0):
        """kind: 0 mp_asum, 171 one-norm, 175 inf-norm (src/blas/asum.cuh:41, norm.cuh:43)"""
        r = self.empty(1)
        self.lib.ref_gpu_asum_norm.restype = ctypes.c_float
        self.lib.ref_gpu_asum_norm(int(kind), int(n), _ip(np.ascontiguousarray(x)), int(incx), _ip(r), int(cfg))
        return r[0]

    def gpu_ge_norm(self, kind, m, n, A, lda):
        r = self.empty(1)
        self.lib.ref_gpu_ge_norm.restype = ctypes.c_float
        self.lib.ref_gpu_ge_norm(int(kind), int(m), int(n), _ip(np.ascontiguousarray(A).reshape(-1)), int(lda), _ip(r))
        return r[0]

    def gpu_spmv_2st(self, fmt, m, n, nnz, ptr, idx, vals, x):
        """fmt 0: CSR (mp_spmv_mpmtx_csr2st), 1: ELLPACK (mp_spmv_mpmtx_ell2st, nnz = maxnzr)"""
        y = self.empty(m)
        self.lib.ref_gpu_spmv_2st.restype = ctypes.c_float
        p = np.ascontiguousarray(ptr, dtype=np.int32) if ptr is not None else None
        t0 = time.time()
        self.last_kernel_ms = self.lib.ref_gpu_spmv_2st(int(fmt), int(m), int(n), int(nnz), _ip(p) if p is not None else None, _ip(np.ascontiguousarray(idx, dtype=np.int32)),
                                                        _ip(np.ascontiguousarray(vals).reshape(-1)), _ip(np.ascontiguousarray(x).reshape(-1)), _ip(y))
        self.last_wall_s = time.time() - t0
        return y

    def gpu_gemm(self, m, n, k, alpha, A, B, beta, C, want_ab=False, repeat=1):
        Cw = np.ascontiguousarray(C).reshape(-1).copy()
        ab = self.empty(m * n) if want_ab else None
        ms = self.lib.ref_gpu_gemm(m, n, k, _ip(np.ascontiguousarray(alpha).reshape(-1)), _ip(np.ascontiguousarray(A).reshape(-1)),
                                   _ip(np.ascontiguousarray(B).reshape(-1)), _ip(np.ascontiguousarray(beta).reshape(-1)), _ip(Cw),
                                   _ip(ab) if want_ab else None, repeat)
        return Cw, ab, ms

    def gpu_gemv(self, trans, m, n, alpha, A, x, beta, y, repeat=1):
        yw = np.ascontiguousarray(y).reshape(-1).copy()
        ms = self.lib.ref_gpu_gemv(trans, m, n, _ip(np.ascontiguousarray(alpha).reshape(-1)), _ip(np.ascontiguousarray(A).reshape(-1)),
                                   _ip(np.ascontiguousarray(x).reshape(-1)), _ip(np.ascontiguousarray(beta).reshape(-1)), _ip(yw), repeat)
        return yw, ms

    def gpu_dot(self, x, y, cfg=0, repeat=1):
        r = self.empty(1)
        ms = self.lib.ref_gpu_dot(int(x.size), _ip(np.ascontiguousarray(x)), _ip(np.ascontiguousarray(y)), _ip(r), cfg, repeat)
        return r[0], ms

    def gpu_probe(self, op, x, y=None, bits=None):
        x = np.ascontiguousarray(x)
        r = np.zeros_like(x)
        b = np.ascontiguousarray(bits, dtype=np.int32) if bits is not None else None
        self.lib.ref_gpu_probe(op, _ip(r), _ip(x), _ip(np.ascontiguousarray(y)) if y is not None else None,
                               _ip(b) if b is not None else None, int(x.size))
        return r
