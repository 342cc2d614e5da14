"""mp_scal and mp_axpy through the C-ABI against the C oracle's mp_mul / mp_add (the DEVICE flavour, itself pinned
against the reference's cuda:: functions): digits, sign and exponent bit for bit, with roundings (p-bit inputs), BLAS strides and the
silent-return cases of src/blas/scal.cuh:47-49 and src/blas/axpy.cuh:49-51."""
import numpy as np
import pytest

import oracle
from util import diff_fields, get_oracle, random_records

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _blas_view(n, inc):
    """indices of a BLAS vector of n elements with stride inc (negative: reversed, mpvector.cuh:68-70)"""
    return [i * inc if inc > 0 else (-n + i + 1) * inc for i in range(n)]


@pytest.mark.parametrize("N,full,incx", [(8, False, 1), (8, True, 2), (32, False, 3), (32, True, 1), (16, True, 1)])
def test_scal(pkg, N, full, incx):
    ctx = pkg.Context(N, 0)
    orc = get_oracle(N, oracle.DEVICE)
    bits = orc.precision if full else orc.precision // 4
    n = 257
    x = random_records(N, n * incx, bits, 11)
    al = random_records(N, 1, bits, 12)
    dx, dal = ctx.mp_array_from_host(x), ctx.mp_array_from_host(al)
    pkg.mp_scal(ctx, n, dal, dx, incx)
    got = dx.device2host()
    idx = _blas_view(n, incx)
    want = x.copy()
    want[idx] = orc.mul(x[idx], np.repeat(al, n))
    assert diff_fields(got, want, ("digits", "sign", "exp")).size == 0
    # silent returns: nothing changes
    pkg.mp_scal(ctx, 0, dal, dx, incx)
    pkg.mp_scal(ctx, n, dal, dx, 0)
    assert diff_fields(dx.device2host(), got).size == 0
    ctx.close()


@pytest.mark.parametrize("N,full,incx,incy", [(8, False, 1, 1), (8, True, 1, 1), (32, False, 2, -1), (32, True, -1, 3), (24, True, 1, 1)])
def test_axpy(pkg, N, full, incx, incy):
    ctx = pkg.Context(N, 0)
    orc = get_oracle(N, oracle.DEVICE)
    bits = orc.precision if full else orc.precision // 4
    n = 300
    x = random_records(N, n * abs(incx), bits, 21)
    y = random_records(N, n * abs(incy), bits, 22)
    al = random_records(N, 1, bits, 23)
    ix, iy = _blas_view(n, incx), _blas_view(n, incy)
    # a few exact cancellations and zeros: y = -(alpha * x) after rounding, x = 0, y = 0
    prod = orc.mul(x[ix], np.repeat(al, n))
    y[iy[5]] = prod[5]; y[iy[5]]["sign"] ^= 1
    zero = orc.set_ints([0], [0], [0])[0]
    x[ix[7]] = zero
    y[iy[9]] = zero
    prod = orc.mul(x[ix], np.repeat(al, n))
    dx, dy, dal = ctx.mp_array_from_host(x), ctx.mp_array_from_host(y), ctx.mp_array_from_host(al)
    pkg.mp_axpy(ctx, n, dal, dx, incx, dy, incy)
    got = dy.device2host()
    want = y.copy()
    want[iy] = orc.add(prod, y[iy])
    bad = diff_fields(got, want, ("digits", "sign", "exp"))
    assert bad.size == 0, "%d differ, first %d\n%s\n%s" % (bad.size, bad[0], got[bad[0]], want[bad[0]])
    assert orc.to_fraction(got[iy[5]]) == 0
    pkg.mp_axpy(ctx, 0, dal, dx, incx, dy, incy)
    assert diff_fields(dy.device2host(), got).size == 0
    ctx.close()


def test_gemv_t_row_sharded_composition(pkg):
    """The multi-GPU GEMV (T) recipe of bench.py / SURVEY 8(e), run on one device: t_r = alpha A_r^T x_r per row block, y = round(beta y),
    y += t_r in block order (mp_axpy with the scalar one) == one mp_gemv(T) on the whole matrix (p/4-bit inputs: nothing rounds)."""
    N, m, n, G = 16, 96, 70, 3
    ctx = pkg.Context(N, 0)
    orc = get_oracle(N, oracle.DEVICE)
    bits = orc.precision // 4
    A = random_records(N, m * n, bits, 31)      # column-major m x n
    x, y = random_records(N, m, bits, 32), random_records(N, n, bits, 33)
    al, be = random_records(N, 1, bits, 34), random_records(N, 1, bits, 35)
    one = orc.set_ints([0], [1], [0])
    zero = orc.set_ints([0] * n, [0] * n, [0] * n)
    dal, dbe, done = ctx.mp_array_from_host(al), ctx.mp_array_from_host(be), ctx.mp_array_from_host(one)
    dA, dx, dy = ctx.mp_array_from_host(A), ctx.mp_array_from_host(x), ctx.mp_array_from_host(y)
    pkg.mp_gemv(ctx, pkg.mblas_trans, m, n, dal, dA, m, dx, 1, dbe, dy, 1)
    want = dy.device2host()
    dy2 = ctx.mp_array_from_host(y)
    pkg.mp_scal(ctx, n, dbe, dy2, 1)
    A2 = A.reshape(n, m)
    for r in range(G):
        lo, hi = m * r // G, m * (r + 1) // G
        dAr = ctx.mp_array_from_host(np.ascontiguousarray(A2[:, lo:hi]).reshape(-1))     # compact shard, lda = hi - lo
        dxr = ctx.mp_array_from_host(x[lo:hi])
        dt = ctx.mp_array_from_host(zero)
        pkg.mp_gemv(ctx, pkg.mblas_trans, hi - lo, n, dal, dAr, hi - lo, dxr, 1, dbe, dt, 1)
        pkg.mp_axpy(ctx, n, done, dt, 1, dy2, 1)
    got = dy2.device2host()
    bad = diff_fields(got, want, ("digits", "sign", "exp"))
    assert bad.size == 0, "%d/%d differ, first %d\n%s\n%s" % (bad.size, n, bad[0], got[bad[0]], want[bad[0]])
    ctx.close()


@pytest.mark.parametrize("N,full,incs", [(8, False, (1, 1, 1)), (16, True, (2, -1, 1)), (32, True, (1, 1, 2))])
def test_waxpby(pkg, N, full, incs):
    ctx = pkg.Context(N, 0)
    orc = get_oracle(N, oracle.DEVICE)
    bits = orc.precision if full else orc.precision // 4
    n = 211
    incx, incy, incw = incs
    x, y, w = random_records(N, n * abs(incx), bits, 41), random_records(N, n * abs(incy), bits, 42), random_records(N, n * abs(incw), bits, 43)
    al, be = random_records(N, 1, bits, 44), random_records(N, 1, bits, 45)
    ix, iy, iw = _blas_view(n, incx), _blas_view(n, incy), _blas_view(n, incw)
    dx, dy, dw = ctx.mp_array_from_host(x), ctx.mp_array_from_host(y), ctx.mp_array_from_host(w)
    pkg.mp_waxpby(ctx, n, ctx.mp_array_from_host(al), dx, incx, ctx.mp_array_from_host(be), dy, incy, dw, incw)
    got = dw.device2host()
    want = w.copy()
    want[iw] = orc.add(orc.mul(y[iy], np.repeat(be, n)), orc.mul(x[ix], np.repeat(al, n)))
    assert diff_fields(got, want, ("digits", "sign", "exp")).size == 0
    ctx.close()


@pytest.mark.parametrize("N,full", [(8, False), (24, True)])
def test_ge_add_ge_acc_ger(pkg, N, full):
    ctx = pkg.Context(N, 0)
    orc = get_oracle(N, oracle.DEVICE)
    bits = orc.precision if full else orc.precision // 4
    m, n, lda, ldb, ldc = 19, 13, 23, 19, 21
    A, B, C = random_records(N, lda * n, bits, 51), random_records(N, ldb * n, bits, 52), random_records(N, ldc * n, bits, 53)
    al, be = random_records(N, 1, bits, 54), random_records(N, 1, bits, 55)
    dal, dbe = ctx.mp_array_from_host(al), ctx.mp_array_from_host(be)

    def sub(X, ld):      # the m x n part of a column-major array with leading dimension ld, flattened column by column
        return X.reshape(n, ld)[:, :m].reshape(-1)
    want_sum = orc.add(orc.mul(sub(B, ldb), np.repeat(be, m * n)), orc.mul(sub(A, lda), np.repeat(al, m * n)))
    # C = alpha A + beta B
    dA, dB, dC = ctx.mp_array_from_host(A), ctx.mp_array_from_host(B), ctx.mp_array_from_host(C)
    pkg.mp_ge_add(ctx, m, n, dal, dA, lda, dbe, dB, ldb, dC, ldc)
    got = dC.device2host()
    assert diff_fields(sub(got, ldc), want_sum, ("digits", "sign", "exp")).size == 0
    pad = got.reshape(n, ldc)[:, m:]
    assert diff_fields(pad.reshape(-1), C.reshape(n, ldc)[:, m:].reshape(-1)).size == 0      # outside the m x n part nothing is written
    # B = alpha A + beta B
    pkg.mp_ge_acc(ctx, m, n, dal, dA, lda, dbe, dB, ldb)
    assert diff_fields(sub(dB.device2host(), ldb), want_sum, ("digits", "sign", "exp")).size == 0
    # A = alpha x y^T + A
    x, y = random_records(N, m, bits, 56), random_records(N, 2 * n, bits, 57)
    incy = -2
    iy = _blas_view(n, incy)
    ay = orc.mul(y[iy], np.repeat(al, n))
    prod = orc.mul(np.tile(x, n), np.repeat(ay, m))          # entry (i, j) at i + j m: x_i * (alpha y_j)
    want_A = orc.add(sub(A, lda), prod)
    dA2 = ctx.mp_array_from_host(A)
    pkg.mp_ger(ctx, m, n, dal, ctx.mp_array_from_host(x), 1, ctx.mp_array_from_host(y), incy, dA2, lda)
    assert diff_fields(sub(dA2.device2host(), lda), want_A, ("digits", "sign", "exp")).size == 0
    ctx.close()
