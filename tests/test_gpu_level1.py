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
