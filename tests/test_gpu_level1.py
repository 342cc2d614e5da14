"""The level-1 / elementwise operations (mp_scal, mp_axpy, mp_waxpby, mp_axpy_dot, mp_rot, mp_ge_add, mp_ge_acc, mp_ger,
mp_ge_diag_scale, mp_ge_lr_scale) through the C-ABI against the C oracle's mp_mul / mp_add (the DEVICE flavour, itself pinned
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


@pytest.mark.parametrize("N,full,incd", [(8, False, 1), (8, True, -1), (32, True, 2), (16, True, 1)])
def test_ge_diag_scale_lr_scale(pkg, N, full, incd):
    """cuda::mp_ge_diag_scale (both sides) and cuda::mp_ge_lr_scale against the oracle's mp_mul, entry by entry (row i times d_i / column j
    times d_j, rounded after every product: src/blas/gediagscale.cuh:75-98, gelrscale.cuh:75-92)"""
    ctx = pkg.Context(N, 0)
    orc = get_oracle(N, oracle.DEVICE)
    bits = orc.precision if full else orc.precision // 4
    m, n, lda = 37, 21, 40
    A = random_records(N, lda * n, bits, 61)
    DL = random_records(N, m * abs(incd), bits, 62)
    DR = random_records(N, n * abs(incd), bits, 63)
    il, ir = _blas_view(m, incd), _blas_view(n, incd)
    zero = orc.set_ints([0], [0], [0])[0]
    DL[il[3]] = zero
    A[5 + 2 * lda] = zero
    rows = np.array([i + j * lda for j in range(n) for i in range(m)])
    dl_e = DL[[il[i] for j in range(n) for i in range(m)]]
    dr_e = DR[[ir[j] for j in range(n) for i in range(m)]]
    dDL, dDR = ctx.mp_array_from_host(DL), ctx.mp_array_from_host(DR)
    # left
    dA = ctx.mp_array_from_host(A)
    pkg.mp_ge_diag_scale(ctx, pkg.LEFT_SIDE, m, n, dDL, incd, dA, lda)
    want = A.copy(); want[rows] = orc.mul(A[rows], dl_e)
    got_l = dA.device2host()
    assert diff_fields(got_l, want, ("digits", "sign", "exp")).size == 0
    # right
    dA = ctx.mp_array_from_host(A)
    pkg.mp_ge_diag_scale(ctx, pkg.RIGHT_SIDE, m, n, dDR, incd, dA, lda)
    want = A.copy(); want[rows] = orc.mul(A[rows], dr_e)
    assert diff_fields(dA.device2host(), want, ("digits", "sign", "exp")).size == 0
    # both
    dA = ctx.mp_array_from_host(A)
    pkg.mp_ge_lr_scale(ctx, m, n, dDL, incd, dDR, incd, dA, lda)
    want = A.copy(); want[rows] = orc.mul(orc.mul(A[rows], dl_e), dr_e)
    got = dA.device2host()
    bad = diff_fields(got, want, ("digits", "sign", "exp"))
    assert bad.size == 0, "%d differ, first %d" % (bad.size, bad[0])
    # the padding rows of A (m <= i < lda) are never touched
    pad = np.array([i + j * lda for j in range(n) for i in range(m, lda)])
    assert diff_fields(got[pad], A[pad]).size == 0
    # silent returns of the reference: nothing is written
    pkg.mp_ge_lr_scale(ctx, 0, n, dDL, incd, dDR, incd, dA, lda)
    pkg.mp_ge_diag_scale(ctx, pkg.LEFT_SIDE, m, 0, dDL, incd, dA, lda)
    with pytest.raises(Exception):
        pkg.mp_ge_diag_scale(ctx, pkg.LEFT_SIDE, m, n, dDL, 0, dA, lda)
    with pytest.raises(Exception):
        pkg.mp_ge_lr_scale(ctx, m, n, dDL, incd, dDR, incd, dA, m - 1)
    assert diff_fields(dA.device2host(), got).size == 0
    ctx.close()


@pytest.mark.parametrize("N,full,incx,incy", [(8, False, 1, 1), (8, True, 1, 1), (32, True, 1, 1), (16, True, 2, -1), (24, True, -1, 3)])
def test_rot(pkg, N, full, incx, incy):
    """cuda::mp_rot against the oracle: x = round(round(c x) + round(s y)), y = round(round(c y) - round(s x)) (src/blas/rot.cuh:60-99).  With
    non-unit increments the reference scales the first n contiguous elements by c (its mp_scal(n, c, x, 1) calls) -- reproduced as is."""
    ctx = pkg.Context(N, 0)
    orc = get_oracle(N, oracle.DEVICE)
    bits = orc.precision if full else orc.precision // 4
    n = 211
    x = random_records(N, n * abs(incx), bits, 71)
    y = random_records(N, n * abs(incy), bits, 72)
    c, s = random_records(N, 1, bits, 73), random_records(N, 1, bits, 74)
    ix, iy = _blas_view(n, incx), _blas_view(n, incy)
    zero = orc.set_ints([0], [0], [0])[0]
    x[ix[4]] = zero
    y[iy[6]] = zero
    x[ix[8]] = zero; y[iy[8]] = zero
    cn, sn = np.repeat(c, n), np.repeat(s, n)
    b1, b2 = orc.mul(x[ix], sn), orc.mul(y[iy], sn)
    wx, wy = x.copy(), y.copy()
    wx[:n] = orc.mul(x[:n], cn)
    wy[:n] = orc.mul(y[:n], cn)
    nb1 = b1.copy(); nb1["sign"] ^= 1
    wx[ix] = orc.add(wx[ix], b2)
    wy[iy] = orc.add(wy[iy], nb1)
    dx, dy, dc, ds = ctx.mp_array_from_host(x), ctx.mp_array_from_host(y), ctx.mp_array_from_host(c), ctx.mp_array_from_host(s)
    buf1 = buf2 = None
    if incx != 1 or incy != 1:
        buf1, buf2 = ctx.mp_array_from_host(random_records(N, n, bits, 75)), ctx.mp_array_from_host(random_records(N, n, bits, 76))
    pkg.mp_rot(ctx, n, dx, incx, dy, incy, dc, ds, buf1, buf2)
    gx, gy = dx.device2host(), dy.device2host()
    bad = diff_fields(gx, wx, ("digits", "sign", "exp"))
    assert bad.size == 0, "x: %d differ, first %d\n%s\n%s" % (bad.size, bad[0], gx[bad[0]], wx[bad[0]])
    bad = diff_fields(gy, wy, ("digits", "sign", "exp"))
    assert bad.size == 0, "y: %d differ, first %d\n%s\n%s" % (bad.size, bad[0], gy[bad[0]], wy[bad[0]])
    pkg.mp_rot(ctx, 0, dx, incx, dy, incy, dc, ds, buf1, buf2)
    assert diff_fields(dx.device2host(), gx).size == 0
    ctx.close()


@pytest.mark.parametrize("N,incw,incv,incu", [(8, 1, 1, 1), (32, 1, 1, 1), (16, 2, -1, 3)])
def test_axpy_dot(pkg, N, incw, incv, incu):
    """cuda::mp_axpy_dot: w = round(w - round(alpha v)) record for record against the oracle, then r = u^T w (p/8-bit w, v, alpha and
    p/4-bit u: neither the products nor the sums round, so every summation order gives the oracle's sequential dot product)"""
    ctx = pkg.Context(N, 0)
    orc = get_oracle(N, oracle.DEVICE)
    bits = orc.precision // 8
    n = 1500
    w = random_records(N, n * abs(incw), bits, 81)
    v = random_records(N, n * abs(incv), bits, 82)
    u = random_records(N, n * abs(incu), 2 * bits, 83)
    al = random_records(N, 1, bits, 84)
    iw, iv, iu = _blas_view(n, incw), _blas_view(n, incv), _blas_view(n, incu)
    prod = orc.mul(v[iv], np.repeat(al, n))
    w[iw[3]] = prod[3]                       # exact cancellation
    prod["sign"] ^= 1
    want_w = w.copy()
    want_w[iw] = orc.add(w[iw], prod)
    assert orc.to_fraction(want_w[iw[3]]) == 0
    want_r = orc.dot_seq(u[iu], want_w[iw])
    dw, dv, du, dal, dr = (ctx.mp_array_from_host(w), ctx.mp_array_from_host(v), ctx.mp_array_from_host(u), ctx.mp_array_from_host(al),
                           ctx.mp_array_init(1))
    pkg.mp_axpy_dot(ctx, n, dal, dw, incw, dv, incv, du, incu, dr)
    got_w, got_r = dw.device2host(), dr.device2host()
    bad = diff_fields(got_w, want_w, ("digits", "sign", "exp"))
    assert bad.size == 0, "w: %d differ, first %d" % (bad.size, bad[0])
    assert diff_fields(got_r, np.array([want_r]), ("digits", "sign", "exp")).size == 0
    pkg.mp_axpy_dot(ctx, 0, dal, dw, incw, dv, incv, du, incu, dr)
    assert diff_fields(dw.device2host(), got_w).size == 0
    ctx.close()
