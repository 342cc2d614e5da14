"""GPU parity of the single-pass exact-window mp_gemv / mp_dot kernels (csrc/kernels_vec.cuh) through the
C-ABI: AUTO / FAST mode against the reference-order kernels, the C oracle and the reference's own CUDA
kernels (oracle/_ref), plus the fallback and the special cases."""
from fractions import Fraction

import numpy as np
import pytest

import oracle
from util import diff_fields, get_oracle, random_records, unit_roundoff

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _dot(pkg, ctx, x, y, mode):
    ctx.set_mode(mode)
    n = len(x)
    dx, dy, dr = ctx.mp_array_from_host(x), ctx.mp_array_from_host(y), ctx.mp_array_init(1)
    pkg.mp_dot(ctx, n, dx, 1, dy, 1, dr)
    return dr.device2host()[0]


def _gemv(pkg, ctx, trans, m, n, alpha, A, x, beta, y, mode, incy=1):
    ctx.set_mode(mode)
    dA, dx, dal, dbe = ctx.mp_array_from_host(A), ctx.mp_array_from_host(x), ctx.mp_array_from_host(alpha), ctx.mp_array_from_host(beta)
    leny = len(y)
    if incy != 1:
        zero = get_oracle(ctx.N).set_ints([0], [0], [0])[0]
        ys = np.full((leny - 1) * abs(incy) + 1, zero, dtype=y.dtype)
        pos = np.arange(leny) * incy if incy > 0 else (-leny + np.arange(leny) + 1) * incy
        ys[pos] = y
        dy = ctx.mp_array_from_host(ys)
        pkg.mp_gemv(ctx, trans, m, n, dal, dA, m, dx, 1, dbe, dy, incy)
        return dy.device2host()[pos]
    dy = ctx.mp_array_from_host(y)
    pkg.mp_gemv(ctx, trans, m, n, dal, dA, m, dx, 1, dbe, dy, 1)
    return dy.device2host()


@pytest.mark.parametrize("N,n", [(8, 1), (8, 5000), (16, 777), (16, 200000), (24, 3001), (32, 40000), (64, 2500)])
def test_dot_fast_bit_exact(pkg, N, n):
    """p/4-bit inputs: the one-pass exact accumulation gives the digits, sign and exponent of the reference's
    mul / round / two-pass tree sum (nothing rounds), without touching the reference-order kernels."""
    ctx = pkg.Context(N, 0)
    orc = get_oracle(N, oracle.DEVICE)
    bits = orc.precision // 4
    x = random_records(N, n, bits, 111)
    y = random_records(N, n, bits, 112)
    got = _dot(pkg, ctx, x, y, pkg.MODE_AUTO)
    assert ctx.last_fallback_count() == 0
    want = _dot(pkg, ctx, x, y, pkg.MODE_REFERENCE_ORDER)
    assert diff_fields(np.array([got]), np.array([want]), ("digits", "sign", "exp")).size == 0, (got, want)
    if n <= 5000:
        w2 = orc.dot_seq(x, y)
        assert diff_fields(np.array([got]), np.array([w2]), ("digits", "sign", "exp")).size == 0
    if oracle.have_ref(N) and n <= 200000:
        r, _ = oracle.RefLib(N, gpu=True).gpu_dot(x, y)
        assert diff_fields(np.array([got]), np.array([r]), ("digits", "sign", "exp")).size == 0
    # the interval evaluation must enclose the value
    v = orc.to_fraction(got)
    M = orc.c["M"]
    lo = Fraction(float(got["eval"]["frac"][0])) * Fraction(2) ** int(got["eval"]["exp"][0])
    up = Fraction(float(got["eval"]["frac"][1])) * Fraction(2) ** int(got["eval"]["exp"][1])
    X = abs(v) / Fraction(2) ** int(got["exp"])
    assert lo <= X / M <= up
    ctx.close()


@pytest.mark.parametrize("N,m,n", [(8, 70, 50), (8, 300, 1000), (16, 129, 257), (16, 1000, 64), (24, 50, 90), (32, 64, 700), (64, 40, 33), (16, 3, 5)])
@pytest.mark.parametrize("trans", [111, 112])
def test_gemv_fast_bit_exact(pkg, N, m, n, trans):
    ctx = pkg.Context(N, 0)
    orc = get_oracle(N, oracle.DEVICE)
    bits = orc.precision // 4
    A = random_records(N, m * n, bits, 121)
    alpha = random_records(N, 1, bits, 122)
    beta = random_records(N, 1, bits, 123)
    lenx, leny = (n, m) if trans == 111 else (m, n)
    x = random_records(N, lenx, bits, 124)
    y = random_records(N, leny, bits, 125)
    got = _gemv(pkg, ctx, trans, m, n, alpha, A, x, beta, y, pkg.MODE_AUTO)
    assert ctx.last_fallback_count() == 0
    want = _gemv(pkg, ctx, trans, m, n, alpha, A, x, beta, y, pkg.MODE_REFERENCE_ORDER)
    bad = diff_fields(got, want, ("digits", "sign", "exp"))
    assert bad.size == 0, "%d/%d outputs differ from the reference order, first %d\n%s\n%s" % (bad.size, leny, bad[0], got[bad[0]], want[bad[0]])
    if m * n <= 40000:
        w2 = orc.gemv(trans, m, n, alpha, A, x, beta, y)
        assert diff_fields(got, w2, ("digits", "sign", "exp")).size == 0
        if oracle.have_ref(N):
            r, _ = oracle.RefLib(N, gpu=True).gpu_gemv(trans, m, n, alpha, A, x, beta, y)
            assert diff_fields(got, r, ("digits", "sign", "exp")).size == 0
    ctx.close()


@pytest.mark.parametrize("trans", [111, 112])
def test_gemv_fast_strided_y(pkg, trans):
    N, m, n = 16, 90, 70
    ctx = pkg.Context(N, 0)
    orc = get_oracle(N, oracle.DEVICE)
    bits = orc.precision // 4
    A = random_records(N, m * n, bits, 131)
    alpha = random_records(N, 1, bits, 132)
    beta = random_records(N, 1, bits, 133)
    lenx, leny = (n, m) if trans == 111 else (m, n)
    x = random_records(N, lenx, bits, 134)
    y = random_records(N, leny, bits, 135)
    want = _gemv(pkg, ctx, trans, m, n, alpha, A, x, beta, y, pkg.MODE_REFERENCE_ORDER)
    for incy in (2, -3):
        got = _gemv(pkg, ctx, trans, m, n, alpha, A, x, beta, y, pkg.MODE_AUTO, incy=incy)
        assert diff_fields(got, want, ("digits", "sign", "exp")).size == 0, incy
    ctx.close()


def _special_vec_inputs(N, m, n, bits, seed):
    """zeros, rows / columns scaled by powers of two (alignment shifts), an exactly cancelling row.  The scalings
    stay below bits / 6 so that the sums still fit the working precision where the test demands bit-exactness."""
    orc = get_oracle(N, oracle.DEVICE)
    A = random_records(N, m * n, bits, seed).reshape(n, m).copy()   # A[j, i]
    x = random_records(N, n, bits, seed + 1)
    zero = orc.set_ints([0], [0], [0])[0]
    k = max(3, bits // 6)
    A[:, 1] = zero              # zero row
    A[3, :] = zero              # zero column
    A[5, 2] = zero
    x[7] = zero
    A[:, 4]["exp"] += k
    A[:, 6]["exp"] -= k + 1
    A[9, :]["exp"] += k + 2
    x[11]["exp"] -= k + 3
    A[1, 8] = A[0, 8]           # row 8: x0 a - x0' a with x1 := x0 -> cancels exactly
    A[1, 8]["sign"] ^= 1
    A[2:, 8] = zero
    x[1] = x[0]
    return A.reshape(-1), x


@pytest.mark.parametrize("N", [8, 16, 32])
def test_gemv_fast_special_cases_value_exact(pkg, N):
    """Where exact zeros take part the reference's exponent depends on the zero's own exponent (DESIGN section 7
    "zeros"); the value is the same.  Everything else must stay bit-identical."""
    m, n = 40, 30
    ctx = pkg.Context(N, 0)
    orc = get_oracle(N, oracle.DEVICE)
    bits = orc.precision // 4
    A, x = _special_vec_inputs(N, m, n, bits, 141)
    alpha = random_records(N, 1, bits, 144)
    beta = random_records(N, 1, bits, 145)
    y = random_records(N, m, bits, 146)
    got = _gemv(pkg, ctx, 111, m, n, alpha, A, x, beta, y, pkg.MODE_AUTO)
    assert ctx.last_fallback_count() == 0
    want = _gemv(pkg, ctx, 111, m, n, alpha, A, x, beta, y, pkg.MODE_REFERENCE_ORDER)
    u = unit_roundoff(orc)

    def same(g, w):
        # N = 8: the far-scaled terms make partial sums exceed the working precision, the reference order rounds
        # (truncates) them step by step while the fast path rounds the exact sum once -- a few units apart
        if N >= 16:
            return orc.to_fraction(g) == orc.to_fraction(w)
        return abs(orc.to_fraction(g) - orc.to_fraction(w)) <= 4 * max(m, n) * u * abs(orc.to_fraction(w))
    for i in range(m):
        assert same(got[i], want[i]), i
    # transposed: the same matrix seen as n x m^T
    yt = random_records(N, n, bits, 147)
    xt = random_records(N, m, bits, 148)
    got = _gemv(pkg, ctx, 112, m, n, alpha, A, xt, beta, yt, pkg.MODE_AUTO)
    want = _gemv(pkg, ctx, 112, m, n, alpha, A, xt, beta, yt, pkg.MODE_REFERENCE_ORDER)
    for j in range(n):
        assert same(got[j], want[j]), j
    # alpha == 0 / beta == 0
    zero = orc.set_ints([0], [0], [0])
    for al, be in ((alpha, zero), (zero, beta)):
        got = _gemv(pkg, ctx, 111, m, n, al, A, x, be, y, pkg.MODE_AUTO)
        want = _gemv(pkg, ctx, 111, m, n, al, A, x, be, y, pkg.MODE_REFERENCE_ORDER)
        for i in range(m):
            assert same(got[i], want[i]), i
    ctx.close()


@pytest.mark.parametrize("N", [8, 32])
def test_vec_full_precision_inputs_fall_back(pkg, N):
    """p-bit inputs: the exact sums do not fit below M/4, so AUTO must hand every output back to the
    reference-order kernels -- bit-identical records including the interval evaluations."""
    ctx = pkg.Context(N, 0)
    orc = get_oracle(N, oracle.DEVICE)
    bits = orc.precision
    n = 3000
    x = random_records(N, n, bits, 151)
    y = random_records(N, n, bits, 152)
    got = _dot(pkg, ctx, x, y, pkg.MODE_AUTO)
    assert ctx.last_fallback_count() == 1
    want = _dot(pkg, ctx, x, y, pkg.MODE_REFERENCE_ORDER)
    assert diff_fields(np.array([got]), np.array([want])).size == 0
    m, nn = 33, 47
    A = random_records(N, m * nn, bits, 153)
    alpha = random_records(N, 1, bits, 154)
    beta = random_records(N, 1, bits, 155)
    for trans in (111, 112):
        lenx, leny = (nn, m) if trans == 111 else (m, nn)
        xv = random_records(N, lenx, bits, 156)
        yv = random_records(N, leny, bits, 157)
        got = _gemv(pkg, ctx, trans, m, nn, alpha, A, xv, beta, yv, pkg.MODE_AUTO)
        assert ctx.last_fallback_count() == leny
        want = _gemv(pkg, ctx, trans, m, nn, alpha, A, xv, beta, yv, pkg.MODE_REFERENCE_ORDER)
        assert diff_fields(got, want).size == 0
    ctx.close()


def test_vec_mixed_width_accuracy(pkg):
    """~p/3-bit inputs with a wide exponent spread: some partial sums round in the reference, the fast path rounds
    once.  Error model of tests/blas/accuracy/test_dot_accuracy.cu:41-72."""
    N = 8
    ctx = pkg.Context(N, 0)
    orc = get_oracle(N, oracle.DEVICE)
    bits = orc.precision // 3
    n = 4000
    x = random_records(N, n, bits, 161)
    y = random_records(N, n, bits, 162)
    got = _dot(pkg, ctx, x, y, pkg.MODE_AUTO)
    fx = [orc.to_fraction(v) for v in x]
    fy = [orc.to_fraction(v) for v in y]
    exact = sum(a * b for a, b in zip(fx, fy))
    u = unit_roundoff(orc)
    gam = n * u / (1 - n * u)
    assert abs(orc.to_fraction(got) - exact) <= gam * sum(abs(a * b) for a, b in zip(fx, fy))
    ctx.close()


def test_dot_fast_collection_and_partial(pkg):
    """mp_collection_t operands and the packed partial used by the multi-GPU DOT take the same fast path"""
    N, n = 16, 9000
    ctx = pkg.Context(N, 0)
    orc = get_oracle(N, oracle.DEVICE)
    bits = orc.precision // 4
    x = random_records(N, n, bits, 171)
    y = random_records(N, n, bits, 172)
    want = _dot(pkg, ctx, x, y, pkg.MODE_REFERENCE_ORDER)
    ctx.set_mode(pkg.MODE_AUTO)
    cx, cy = ctx.mp_collection_from_host(x), ctx.mp_collection_from_host(y)
    cr = pkg.MpCollection(ctx, 1)
    pkg.mp_dot(ctx, n, cx, 1, cy, 1, cr)
    got = cr.device2host()[0]
    assert diff_fields(np.array([got]), np.array([want]), ("digits", "sign", "exp")).size == 0
    ctx.close()
