"""GPU parity of mp_gemm / mp_gemv / mp_dot through the C-ABI against the reference's CUDA kernels
(oracle/_ref) and the C oracle."""
import numpy as np
import pytest

import oracle
from util import diff_fields, get_oracle, random_records, unit_roundoff

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _gemm(pkg, ctx, m, n, k, alpha, A, B, beta, C, mode, ta=111, tb=111):
    ctx.set_mode(mode)
    dA, dB, dC = ctx.mp_array_from_host(A), ctx.mp_array_from_host(B), ctx.mp_array_from_host(C)
    dal, dbe = ctx.mp_array_from_host(alpha), ctx.mp_array_from_host(beta)
    lda = m if ta == 111 else k
    ldb = k if tb == 111 else n
    pkg.mp_gemm(ctx, ta, tb, m, n, k, dal, dA, lda, dB, ldb, dbe, dC, m)
    return dC.device2host()


@pytest.mark.parametrize("N,full", [(8, False), (8, True), (16, False), (32, False), (32, True)])
def test_gemm_reference_order_bit_exact(pkg, N, full):
    """REFERENCE_ORDER mode reproduces the reference v1 mp_gemm bit for bit (digits, sign, exp, eval)"""
    ctx = pkg.Context(N, 0)
    orc = get_oracle(N, oracle.DEVICE)
    bits = orc.precision if full else orc.precision // 4
    m, n, k = 33, 17, 41
    A = random_records(N, m * k, bits, 1)
    B = random_records(N, k * n, bits, 2)
    C = random_records(N, m * n, bits, 3)
    alpha = random_records(N, 1, bits, 4)
    beta = random_records(N, 1, bits, 5)
    got = _gemm(pkg, ctx, m, n, k, alpha, A, B, beta, C, pkg.MODE_REFERENCE_ORDER)
    want, _ = orc.gemm(m, n, k, alpha, A, B, beta, C)
    bad = diff_fields(got, want, ("digits", "sign", "exp"))
    assert bad.size == 0, "%d/%d elements differ from the oracle" % (bad.size, m * n)
    if oracle.have_ref(N):
        ref = oracle.RefLib(N, gpu=True)
        r, ab, _ = ref.gpu_gemm(m, n, k, alpha, A, B, beta, C, want_ab=True)
        bad = diff_fields(got, r)
        assert bad.size == 0, "%d/%d elements differ from the reference kernels\n%s\n%s" % (bad.size, m * n, got[bad[0]], r[bad[0]])
    ctx.close()


@pytest.mark.parametrize("N", [8, 16])
def test_dot_and_gemv_reference_order(pkg, N):
    ctx = pkg.Context(N, 0)
    ctx.set_mode(pkg.MODE_REFERENCE_ORDER)
    orc = get_oracle(N, oracle.DEVICE)
    bits = orc.precision // 4
    n = 5000
    x = random_records(N, n, bits, 11)
    y = random_records(N, n, bits, 12)
    dx, dy, dr = ctx.mp_array_from_host(x), ctx.mp_array_from_host(y), ctx.mp_array_init(1)
    pkg.mp_dot(ctx, n, dx, 1, dy, 1, dr)
    got = dr.device2host()[0]
    want = orc.dot_seq(x, y)
    # p/4-bit inputs: nothing rounds, so every summation order gives the same digits/sign/exp
    assert diff_fields(np.array([got]), np.array([want]), ("digits", "sign", "exp")).size == 0
    if oracle.have_ref(N):
        r, _ = oracle.RefLib(N, gpu=True).gpu_dot(x, y)
        assert diff_fields(np.array([got]), np.array([r]), ("digits", "sign", "exp")).size == 0
    m, nn = 70, 50
    A = random_records(N, m * nn, bits, 13)
    alpha = random_records(N, 1, bits, 14)
    beta = random_records(N, 1, bits, 15)
    for trans in (111, 112):
        lenx, leny = (nn, m) if trans == 111 else (m, nn)
        xv = random_records(N, lenx, bits, 16)
        yv = random_records(N, leny, bits, 17)
        dA, dxv, dyv = ctx.mp_array_from_host(A), ctx.mp_array_from_host(xv), ctx.mp_array_from_host(yv)
        dal, dbe = ctx.mp_array_from_host(alpha), ctx.mp_array_from_host(beta)
        pkg.mp_gemv(ctx, trans, m, nn, dal, dA, m, dxv, 1, dbe, dyv, 1)
        got = dyv.device2host()
        want = orc.gemv(trans, m, nn, alpha, A, xv, beta, yv)
        assert diff_fields(got, want, ("digits", "sign", "exp")).size == 0
        if oracle.have_ref(N):
            r, _ = oracle.RefLib(N, gpu=True).gpu_gemv(trans, m, nn, alpha, A, xv, beta, yv)
            assert diff_fields(got, r, ("digits", "sign", "exp")).size == 0
    ctx.close()
