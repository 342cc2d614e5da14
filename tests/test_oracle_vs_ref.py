"""Where oracle/_ref exists (the build container, and the GPU box through the snapshot): larger random
bit-exact comparison of the C oracle's HOST flavour with the unmodified reference host functions."""
import numpy as np
import pytest

import oracle
from oracle import gen
from util import diff_fields, get_oracle


@pytest.mark.parametrize("N", [8, 16, 24, 32])
def test_host_flavour_bit_exact_against_reference(N):
    if not oracle.have_ref(N):
        pytest.skip("oracle/_ref/libmpres_ref_N%d.so not built" % N)
    ref = oracle.RefLib(N)
    orc = get_oracle(N, oracle.HOST)
    p = orc.precision
    for bits in (p // 4, p // 2, p):
        s, m, e = gen.random_values(1200, bits, 99 + bits)
        xr = ref.set_ints(s, m, e, bits)
        assert diff_fields(xr, orc.set_ints(s, m, e)).size == 0
        x, y = xr[:600], xr[600:]
        mr = ref.host_mul(x, y)
        assert diff_fields(mr, orc.mul(x, y)).size == 0
        assert diff_fields(ref.host_add(x, y), orc.add(x, y)).size == 0
        assert diff_fields(ref.host_add(mr, x), orc.add(mr, x)).size == 0
        assert diff_fields(np.array([ref.host_dot(x, y)]), np.array([orc.dot_seq(x, y)])).size == 0
    A, B, C = xr[:7 * 9], xr[100:100 + 9 * 5], xr[200:200 + 7 * 5]
    cr, _ = ref.host_gemm(7, 5, 9, xr[300:301], A, B, xr[301:302], C)
    co, _ = orc.gemm(7, 5, 9, xr[300:301], A, B, xr[301:302], C)
    assert diff_fields(cr, co).size == 0
    yr, _ = ref.host_gemv(7, 9, xr[300:301], A, xr[400:409], xr[301:302], xr[500:507])
    # host gemv helper sums sequentially; the oracle's v1 structure with one "thread" is the same order
    yo = orc.gemv(111, 7, 9, xr[300:301], A, xr[400:409], xr[301:302], xr[500:507], block=1)
    assert diff_fields(yr, yo, ("digits", "sign", "exp")).size == 0
