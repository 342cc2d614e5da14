"""BASELINE config 1: mp_dot, n = 1e6, the 8-moduli / 106-bit set, host mp_float_t arithmetic (the
reference's CPU path restated in the oracle's HOST flavour), checked against exact arithmetic and the
reference's error model (tests/blas/accuracy/test_dot_accuracy.cu:39-87)."""
from fractions import Fraction

import numpy as np
import pytest

import oracle
from util import get_oracle, unit_roundoff


def _exact_dot_small_significands(x, y):
    """p/4-bit significands are below every modulus, so digits[0] IS the significand"""
    dx, dy = x["digits"][:, 0].astype(object), y["digits"][:, 0].astype(object)
    sg = np.where((x["sign"] ^ y["sign"]) == 1, -1, 1).astype(object)
    e = x["exp"].astype(np.int64) + y["exp"].astype(np.int64)
    emin = int(e.min())
    total = 0
    for p_, s_, e_ in zip(dx * dy, sg, e):
        total += (p_ * s_) << int(e_ - emin)
    return Fraction(total) * Fraction(2) ** emin


def test_dot_1e6_quarter_precision_is_exact():
    N, n = 8, 1000000
    orc = get_oracle(N, oracle.HOST)
    bits = orc.precision // 4
    x, y = orc.random_records(n, bits, 1), orc.random_records(n, bits, 2)
    assert int(x["digits"].max()) < min(orc.c["moduli"])
    want = _exact_dot_small_significands(x, y)
    r_omp, nt = orc.dot_omp(x, y)
    assert orc.to_fraction(r_omp) == want          # nothing rounds: any summation order is exact
    r_seq = orc.dot_seq(x[:200000], y[:200000])
    assert orc.to_fraction(r_seq) == _exact_dot_small_significands(x[:200000], y[:200000])
    if oracle.have_ref(N):
        r_ref, _ = oracle.RefLib(N).host_dot_omp(x, y)
        assert orc.to_fraction(r_ref) == want


def test_dot_full_precision_error_bound():
    N, n = 8, 20000
    orc = get_oracle(N, oracle.HOST)
    x, y = orc.random_records(n, orc.precision, 3), orc.random_records(n, orc.precision, 4)
    fx = [orc.to_fraction(v) for v in x]
    fy = [orc.to_fraction(v) for v in y]
    exact = sum(a * b for a, b in zip(fx, fy))
    cond = sum(abs(a * b) for a, b in zip(fx, fy))
    u = unit_roundoff(orc)
    gamma = n * u / (1 - n * u)
    got = orc.to_fraction(orc.dot_seq(x, y))
    assert abs(got - exact) <= gamma * cond
    if oracle.have_ref(N):
        ref = orc.to_fraction(oracle.RefLib(N).host_dot(x, y))
        assert ref == got
