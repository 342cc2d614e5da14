"""Shared helpers for the parity tests (test infrastructure)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from oracle import gen  # noqa: E402

_ORACLES = {}


def get_oracle(N, flavor=oracle.DEVICE):
    key = (N, flavor)
    if key not in _ORACLES:
        _ORACLES[key] = oracle.Oracle(N, flavor)
    return _ORACLES[key]


def random_records(N, count, bits, seed):
    """tsthelper-style uniform(-1,1) values with `bits`-bit significands as mp_float_t records"""
    orc = get_oracle(N)
    s, m, e = gen.random_values(count, bits, seed)
    return orc.set_ints(s, m, e)


def diff_fields(a, b, fields=("digits", "sign", "exp", "eval")):
    """indices where records differ in any of `fields`"""
    bad = np.zeros(a.shape, dtype=bool)
    for f in fields:
        if f == "digits":
            bad |= (a["digits"] != b["digits"]).any(axis=-1)
        elif f == "eval":
            bad |= (a["eval"]["frac"].view(np.int64) != b["eval"]["frac"].view(np.int64)).any(axis=-1)
            bad |= (a["eval"]["exp"] != b["eval"]["exp"]).any(axis=-1)
        else:
            bad |= a[f] != b[f]
    return np.nonzero(bad.reshape(-1))[0]


def value(orc, rec):
    return orc.to_fraction(rec)


def rel_err(orc, got, want):
    """|got - want| / |want| as a float (exact rational arithmetic underneath)"""
    g, w = value(orc, got), value(orc, want)
    if w == 0:
        return float(abs(g))
    return float(abs(g - w) / abs(w))


def unit_roundoff(orc):
    """u = 4 / sqrt(M) (reference tests/blas/accuracy/test_dot_accuracy.cu:41-44)"""
    from fractions import Fraction
    import math
    M = orc.c["M"]
    return Fraction(4, math.isqrt(M))
