"""The C oracle (HOST flavour) against golden vectors minted from the reference's own host functions
(tests/golden/make_golden.py): bit-exact records, interval evaluations included."""
import glob
import os

import numpy as np
import pytest

import oracle
from util import diff_fields, get_oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FILES = sorted(glob.glob(os.path.join(GOLD, "ref_host_N*.npz")))


def _recs(orc, raw):
    return np.frombuffer(raw.tobytes(), dtype=orc.dtype).copy()


@pytest.mark.parametrize("path", FILES)
@pytest.mark.parametrize("tag", ["q", "f"])
def test_oracle_reproduces_reference_host(path, tag):
    N = int(os.path.basename(path)[len("ref_host_N"):-4])
    g = np.load(path)
    orc = get_oracle(N, oracle.HOST)
    mants = [int.from_bytes(bytes(v).ljust(256, b"\0"), "little") for v in g[tag + "_mant"]]
    recs = orc.set_ints(g[tag + "_sign"], mants, g[tag + "_exp"])
    want = _recs(orc, g[tag + "_recs"])
    assert diff_fields(recs, want).size == 0, "mp_set_mpfr"
    x, y = want[:48], want[48:]
    mul = orc.mul(x, y)
    assert diff_fields(mul, _recs(orc, g[tag + "_mul"])).size == 0, "mp_mul"
    assert diff_fields(orc.add(x, y), _recs(orc, g[tag + "_add"])).size == 0, "mp_add"
    assert diff_fields(orc.add(mul, x), _recs(orc, g[tag + "_add_mixed"])).size == 0, "mp_add (mixed exponents)"
    assert diff_fields(orc.round(x, g[tag + "_round_bits"]), _recs(orc, g[tag + "_round"])).size == 0, "mp_round"
    assert diff_fields(orc.eval(mul), _recs(orc, g[tag + "_eval"])).size == 0, "rns_eval_compute"
    assert diff_fields(orc.eval(x, fast=True), _recs(orc, g[tag + "_eval_fast"])).size == 0, "rns_eval_compute_fast"
    assert diff_fields(np.array([orc.dot_seq(x, y)]), _recs(orc, g[tag + "_dot"])).size == 0, "dot"
    mm, nn, kk = 5, 4, 6
    A, B, C = want[:mm * kk], want[30:30 + kk * nn], want[60:60 + mm * nn]
    Cg, _ = orc.gemm(mm, nn, kk, want[90:91], A, B, want[91:92], C)
    assert diff_fields(Cg, _recs(orc, g[tag + "_gemm"])).size == 0, "gemm (host semantics)"


@pytest.mark.parametrize("N", [8, 32])
def test_device_flavour_agrees_in_value(N):
    """DEVICE flavour (IEEE directed rounding, exact %) differs from HOST only in the last bits of the
    interval bounds; values (digits, sign, exp) agree wherever no rounding decision sits on a boundary"""
    g = np.load(os.path.join(GOLD, "ref_host_N%d.npz" % N))
    h, d = get_oracle(N, oracle.HOST), get_oracle(N, oracle.DEVICE)
    for tag in ("q", "f"):
        recs = _recs(h, g[tag + "_recs"])
        x, y = recs[:48], recs[48:]
        assert diff_fields(h.mul(x, y), d.mul(x, y), ("digits", "sign", "exp")).size == 0
        assert diff_fields(h.add(x, y), d.add(x, y), ("digits", "sign", "exp")).size == 0
