"""Constants: product (C++ BigUInt, no GMP) == Python-integer oracle == the reference's
rns_const_init/mp_const_init output frozen in tests/golden/constants_N*.json (and, where oracle/_ref
is present, the live reference)."""
import glob
import hashlib
import json
import os

import numpy as np
import pytest

import oracle
from oracle import constants

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _sha(a):
    return hashlib.sha256(np.asarray(a, dtype=np.int64).astype(np.int32).tobytes()).hexdigest()


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "constants_N*.json"))))
def test_python_constants_match_reference_golden(path):
    g = json.load(open(path))
    c = constants.compute(g["moduli"])
    assert oracle.moduli_sets()[g["N"]] == g["moduli"]
    assert (c["log2M"], c["mp_precision"], c["mp_h"], c["mp_j"]) == (g["log2M"], g["mp_precision"], g["mp_h"], g["mp_j"])
    assert c["part_inverse"] == g["part_inverse"]
    assert c["m_pow2_residues"] == g["m_pow2_residues"]
    assert _sha(c["pow2"]) == g["pow2_sha256"]
    assert _sha(c["mi_pow2_residues"]) == g["mi_pow2_residues_sha256"]
    assert _sha(c["pow2_inverse"]) == g["pow2_inverse_sha256"]
    assert _sha(c["mrc_mult_inv"]) == g["mrc_mult_inv_sha256"]
    assert [constants.double_bits(v) for v in c["recip_rd"]] == g["recip_rd_bits"]
    assert [constants.double_bits(v) for v in c["recip_ru"]] == g["recip_ru_bits"]
    ev = [c["eval_accuracy"], c["eval_unit_low"][0], c["eval_unit_upp"][0], c["eval_inv_unit_low"][0], c["eval_inv_unit_upp"][0]]
    assert [constants.double_bits(v) for v in ev] == g["eval_doubles_bits"]
    assert [c["eval_ref_factor"], c["eval_unit_low"][1], c["eval_unit_upp"][1], c["eval_inv_unit_low"][1], c["eval_inv_unit_upp"][1]] == g["eval_ints"]


@pytest.mark.parametrize("N", sorted(oracle.moduli_sets()))
def test_product_constants_match_oracle(pkg, N):
    """the library's GMP-free derivation, through a constants-only context (no device needed)"""
    ctx = pkg.Context(N, -1)
    c = constants.compute(oracle.moduli_sets()[N])
    assert (ctx.log2M, ctx.precision, ctx.mp_h, ctx.mp_j) == (c["log2M"], c["mp_precision"], c["mp_h"], c["mp_j"])
    for which, key, dt in [(0, "moduli", np.int32), (1, "part_inverse", np.int32), (2, "pow2", np.int32), (3, "m_pow2_residues", np.int32),
                           (4, "mi_pow2_residues", np.int32), (5, "pow2_inverse", np.int32), (6, "mrc_mult_inv", np.int32),
                           (7, "recip_rd", np.float64), (8, "recip_ru", np.float64)]:
        want = np.array(c[key], dtype=np.int64 if dt == np.int32 else dt).astype(dt).reshape(-1)
        got = ctx.constant(which, dt, want.size + 8)
        assert got.size == want.size and np.array_equal(got.view(np.uint8), want.view(np.uint8)), key
    d, e = ctx.constant(9, np.float64, 5), ctx.constant(10, np.float64, 5)
    assert list(d) == [c["eval_accuracy"], c["eval_unit_low"][0], c["eval_unit_upp"][0], c["eval_inv_unit_low"][0], c["eval_inv_unit_upp"][0]]
    assert list(e) == [c["eval_ref_factor"], c["eval_unit_low"][1], c["eval_unit_upp"][1], c["eval_inv_unit_low"][1], c["eval_inv_unit_upp"][1]]
    ctx.close()


def test_bad_moduli_rejected(pkg):
    with pytest.raises(pkg.MpresError):
        pkg.Context(moduli=[4, 9, 25], device=-1)          # even modulus
    with pytest.raises(pkg.MpresError):
        pkg.Context(moduli=[15, 21, 1000003], device=-1)    # not coprime
    with pytest.raises(pkg.MpresError):
        pkg.Context(7, -1)                                  # no such predefined set
    with pytest.raises(pkg.MpresError):
        pkg.Context(moduli=[1000003, 1000033, 1000037], device=-1)   # odd count: the reference's mp_float_t gets padding there, records would not line up


def test_small_modulus_base(pkg):
    """the one-byte moduli of the tensor-core stage 2: pairwise coprime, one byte each, product about 2^362, and coprime to nothing they
    have to be (the extension to the format's moduli only needs the small moduli coprime among themselves)"""
    import math
    ctx = pkg.Context(32, -1)          # constants-only context: no device needed
    ps = ctx.small_moduli(54)
    assert ps[0] == 256 and len(set(ps)) == 54 and all(2 <= p <= 256 for p in ps)
    for i in range(54):
        for j in range(i):
            assert math.gcd(ps[i], ps[j]) == 1, (ps[i], ps[j])
    assert ctx.lib.mpres_small_modulus(ctx.h, 54) == 0 and ctx.lib.mpres_small_modulus(ctx.h, -1) == 0
    bits = sum(math.log2(p) for p in ps)
    assert 362 < bits < 363.5
    assert sorted(ps, reverse=True) == ps
    ctx.close()
