"""oracle/constants.py -- TEST INFRASTRUCTURE ONLY (checker, never on the product path).

Exact restatement, with Python integers and fractions, of every precomputed constant of the
reference: rns_const_init (/root/reference/src/rns.cuh:324-442) and mp_const_init
(/root/reference/src/arith/arith_utils.cuh:44-85).  The reference derives them with GMP/MPFR; nothing
here needs either.  Pinned bit-for-bit against the reference itself (oracle/_ref) by
tests/test_constants.py and frozen in tests/golden/constants_N*.json.
"""
from fractions import Fraction
import math
import struct

# The predefined "n-double" moduli sets (/root/reference/src/params/32-bit-n-double-moduli/
# params.<N>_<n>double.h, macro RNS_MODULI_VALUES).  They are data of the file format, like a CRC
# polynomial: a drop-in replacement has to use exactly these numbers.
MODULI_SETS = {
    8: [113812103, 113812105, 113812107, 113812109, 113812111, 113812117, 113812121, 113812123],
}

RNS_P2_SCALING_THRESHOLD = 30          # params.*.h: RNS_P2_SCALING_THRESHOLD
RNS_EVAL_RELATIVE_ERROR = 0.0000001    # params.*.h: RNS_EVAL_RELATIVE_ERROR
MP_EXP_MAX = (2 ** 31 - 1) // 2        # types.cuh:35


def _round_frac_2exp(v: Fraction, up: bool):
    """mpfr_get_d_2exp(v, RNDU/RNDD): v = frac * 2**exp with frac in [0.5, 1), 53-bit frac."""
    assert v > 0
    e = v.numerator.bit_length() - v.denominator.bit_length()
    while Fraction(2) ** e <= v:
        e += 1
    while Fraction(2) ** (e - 1) > v:
        e -= 1
    scaled = v * Fraction(2) ** (53 - e)            # in [2^52, 2^53)
    n = math.ceil(scaled) if up else math.floor(scaled)
    if n == 2 ** 53:                                 # carry: MPFR returns 0.5 * 2**(e+1)
        return 0.5, e + 1
    return n / 2.0 ** 53, e


def _pred_succ_div(a: float, b: float):
    """dinterval.cuh:142-154 -- the host-emulated ddiv_rd / ddiv_ru (bitwise.cuh:51-61 constants)."""
    eps = math.ldexp(1.0, -53)
    phi1 = eps * (1 + 2 * eps)
    eta = math.ldexp(1.0, -1074)
    c = a / b
    e = phi1 * abs(c) + eta
    return c - e, c + e


def compute(moduli):
    N = len(moduli)
    M = 1
    for m in moduli:
        M *= m
    L = M.bit_length()
    log2M = L - 1                                    # RNS_MODULI_PRODUCT_LOG2 (floor(log2 M))
    # rns.cuh:336 -- mpfr_init() default precision is 53 bits, RNDD
    Md = (M >> (L - 53)) << (L - 53) if L > 53 else M
    Mi = [M // m for m in moduli]
    c = {"N": N, "moduli": list(moduli), "log2M": log2M, "M": M}
    c["part_inverse"] = [pow(Mi[i] % moduli[i], -1, moduli[i]) for i in range(N)]          # :343-346
    c["pow2"] = [[pow(2, j, m) for m in moduli] for j in range(log2M + 1)]                  # :352-358
    T = RNS_P2_SCALING_THRESHOLD
    c["m_pow2_residues"] = [M % (1 << (j + 1)) for j in range(T)]                           # :362-365
    c["mi_pow2_residues"] = [[Mi[i] % (1 << (j + 1)) for i in range(N)] for j in range(T)]  # :367-372
    c["pow2_inverse"] = [[pow(pow(2, j + 1, m), -1, m) for m in moduli] for j in range(T)]  # :374-382
    eps = RNS_EVAL_RELATIVE_ERROR
    acc = 4 * math.pow(2.0, 1 - 53) * N * math.log2(float(N)) * (1 + eps / 2) / eps         # :385
    c["eval_accuracy"] = acc
    c["eval_ref_factor"] = math.floor(math.log2(1 / (2 * acc)))                             # :387
    inv = Fraction(1, Md)
    c["eval_unit_upp"] = _round_frac_2exp(inv, True)                                        # :394-396
    c["eval_unit_low"] = _round_frac_2exp(inv, False)                                       # :402-404
    # :398-399 / :406-407 -- at 10000-bit working precision 1 - 1/M is inexact in both directions, and
    # the final 53-bit rounding gives 1.0 (-> 0.5 * 2^1) upwards and 1 - 2^-53 downwards.
    c["eval_inv_unit_upp"] = _round_frac_2exp(1 - inv, True)
    c["eval_inv_unit_low"] = _round_frac_2exp(1 - inv, False)
    rd, ru = [], []
    for m in moduli:                                                                       # :412-415
        lo, hi = _pred_succ_div(1.0, float(m))
        rd.append(lo)
        ru.append(hi)
    c["recip_rd"], c["recip_ru"] = rd, ru
    c["mrc_mult_inv"] = [[pow(moduli[i], -1, moduli[j]) if j > i else 0 for j in range(N)]
                         for i in range(N)]                                                 # :417-426
    # mp_const_init, arith_utils.cuh:44-61 (all RNDD on the 53-bit Md)
    h = 0
    while (1 << (2 * h)) < Md:
        h += 1                                       # smallest h with 4^h >= Md  == ceil(log2(Md)/2)
    c["mp_precision"] = (Md.bit_length() - 1) // 2 - 1
    c["mp_h"] = -h
    # log2((Md - sqrt(Md)) / Md) = log2(1 - Md^-1/2) in (-1, 0) for every Md >= 5
    c["mp_j"] = -1
    return c


def double_bits(x: float) -> int:
    return struct.unpack("<q", struct.pack("<d", x))[0]


def read_moduli_sets(path="/root/reference/src/params/32-bit-n-double-moduli"):
    """Parse RNS_MODULI_VALUES out of the reference's params files (container only)."""
    import glob
    import os
    import re
    out = {}
    for f in sorted(glob.glob(os.path.join(path, "params.*double.h"))):
        txt = open(f).read()
        vals = re.search(r"RNS_MODULI_VALUES\s*\{([^}]*)\}", txt).group(1)
        mods = [int(v) for v in vals.replace("\\", " ").split(",") if v.strip()]
        out[len(mods)] = mods
    return out
