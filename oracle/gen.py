"""oracle/gen.py -- TEST INFRASTRUCTURE ONLY.  Deterministic synthetic inputs.

Follows create_random_array (/root/reference/tests/tsthelper.cuh:37-70): a uniform `bits`-bit integer
times a uniform double in (-1, 1), rounded to `bits` bits, scaled by 2^-bits -- but with a fixed seed
instead of time(NULL).  Returns (sign, integer significand, exponent) triples, value =
(-1)^sign * mant * 2^exp, which Oracle.set_ints / RefLib.set_ints turn into mp_float_t records.
"""
import random


def _round_to_bits(v, bits):
    """round-to-nearest-even of a positive integer to `bits` significant bits -> (mant, shift)"""
    n = v.bit_length()
    if n <= bits:
        return v, 0
    sh = n - bits
    q, r = v >> sh, v & ((1 << sh) - 1)
    half = 1 << (sh - 1)
    if r > half or (r == half and (q & 1)):
        q += 1
        if q.bit_length() > bits:
            q >>= 1
            sh += 1
    return q, sh


def random_values(count, bits, seed):
    rng = random.Random(seed)
    signs, mants, exps = [], [], []
    for _ in range(count):
        z = rng.getrandbits(bits)
        u = rng.uniform(-1.0, 1.0)
        num, den_exp = abs(u).hex(), 0
        # exact integer form of |u| = f * 2^e with 53-bit f
        import math
        f, e = math.frexp(abs(u))
        fi = int(f * (1 << 53))
        prod = z * fi                      # exact, value = prod * 2^(e - 53)
        mant, sh = _round_to_bits(prod, bits)
        exp = e - 53 + sh - bits           # the final * 2^-bits is exact
        if mant == 0:
            signs.append(0); mants.append(0); exps.append(0)
            continue
        signs.append(1 if u < 0 else 0)
        mants.append(mant)
        exps.append(exp)
    return signs, mants, exps
