"""Full-precision (p-bit) inputs: the exact sums of the fast path exceed the number format, stage 3 rebuilds them in binary from the
one-byte base and rounds ONCE (csrc/kernels_bin.cuh) where the reference rounds every product and partial sum
(src/arith/mul.cuh:108-110, src/arith/add.cuh:197-199).  Checked here:
  * alpha = 1, beta = 0: digits, sign, exponent == the exact integer sum rounded to nearest at MP_PRECISION bits (Python integers), the
    interval evaluation encloses T / M tightly;
  * general alpha, beta: C = rn(rn(alpha rn(S)) + rn(beta C)), every rn the rounding to nearest of the exact value at MP_PRECISION bits,
    against a Python-integer model bit for bit (the binary epilogue, k_bin_norm2); with MPRES_BIN_EPILOGUE=0 the residue-parallel
    epilogue: the reference's mp_mul, mp_mul, mp_add (DEVICE oracle) applied to the rounded sums, bit for bit;
  * against exact rationals within the reference's own error model |err| <= gamma_k sum |a||b|, u = 4 / sqrt(M)
    (tests/blas/accuracy/test_dot_accuracy.cu:41-72), and at least as close as the reference-order k-loop on average."""
from fractions import Fraction

import numpy as np
import pytest

import oracle
from util import diff_fields, get_oracle, random_records, unit_roundoff

pytestmark = pytest.mark.gpu


def _gemm(pkg, ctx, m, n, k, alpha, A, B, beta, C, mode, ta=111, tb=111, lda=None, ldb=None):
    ctx.set_mode(mode)
    dA, dB, dC = ctx.mp_array_from_host(A), ctx.mp_array_from_host(B), ctx.mp_array_from_host(C)
    dal, dbe = ctx.mp_array_from_host(alpha), ctx.mp_array_from_host(beta)
    pkg.mp_gemm(ctx, ta, tb, m, n, k, dal, dA, lda or (m if ta == 111 else k), dB, ldb or (k if tb == 111 else n), dbe, dC, m)
    return dC.device2host()


def _signed(orc, recs):
    return [(-1 if int(r["sign"]) else 1) * orc.to_int(r) for r in recs], [int(r["exp"]) for r in recs]


def _one_zero(orc):
    one = orc.set_ints([0], [1], [0])
    zero = orc.set_ints([0], [0], [0])
    return one, zero


def _zeros(orc, count):
    """exact zeros with exponent 0: with alpha = 1, beta = 0 the epilogue's mp_add then returns the rounded sum unchanged (a non-zero
    exponent of C would make it shift the sum, src/arith/add.cuh:139-150)"""
    z = orc.set_ints([0], [0], [0])
    assert int(z[0]["exp"]) == 0
    return np.repeat(z, count)


def _exact_sums(orc, A, B, m, n, k):
    """S(i, j) as (integer, exponent): sum_l a b 2^(ea + eb - base) with base = min over the non-zero entries of the row + of the column"""
    xa, ea = _signed(orc, A)
    xb, eb = _signed(orc, B)
    out = {}
    for j in range(n):
        cb = [eb[l + j * k] for l in range(k) if xb[l + j * k]]
        for i in range(m):
            ra = [ea[i + l * m] for l in range(k) if xa[i + l * m]]
            if not ra or not cb:
                out[i, j] = (0, 0)
                continue
            base = min(ra) + min(cb)
            s = 0
            for l in range(k):
                a, b = xa[i + l * m], xb[l + j * k]
                if a and b:
                    s += a * b << (ea[i + l * m] + eb[l + j * k] - base)
            out[i, j] = (s, base)
    return out


def _round_nearest(s, base, prec):
    """(sign, T, exp): |s| rounded to nearest at prec bits, ties away from zero"""
    sign, mag = (1, -s) if s < 0 else (0, s)
    L = mag.bit_length()
    drop = max(0, L - prec)
    if drop:
        mag = (mag + (1 << (drop - 1))) >> drop
    return sign, mag, base + drop


@pytest.mark.parametrize("N,shape,spread", [(8, (40, 24, 70), 0), (8, (130, 70, 300), 6), (8, (24, 20, 1100), 3),
                                            (16, (40, 24, 70), 0), (24, (20, 30, 130), 4), (32, (33, 20, 50), 0), (32, (130, 40, 300), 9)])
def test_rounded_sums_match_exact_integers(pkg, N, shape, spread):
    """N >= 16: the sums exceed the one-byte base too -- the significands are cut into slices (k_choose_base), stage 2 forms the slice sums
    and stage 3 puts them together in binary"""
    ctx = pkg.Context(N, 0)
    orc = get_oracle(N, oracle.DEVICE)
    prec = orc.precision
    m, n, k = shape
    A, B = random_records(N, m * k, prec, 901), random_records(N, k * n, prec, 902)
    C = _zeros(orc, m * n)
    rng = np.random.RandomState(904)
    if spread:
        A["exp"] += rng.randint(0, spread, size=A.shape).astype(np.int32)
        B["exp"] -= rng.randint(0, spread, size=B.shape).astype(np.int32)
    A[3::17] = orc.set_ints([0], [0], [0])[0]                       # some exact zeros
    A[5 + np.arange(k) * m] = orc.set_ints([0], [0], [0])[0]        # a zero row of A
    one, zero = _one_zero(orc)
    got = _gemm(pkg, ctx, m, n, k, one, A, B, zero, C, pkg.MODE_AUTO)
    assert ctx.last_binary_rounding(), "the call did not take the binary-rounding stage 3"
    assert ctx.last_fallback_count() == 0
    P, nin = ctx.last_small_base()
    assert P > 0
    S = _exact_sums(orc, A, B, m, n, k)
    M = orc.c["M"]
    mods = orc.c["moduli"]
    for j in range(n):
        for i in range(m):
            s, base = S[i, j]
            g = got[i + j * m]
            if s == 0:
                assert not g["digits"].any() and g["eval"]["frac"][1] == 0, (i, j)
                continue
            sign, T, ex = _round_nearest(s, base, prec)
            assert int(g["sign"]) == sign and int(g["exp"]) == ex, (i, j, g, sign, ex)
            assert [int(d) for d in g["digits"]] == [T % q for q in mods], (i, j)
            lo = Fraction(float(g["eval"]["frac"][0])) * Fraction(2) ** int(g["eval"]["exp"][0])
            up = Fraction(float(g["eval"]["frac"][1])) * Fraction(2) ** int(g["eval"]["exp"][1])
            x = Fraction(T, M)
            assert lo <= x <= up and (up - lo) <= x / 2 ** 20, (i, j, float(lo), float(x), float(up))   # (the epilogue multiplies by the interval of alpha = 1: 1e-7 wide)
    ctx.close()


def _rn(mag, e, prec):
    L = mag.bit_length()
    drop = max(0, L - prec)
    if drop:
        mag = (mag + (1 << (drop - 1))) >> drop
    return mag, e + drop


def _model_entry(s, base, al, be, c, prec):
    """(sign, significand, exponent) of rn(rn(alpha rn(S)) + rn(beta C)); al, be, c = (sign, integer, exponent)"""
    m1 = e1 = s1 = 0
    if s and al[1]:
        t, et = _rn(abs(s), base, prec)
        m1, e1 = _rn(t * al[1], et + al[2], prec)
        s1 = (1 if s < 0 else 0) ^ al[0]
    m2 = e2 = s2 = 0
    if be[1] and c[1]:
        m2, e2 = _rn(c[1] * be[1], c[2] + be[2], prec)
        s2 = c[0] ^ be[0]
    if m1 == 0 and m2 == 0:
        return 0, 0, 0
    top1, top2 = e1 + m1.bit_length(), e2 + m2.bit_length()
    if m1 == 0 or (m2 and top2 - top1 > prec + 2):
        return s2, m2, e2
    if m2 == 0 or top1 - top2 > prec + 2:
        return s1, m1, e1
    emin = min(e1, e2)
    v = (-1 if s1 else 1) * (m1 << (e1 - emin)) + (-1 if s2 else 1) * (m2 << (e2 - emin))
    if v == 0:
        return 0, 0, 0
    mr, er = _rn(abs(v), emin, prec)
    return (1 if v < 0 else 0), mr, er


@pytest.mark.parametrize("N,shape,bits_c,spread", [(8, (33, 21, 64), None, 0), (8, (64, 40, 200), 26, 5), (8, (130, 24, 90), None, 40),
                                                   (16, (33, 21, 64), None, 60), (32, (30, 18, 40), None, 0), (32, (20, 12, 600), 106, 200)])
def test_binary_epilogue_matches_the_integer_model(pkg, N, shape, bits_c, spread):
    ctx = pkg.Context(N, 0)
    orc = get_oracle(N, oracle.DEVICE)
    prec = orc.precision
    m, n, k = shape
    A, B = random_records(N, m * k, prec, 931), random_records(N, k * n, prec, 932)
    C = random_records(N, m * n, bits_c or prec, 933)
    rng = np.random.RandomState(934)
    if spread:
        C["exp"] += rng.randint(-spread, spread, size=C.shape).astype(np.int32)     # t2 dominant / negligible / cancelling against t1
    C[2::11] = orc.set_ints([0], [0], [0])[0]
    A[7 + np.arange(k) * m] = orc.set_ints([0], [0], [0])[0]                           # a zero row of A: C = rn(beta C)
    alpha, beta = random_records(N, 1, prec, 935), random_records(N, 1, prec, 936)
    S = _exact_sums(orc, A, B, m, n, k)
    mods = orc.c["moduli"]
    triple = lambda r: (int(r["sign"]), orc.to_int(r), int(r["exp"]))
    for al_, be_ in ((alpha, beta), (alpha, orc.set_ints([0], [0], [0])), (orc.set_ints([1], [3], [-1]), beta)):
        got = _gemm(pkg, ctx, m, n, k, al_, A, B, be_, C, pkg.MODE_AUTO)
        assert ctx.last_binary_rounding() and ctx.last_fallback_count() == 0
        al, be = triple(al_[0]), triple(be_[0])
        for j in range(n):
            for i in range(m):
                s, base = S[i, j]
                sg, mant, ex = _model_entry(s, base, al, be, triple(C[i + j * m]), prec)
                g = got[i + j * m]
                assert (int(g["sign"]), int(g["exp"])) == (sg, ex), (i, j, g, sg, mant, ex)
                assert [int(d) for d in g["digits"]] == [mant % q for q in mods], (i, j)
                if mant:
                    lo = Fraction(float(g["eval"]["frac"][0])) * Fraction(2) ** int(g["eval"]["exp"][0])
                    up = Fraction(float(g["eval"]["frac"][1])) * Fraction(2) ** int(g["eval"]["exp"][1])
                    x = Fraction(mant, orc.c["M"])
                    assert lo <= x <= up and (up - lo) <= x / 2 ** 45, (i, j)
                else:
                    assert g["eval"]["frac"][1] == 0
    ctx.close()


@pytest.mark.parametrize("N,shape,bits_c", [(8, (33, 21, 64), None), (8, (64, 40, 200), 26)])
def test_epilogue_is_the_reference_sequence_on_the_rounded_sums(pkg, N, shape, bits_c, monkeypatch):
    monkeypatch.setenv("MPRES_BIN_EPILOGUE", "0")          # the residue-parallel epilogue (also the fallback for scalars wider than the precision)
    ctx = pkg.Context(N, 0)
    orc = get_oracle(N, oracle.DEVICE)
    prec = orc.precision
    m, n, k = shape
    A, B = random_records(N, m * k, prec, 911), random_records(N, k * n, prec, 912)
    C = random_records(N, m * n, bits_c or prec, 913)
    alpha, beta = random_records(N, 1, prec, 914), random_records(N, 1, prec, 915)
    one, zero = _one_zero(orc)
    T = _gemm(pkg, ctx, m, n, k, one, A, B, zero, _zeros(orc, m * n), pkg.MODE_AUTO)   # the rounded sums, interval evaluations included
    got = _gemm(pkg, ctx, m, n, k, alpha, A, B, beta, C, pkg.MODE_AUTO)
    assert ctx.last_binary_rounding()
    al, be = np.repeat(alpha, m * n), np.repeat(beta, m * n)
    want = orc.add(orc.mul(C, be), orc.mul(T, al))                             # src/blas/gemm.cuh:142-166
    bad = diff_fields(got, want, ("digits", "sign", "exp"))
    assert bad.size == 0, "%d/%d differ, first %d\n%s\n%s" % (bad.size, m * n, bad[0], got[bad[0]], want[bad[0]])
    ctx.close()


@pytest.mark.parametrize("N,shape,trans", [(8, (20, 16, 128), (111, 111)), (8, (17, 13, 90), (112, 112)), (32, (12, 10, 64), (112, 111)), (24, (9, 8, 40), (111, 112))])
def test_accuracy_against_exact_rationals(pkg, N, shape, trans):
    ctx = pkg.Context(N, 0)
    orc = get_oracle(N, oracle.DEVICE)
    prec = orc.precision
    m, n, k = shape
    A, B, C = random_records(N, m * k, prec, 921), random_records(N, k * n, prec, 922), random_records(N, m * n, prec, 923)
    alpha, beta = random_records(N, 1, prec, 924), random_records(N, 1, prec, 925)
    ta, tb = trans
    A_in = A if ta == 111 else np.ascontiguousarray(A.reshape(k, m).T).reshape(-1)
    B_in = B if tb == 111 else np.ascontiguousarray(B.reshape(n, k).T).reshape(-1)
    got = _gemm(pkg, ctx, m, n, k, alpha, A_in, B_in, beta, C, pkg.MODE_AUTO, ta, tb)
    assert ctx.last_binary_rounding() and ctx.last_fallback_count() == 0
    ref = _gemm(pkg, ctx, m, n, k, alpha, A_in, B_in, beta, C, pkg.MODE_REFERENCE_ORDER, ta, tb)
    u = unit_roundoff(orc)
    gam = (k + 3) * u / (1 - (k + 3) * u)
    fa, fb, fc = [orc.to_fraction(x) for x in A], [orc.to_fraction(x) for x in B], [orc.to_fraction(x) for x in C]
    al, be = orc.to_fraction(alpha[0]), orc.to_fraction(beta[0])
    ours, theirs = Fraction(0), Fraction(0)
    for j in range(n):
        for i in range(m):
            exact = al * sum(fa[i + l * m] * fb[l + j * k] for l in range(k)) + be * fc[i + j * m]
            bound = gam * (abs(al) * sum(abs(fa[i + l * m] * fb[l + j * k]) for l in range(k)) + abs(be * fc[i + j * m]))
            err = abs(orc.to_fraction(got[i + j * m]) - exact)
            assert err <= bound, (i, j, float(err), float(bound))
            ours += err / bound
            theirs += abs(orc.to_fraction(ref[i + j * m]) - exact) / bound
    assert ours <= theirs, "one rounding of the exact sum must not be less accurate than the reference-order loop (%g vs %g)" % (float(ours), float(theirs))
    ctx.close()
