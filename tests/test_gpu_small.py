"""GPU parity of the small-modulus stage 2 of mp_gemm (kernels_small.cuh) -- every stage against exact Python
integer arithmetic on the same inputs, and the whole call against the limb-plane path and the reference order."""
import numpy as np
import pytest

import oracle
from util import diff_fields, get_oracle, random_records
from test_gpu_blas import _gemm, _special_case_inputs, _transpose_recs

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _round_up(v, a):
    return (v + a - 1) // a * a


def _signed_ints(orc, recs):
    """(signed significand, exponent, non-zero) per record"""
    out = []
    for r in recs:
        nz = r["eval"][1]["frac"] != 0
        x = orc.to_int(r) if nz else 0
        out.append((-x if int(r["sign"]) else x, int(r["exp"]), bool(nz)))
    return out


@pytest.mark.parametrize("N,shape,div,tiled", [(8, (40, 24, 70), 4, False), (32, (33, 20, 50), 4, False), (32, (20, 12, 40), 3, True), (16, (17, 9, 33), 5, False),
                                               (8, (300, 260, 200), 4, False)])
def test_small_path_stages_exact(pkg, N, shape, div, tiled):
    """one-byte planes of A' and B' (stage 1), per-modulus sums (stage 2) and the extended residue planes (stage 3a)
    against exact integers"""
    ctx = pkg.Context(N, 0)
    orc = get_oracle(N, oracle.DEVICE)
    bits = orc.precision // div
    m, n, k = shape
    A, B, C = _special_case_inputs(N, m, n, k, bits, 301)
    alpha = random_records(N, 1, bits, 304)
    beta = random_records(N, 1, bits, 305)
    ctx.set_stage2_kernel(pkg.STAGE2_SMALL_TILED if tiled else pkg.STAGE2_SMALL)
    ctx.set_stage3_kernel(0)      # base extension as a separate kernel (the default): the residue planes are written in full
    _gemm(pkg, ctx, m, n, k, alpha, A, B, beta, C, pkg.MODE_AUTO)
    P, nin = ctx.last_small_base()
    assert P > 0 and nin > 0, "the small base must be selected for p/%d-bit inputs (got %d, %d)" % (div, P, nin)
    assert ctx.last_base_size() == 0
    ps = ctx.small_moduli(P)
    assert ps[0] == 256 and all(p > 1 for p in ps)
    m_p, m_ps, k_p, n_ps = _round_up(m, 128), _round_up(m, 256), _round_up(k, 128), _round_up(n, 256)
    n_p = n_ps      # one column panel: every per-panel plane is padded to the 256-column tiles of stage 2
    a = _signed_ints(orc, A)     # column-major m x k: entry (i, l) at i + l m
    b = _signed_ints(orc, B)     # column-major k x n: entry (l, j) at l + j k

    def aligned(vals, outer, inner, at):
        """exact pre-shifted integers per line, and the line's exponent base"""
        out, base = [], []
        for o in range(outer):
            es = [vals[at(o, l)][1] for l in range(inner) if vals[at(o, l)][2]]
            emin = min(es) if es else 0
            base.append(emin)
            out.append([(vals[at(o, l)][0] << (vals[at(o, l)][1] - emin)) if vals[at(o, l)][2] else 0 for l in range(inner)])
        return out, base
    Ai, ra = aligned(a, m, k, lambda i, l: i + l * m)
    Bi, cb = aligned(b, n, k, lambda j, l: l + j * k)
    # ---- stage 1 ----
    QA = ctx.debug_read_workspace(8, 0, P * m_ps * k_p).reshape(P, m_ps, k_p)
    QB = ctx.debug_read_workspace(9, 0, P * n_ps * k_p).reshape(P, n_ps, k_p)
    for j, p in enumerate(ps):
        wantA = np.array([[v % p for v in row] for row in Ai], dtype=np.uint8)
        wantB = np.array([[v % p for v in row] for row in Bi], dtype=np.uint8)
        assert np.array_equal(QA[j, :m, :k], wantA), "A' plane of modulus %d" % p
        assert np.array_equal(QB[j, :n, :k], wantB), "B' plane of modulus %d" % p
        assert not QA[j, m:, :].any() and not QA[j, :, k:].any() and not QB[j, n:, :].any() and not QB[j, :, k:].any()
    # ---- stage 2 ----
    S = [[sum(Ai[i][l] * Bi[j][l] for l in range(k)) for i in range(m)] for j in range(n)]    # S[j][i]
    S8 = ctx.debug_read_workspace(10, 0, P * n_ps * m_ps).reshape(P, n_ps, m_ps)
    for j, p in enumerate(ps):
        want = np.array([[v % p for v in row] for row in S], dtype=np.uint8)
        assert np.array_equal(S8[j, :n, :m], want), "sums modulo %d" % p
    # ---- stage 3a ----
    SP = ctx.debug_read_workspace(5, 0, N * n_p * m_p * 4).view(np.int32).reshape(N, n_p, m_p)
    mods = orc.c["moduli"]
    for q, mq in enumerate(mods):
        want = np.array([[v % mq for v in row] for row in S], dtype=np.int64)
        assert np.array_equal(SP[q, :n, :m].astype(np.int64), want), "extended residues modulo %d" % mq
    ctx.close()


@pytest.mark.parametrize("N,bits_div,shape,ta,tb", [(8, 4, (140, 20, 60), 111, 111), (32, 4, (130, 70, 129), 111, 111), (32, 8, (64, 64, 300), 112, 111),
                                                     (16, 4, (40, 300, 50), 111, 112), (24, 4, (33, 5, 20), 112, 112), (64, 4, (20, 6, 24), 111, 111),
                                                     (64, 16, (130, 20, 64), 111, 111), (16, 2, (40, 12, 50), 111, 111), (32, 2, (40, 12, 50), 111, 111)])
def test_gemm_small_base_identical(pkg, N, bits_div, shape, ta, tb):
    """small-modulus stage 2 == limb planes on the format's moduli, record for record (digits, sign, exp, eval), including zero
    lines, far exponents, cancellation and transposed operands; sums beyond the small base (~362 bits) must fall through to the limb planes"""
    ctx = pkg.Context(N, 0)
    orc = get_oracle(N, oracle.DEVICE)
    bits = orc.precision // bits_div - (8 if bits_div == 2 else 0)
    m, n, k = shape
    A, B, C = _special_case_inputs(N, m, n, k, bits, 311)
    if ta != 111:
        A = _transpose_recs(A, m, k)
    if tb != 111:
        B = _transpose_recs(B, k, n)
    alpha = random_records(N, 1, bits, 314)
    beta = random_records(N, 1, bits, 315)
    out, sel = [], []
    for kind in (pkg.STAGE2_UMMA, pkg.STAGE2_SMALL, pkg.STAGE2_SMALL_T128, pkg.STAGE2_SMALL_K64, pkg.STAGE2_SMALL_TILED):
        ctx.set_stage2_kernel(kind)
        out.append(_gemm(pkg, ctx, m, n, k, alpha, A, B, beta, C, pkg.MODE_AUTO, ta, tb))
        sel.append(ctx.last_small_base())
    for other in out[1:]:
        bad = diff_fields(out[0], other)
        assert bad.size == 0, "%d/%d entries differ, first %d\n%s\n%s" % (bad.size, m * n, bad[0], out[0][bad[0]], other[bad[0]])
    assert sel[0] == (0, 0)
    if 2 * bits + 100 < 362:      # the sums certainly fit the product of the one-byte moduli (~2^362)
        assert sel[1][0] > 0, "%d-bit inputs must select the small base" % bits
    if 2 * bits > 362:
        assert sel[1] == (0, 0), "%d-bit inputs do not fit the small base" % bits
    ctx.close()


def test_gemm_small_base_long_k(pkg):
    """k beyond one accumulation chunk of the small-modulus kernel (32768) and a reference-order cross-check on a sample"""
    N = 8
    ctx = pkg.Context(N, 0)
    orc = get_oracle(N, oracle.DEVICE)
    bits = orc.precision // 4
    m, n, k = 6, 5, 33000
    A = random_records(N, m * k, bits, 321)
    B = random_records(N, k * n, bits, 322)
    C = random_records(N, m * n, bits, 323)
    alpha = random_records(N, 1, bits, 324)
    beta = random_records(N, 1, bits, 325)
    ctx.set_stage2_kernel(pkg.STAGE2_SMALL)
    got = _gemm(pkg, ctx, m, n, k, alpha, A, B, beta, C, pkg.MODE_AUTO)
    assert ctx.last_small_base()[0] > 0
    want = _gemm(pkg, ctx, m, n, k, alpha, A, B, beta, C, pkg.MODE_REFERENCE_ORDER)
    bad = diff_fields(got, want, ("digits", "sign", "exp"))
    assert bad.size == 0, "%d/%d differ" % (bad.size, m * n)
    ctx.close()


@pytest.mark.parametrize("N,shape,spread", [(8, (70, 50, 600), 0), (16, (130, 64, 1100), 40), (32, (33, 20, 520), 3)])
def test_minplus_candidate_lists(pkg, N, shape, spread):
    """(min,+) exponent product from candidate lists == dense kernel (records identical), on inputs whose exponents are spread so
    that few / many pairs fall outside the candidates (the dense recomputation list is exercised when spread is large)"""
    ctx = pkg.Context(N, 0)
    orc = get_oracle(N, oracle.DEVICE)
    bits = orc.precision // 4
    m, n, k = shape
    A = random_records(N, m * k, bits, 331)
    B = random_records(N, k * n, bits, 332)
    C = random_records(N, m * n, bits, 333)
    if spread:
        rng = np.random.RandomState(7)
        # every entry gets a large random exponent offset except a few per line: the small shifts (candidates) are rare and scattered
        A["exp"] += rng.randint(0, spread, size=A.shape).astype(np.int32)
        B["exp"] += rng.randint(0, spread, size=B.shape).astype(np.int32)
    alpha = random_records(N, 1, bits, 334)
    beta = random_records(N, 1, bits, 335)
    out, dense = [], []
    for kind in (2, 0):
        ctx.set_stage1_kernel(kind)
        out.append(_gemm(pkg, ctx, m, n, k, alpha, A, B, beta, C, pkg.MODE_AUTO))
        dense.append(ctx.last_minplus_dense_count())
    bad = diff_fields(out[0], out[1])
    assert bad.size == 0, "%d/%d entries differ, first %d" % (bad.size, m * n, bad[0])
    assert dense[0] == 0
    if spread == 0:
        assert dense[1] < m * n // 10
    want = _gemm(pkg, ctx, m, n, k, alpha, A, B, beta, C, pkg.MODE_REFERENCE_ORDER)
    assert ctx.last_fallback_count() == 0
    bad = diff_fields(out[1], want, ("digits", "sign", "exp"))
    assert bad.size == 0
    ctx.close()
