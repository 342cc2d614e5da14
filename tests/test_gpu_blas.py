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


def _transpose_recs(X, rows, cols):
    """column-major rows x cols -> column-major cols x rows"""
    return np.ascontiguousarray(X.reshape(cols, rows).T).reshape(-1)


@pytest.mark.parametrize("N,shape", [(8, (33, 17, 41)), (8, (200, 150, 300)), (16, (130, 70, 129)), (32, (64, 64, 256)), (24, (40, 40, 40))])
def test_gemm_fast_path_bit_exact_quarter_precision(pkg, N, shape):
    """p/4-bit inputs (the reference's benchmark convention): nothing rounds inside the k-loop, so the
    exact-window fast path must give the reference's digits, sign and exponent bit for bit, with no
    element routed to the fallback."""
    ctx = pkg.Context(N, 0)
    orc = get_oracle(N, oracle.DEVICE)
    bits = orc.precision // 4
    m, n, k = shape
    A = random_records(N, m * k, bits, 21)
    B = random_records(N, k * n, bits, 22)
    C = random_records(N, m * n, bits, 23)
    alpha = random_records(N, 1, bits, 24)
    beta = random_records(N, 1, bits, 25)
    got = _gemm(pkg, ctx, m, n, k, alpha, A, B, beta, C, pkg.MODE_AUTO)
    assert ctx.last_fallback_count() == 0
    ref_order = _gemm(pkg, ctx, m, n, k, alpha, A, B, beta, C, pkg.MODE_REFERENCE_ORDER)
    bad = diff_fields(got, ref_order, ("digits", "sign", "exp"))
    assert bad.size == 0, "%d/%d differ from reference order, first %d\n%s\n%s" % (bad.size, m * n, bad[0], got[bad[0]], ref_order[bad[0]])
    if oracle.have_ref(N) and m * n * k < 3000000:
        r, _, _ = oracle.RefLib(N, gpu=True).gpu_gemm(m, n, k, alpha, A, B, beta, C)
        assert diff_fields(got, r, ("digits", "sign", "exp")).size == 0
    ctx.close()


def test_gemm_fast_path_transposes(pkg):
    N = 8
    ctx = pkg.Context(N, 0)
    orc = get_oracle(N, oracle.DEVICE)
    bits = orc.precision // 4
    m, n, k = 37, 29, 53
    A = random_records(N, m * k, bits, 31)
    B = random_records(N, k * n, bits, 32)
    C = random_records(N, m * n, bits, 33)
    alpha = random_records(N, 1, bits, 34)
    beta = random_records(N, 1, bits, 35)
    want = _gemm(pkg, ctx, m, n, k, alpha, A, B, beta, C, pkg.MODE_REFERENCE_ORDER)
    At, Bt = _transpose_recs(A, m, k), _transpose_recs(B, k, n)
    for mode in (pkg.MODE_AUTO, pkg.MODE_REFERENCE_ORDER):
        for ta, tb, a_, b_ in ((112, 111, At, B), (111, 112, A, Bt), (112, 112, At, Bt)):
            got = _gemm(pkg, ctx, m, n, k, alpha, a_, b_, beta, C, mode, ta, tb)
            assert diff_fields(got, want, ("digits", "sign", "exp")).size == 0, (mode, ta, tb)
    ctx.close()


@pytest.mark.parametrize("N", [40])
def test_gemm_full_precision_inputs(pkg, N):
    """p-bit inputs: every product needs a rounding in the reference, the exact window does not fit in
    M.  Formats without the single-rounding stage 3 (more than 32 moduli) must route those elements to the
    reference-order fallback (bit-exact again); N <= 32 is tests/test_gpu_fullprec.py."""
    ctx = pkg.Context(N, 0)
    orc = get_oracle(N, oracle.DEVICE)
    bits = orc.precision
    m, n, k = 24, 20, 50
    A = random_records(N, m * k, bits, 41)
    B = random_records(N, k * n, bits, 42)
    C = random_records(N, m * n, bits, 43)
    alpha = random_records(N, 1, bits, 44)
    beta = random_records(N, 1, bits, 45)
    got = _gemm(pkg, ctx, m, n, k, alpha, A, B, beta, C, pkg.MODE_AUTO)
    assert ctx.last_fallback_count() == m * n
    want = _gemm(pkg, ctx, m, n, k, alpha, A, B, beta, C, pkg.MODE_REFERENCE_ORDER)
    assert diff_fields(got, want).size == 0
    ctx.close()


def test_gemm_mixed_width_accuracy(pkg):
    """half-precision-width inputs: the reference rounds some partial sums, the fast path rounds once.
    Results must agree with the exact rational result within the reference's error model
    (test_dot_accuracy.cu:59-72): |err| <= gamma_k * sum |a||b| with u = 4/sqrt(M)."""
    from fractions import Fraction
    N = 8
    ctx = pkg.Context(N, 0)
    orc = get_oracle(N, oracle.DEVICE)
    bits = orc.precision // 2 - 8
    m, n, k = 12, 10, 64
    A = random_records(N, m * k, bits, 51)
    B = random_records(N, k * n, bits, 52)
    C = random_records(N, m * n, bits, 53)
    alpha = random_records(N, 1, bits, 54)
    beta = random_records(N, 1, bits, 55)
    got = _gemm(pkg, ctx, m, n, k, alpha, A, B, beta, C, pkg.MODE_AUTO)
    u = unit_roundoff(orc)
    gam = (k + 3) * u / (1 - (k + 3) * u)
    fa = [orc.to_fraction(x) for x in A]
    fb = [orc.to_fraction(x) for x in B]
    fc = [orc.to_fraction(x) for x in C]
    al, be = orc.to_fraction(alpha[0]), orc.to_fraction(beta[0])
    for j in range(n):
        for i in range(m):
            exact = al * sum(fa[i + l * m] * fb[l + j * k] for l in range(k)) + be * fc[i + j * m]
            bound = gam * (abs(al) * sum(abs(fa[i + l * m] * fb[l + j * k]) for l in range(k)) + abs(be * fc[i + j * m]))
            err = abs(orc.to_fraction(got[i + j * m]) - exact)
            assert err <= bound, (i, j, float(err), float(bound))
    ctx.close()


@pytest.mark.parametrize("N,shape", [(8, (200, 150, 300)), (32, (130, 70, 129)), (8, (20, 10, 8300)), (16, (256, 128, 1024))])
def test_gemm_stage2_kernels_agree(pkg, N, shape):
    """The three stage-2 kernels (tcgen05 stacked / unstacked, legacy mma.sync) give the same result as
    the reference-order k-loop, bit for bit; k = 8300 crosses the 8064-byte accumulation chunk."""
    ctx = pkg.Context(N, 0)
    orc = get_oracle(N, oracle.DEVICE)
    bits = orc.precision // 4
    m, n, k = shape
    A = random_records(N, m * k, bits, 61)
    B = random_records(N, k * n, bits, 62)
    C = random_records(N, m * n, bits, 63)
    alpha = random_records(N, 1, bits, 64)
    beta = random_records(N, 1, bits, 65)
    want = _gemm(pkg, ctx, m, n, k, alpha, A, B, beta, C, pkg.MODE_REFERENCE_ORDER)
    for kind in (pkg.STAGE2_SMALL, pkg.STAGE2_SMALL_T128, pkg.STAGE2_SMALL_K64, pkg.STAGE2_SMALL_TILED, pkg.STAGE2_UMMA, pkg.STAGE2_UMMA_UNSTACKED, pkg.STAGE2_MMA_SYNC):
        ctx.set_stage2_kernel(kind)
        got = _gemm(pkg, ctx, m, n, k, alpha, A, B, beta, C, pkg.MODE_AUTO)
        assert ctx.last_fallback_count() == 0
        bad = diff_fields(got, want, ("digits", "sign", "exp"))
        assert bad.size == 0, "stage2 kernel %d: %d/%d differ, first %d" % (kind, bad.size, m * n, bad[0])
    ctx.close()


def _special_case_inputs(N, m, n, k, bits, seed):
    """random matrices with the special cases of the normalisation stage mixed in: a zero row of A, a
    zero column of B, zero entries of C, rows scaled by large powers of two (alignment shifts in the
    alpha*S + beta*C addition), an exactly cancelling row."""
    orc = get_oracle(N, oracle.DEVICE)
    A = random_records(N, m * k, bits, seed)
    B = random_records(N, k * n, bits, seed + 1)
    C = random_records(N, m * n, bits, seed + 2)
    zero = orc.set_ints([0], [0], [0])[0]
    A2 = A.reshape(k, m).copy()      # column-major m x k: A2[l, i]
    B2 = B.reshape(n, k).copy()      # B2[j, l]
    C2 = C.reshape(n, m).copy()
    A2[:, 1] = zero                  # zero row of A
    B2[2, :] = zero                  # zero column of B
    C2[0, :] = zero                  # zero column of C
    C2[3, 5] = zero
    A2[:, 4]["exp"] += 37            # row scaled up: alpha*S dominates beta*C
    A2[:, 6]["exp"] -= 45            # row scaled down
    C2[:, 7]["exp"] -= 300           # beta*C negligible: far alignment shift / zeroing
    if k >= 2:                       # S(8, :) == 0 with non-zero terms
        A2[1, 8] = A2[0, 8]
        A2[1, 8]["sign"] ^= 1
        A2[2:, 8] = zero
        B2[:, 1] = B2[:, 0]
    return A2.reshape(-1), B2.reshape(-1), C2.reshape(-1)


@pytest.mark.parametrize("N,bits_div,shape", [(8, 4, (140, 20, 60)), (32, 4, (130, 9, 40)), (8, 2, (40, 12, 50)), (16, 3, (70, 10, 33)),
                                              (24, 4, (33, 5, 20)), (64, 4, (20, 6, 24))])
def test_gemm_stage3_kernels_identical(pkg, N, bits_div, shape):
    """The entry-per-thread normalisation kernel (+ list kernel) and the residue-parallel tile kernel give
    identical records -- digits, sign, exponent AND interval evaluations -- on inputs that exercise the
    special cases (zeros, far exponents, exact cancellation, roundings)."""
    ctx = pkg.Context(N, 0)
    orc = get_oracle(N, oracle.DEVICE)
    bits = orc.precision // bits_div - (8 if bits_div == 2 else 0)
    m, n, k = shape
    A, B, C = _special_case_inputs(N, m, n, k, bits, 71)
    alpha = random_records(N, 1, bits, 74)
    beta = random_records(N, 1, bits, 75)
    out = []
    for kind in (1, 0, 2, 3, 4):  # tile kernel, entry-per-thread (32-bit Barrett products, staged sums), generic products, fused base extension, unstaged sums
        ctx.set_stage3_kernel(kind)
        out.append(_gemm(pkg, ctx, m, n, k, alpha, A, B, beta, C, pkg.MODE_FAST))
    for other in out[1:]:
        bad = diff_fields(out[0], other)
        assert bad.size == 0, "%d/%d entries differ, first %d\n%s\n%s" % (bad.size, m * n, bad[0], out[0][bad[0]], other[bad[0]])
    if bits_div == 4:
        want = _gemm(pkg, ctx, m, n, k, alpha, A, B, beta, C, pkg.MODE_REFERENCE_ORDER)
        # Bit-identical except where an exact zero takes part (zero row / column / cancelling row): there the
        # reference's exponent depends on the zero's own exponent (DESIGN.md section 7 "zeros") -- value-equal.
        bad = diff_fields(out[1], want, ("digits", "sign", "exp"))
        for e in bad:
            row, col = e % m, e // m
            assert row in (1, 8) or col == 2, "entry (%d, %d) differs from the reference order\n%s\n%s" % (row, col, out[1][e], want[e])
            assert orc.to_fraction(out[1][e]) == orc.to_fraction(want[e]), (row, col)
    # beta == 0 and alpha == 0
    zero = orc.set_ints([0], [0], [0])
    for al, be in ((alpha, zero), (zero, beta)):
        res = []
        for kind in (1, 0, 2, 3, 4):
            ctx.set_stage3_kernel(kind)
            res.append(_gemm(pkg, ctx, m, n, k, al, A, B, be, C, pkg.MODE_FAST))
        assert all(diff_fields(res[0], r).size == 0 for r in res[1:])
    ctx.close()


@pytest.mark.parametrize("N,shape,ta,tb", [(8, (150, 70, 200), 111, 111), (24, (40, 33, 140), 112, 111), (32, (129, 65, 130), 111, 112), (64, (20, 10, 130), 112, 112)])
def test_gemm_stage1_kernels_identical(pkg, N, shape, ta, tb):
    """vectorised alignment kernel == one-residue-per-thread alignment kernel (whole-result comparison)"""
    ctx = pkg.Context(N, 0)
    orc = get_oracle(N, oracle.DEVICE)
    bits = orc.precision // 4
    m, n, k = shape
    A, B, C = _special_case_inputs(N, m, n, k, bits, 81)
    if ta != 111:
        A = _transpose_recs(A, m, k)
    if tb != 111:
        B = _transpose_recs(B, k, n)
    alpha = random_records(N, 1, bits, 84)
    beta = random_records(N, 1, bits, 85)
    out = []
    for kind in (1, 0, 2, 3):   # round-1 alignment + dense (min,+); default; vectorised + dense (min,+); tensor-core residues in the small-modulus alignment
        ctx.set_stage1_kernel(kind)
        out.append(_gemm(pkg, ctx, m, n, k, alpha, A, B, beta, C, pkg.MODE_AUTO, ta, tb))
    assert all(diff_fields(out[0], o).size == 0 for o in out[1:])
    ctx.close()


@pytest.mark.parametrize("N,bits_div,shape", [(8, 4, (140, 20, 60)), (32, 4, (130, 70, 129)), (32, 8, (64, 64, 300)), (16, 2, (40, 12, 50)),
                                              (24, 4, (33, 5, 20)), (64, 4, (20, 6, 24)), (64, 16, (130, 20, 64))])
def test_gemm_reduced_base_identical(pkg, N, bits_div, shape):
    """Stages 1-2 on the first n' moduli + mixed-radix base extension == stages 1-2 on all N moduli,
    record for record (digits, sign, exp, eval), including zero lines, far exponents and cancellation."""
    ctx = pkg.Context(N, 0)
    orc = get_oracle(N, oracle.DEVICE)
    bits = orc.precision // bits_div - (8 if bits_div == 2 else 0)
    m, n, k = shape
    A, B, C = _special_case_inputs(N, m, n, k, bits, 91)
    alpha = random_records(N, 1, bits, 94)
    beta = random_records(N, 1, bits, 95)
    out, base = [], []
    ctx.set_stage2_kernel(pkg.STAGE2_UMMA)      # the limb planes on the format's own moduli (the default small-modulus path has its own test)
    for on in (False, True):
        ctx.set_reduced_base(on)
        out.append(_gemm(pkg, ctx, m, n, k, alpha, A, B, beta, C, pkg.MODE_AUTO))
        base.append(ctx.last_base_size())
    bad = diff_fields(out[0], out[1])
    assert bad.size == 0, "%d/%d entries differ, first %d\n%s\n%s" % (bad.size, m * n, bad[0], out[0][bad[0]], out[1][bad[0]])
    assert base[0] == N and 1 <= base[1] <= N
    if bits_div >= 4 and N >= 16:
        assert base[1] < N, "p/%d-bit inputs must not need all %d moduli (got %d)" % (bits_div, N, base[1])
    ctx.close()


def test_gemm_workspace_limit_falls_back_to_reference_order(pkg):
    """A fast-path call whose workspaces do not fit the pool's budget is served in reference order (the m x n scratch matrix of
    src/blas/gemm.cuh:98-139) instead of failing with an allocation error; the next call without the cap takes the fast path again."""
    N = 8
    ctx = pkg.Context(N, 0)
    orc = get_oracle(N, oracle.DEVICE)
    bits = orc.precision // 4
    m, n, k = 96, 80, 160
    A = random_records(N, m * k, bits, 71)
    B = random_records(N, k * n, bits, 72)
    C = random_records(N, m * n, bits, 73)
    alpha = random_records(N, 1, bits, 74)
    beta = random_records(N, 1, bits, 75)
    want = _gemm(pkg, ctx, m, n, k, alpha, A, B, beta, C, pkg.MODE_REFERENCE_ORDER)
    held = ctx.workspace_bytes()
    ctx.set_workspace_limit(held + 4096)              # room for nothing beyond the reference-order scratch already held
    got = _gemm(pkg, ctx, m, n, k, alpha, A, B, beta, C, pkg.MODE_AUTO)
    assert ctx.workspace_fallbacks() == 1
    assert diff_fields(got, want).size == 0
    ctx.set_workspace_limit(0)
    got = _gemm(pkg, ctx, m, n, k, alpha, A, B, beta, C, pkg.MODE_AUTO)
    assert ctx.workspace_fallbacks() == 1 and ctx.last_fallback_count() == 0
    assert ctx.workspace_bytes() > held
    assert diff_fields(got, want, ("digits", "sign", "exp")).size == 0
    ctx.close()
