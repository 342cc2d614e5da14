"""mp_asum, mp_norm, mp_ge_norm, the two-stage SpMV over mp_collection_t (rank 4) and the double conversions (rank 1)
against the oracle's mp_mul / mp_add sequences, the reference's own CUDA kernels (oracle/_ref) and exact rationals."""
import struct
from fractions import Fraction

import numpy as np
import pytest

import oracle
from util import diff_fields, get_oracle, random_records, unit_roundoff

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _abs(recs):
    out = recs.copy()
    out["sign"] = 0
    return out


def _seq_sum(orc, recs):
    """mp_add chain from MP_ZERO in index order (DEVICE oracle)"""
    acc = orc.set_ints([0], [0], [0])
    for r in recs:
        acc = orc.add(acc, r.reshape(1))
    return acc[0]


def _scalar(ctx):
    return ctx.mp_array_init(1)


@pytest.mark.parametrize("N,n,incx", [(8, 1, 1), (8, 777, 1), (16, 5000, 1), (32, 1300, 1), (8, 300, 3), (24, 257, 2)])
def test_asum_quarter_precision_bit_exact(pkg, N, n, incx):
    """p/4-bit inputs: no addition of the reference rounds, so any summation order gives its digits, sign and exponent"""
    ctx = pkg.Context(N, 0)
    orc = get_oracle(N, oracle.DEVICE)
    x = random_records(N, n * incx, orc.precision // 4, 1201)
    x[::7] = orc.set_ints([0], [0], [0])[0]
    dx, r = ctx.mp_array_from_host(x), _scalar(ctx)
    pkg.mp_asum(ctx, n, dx, incx, r)
    got = r.device2host()[0]
    want = _seq_sum(orc, _abs(x[::incx][:n]))
    assert diff_fields(np.array([got]), np.array([want]), ("digits", "sign", "exp")).size == 0, (got, want)
    if oracle.have_ref(N) and incx == 1:
        ref = oracle.RefLib(N, gpu=True).gpu_asum_norm(0, x, n, incx)
        assert diff_fields(np.array([got]), np.array([ref]), ("digits", "sign", "exp")).size == 0
    # one norm = the same
    pkg.mp_norm(ctx, pkg.mblas_one_norm, n, dx, incx, r)
    assert diff_fields(r.device2host(), np.array([want]), ("digits", "sign", "exp")).size == 0
    ctx.close()


@pytest.mark.parametrize("N,n", [(8, 4000), (32, 900)])
def test_asum_full_precision_within_the_error_model(pkg, N, n):
    ctx = pkg.Context(N, 0)
    orc = get_oracle(N, oracle.DEVICE)
    x = random_records(N, n, orc.precision, 1211)
    dx, r = ctx.mp_array_from_host(x), _scalar(ctx)
    pkg.mp_asum(ctx, n, dx, 1, r)
    got = orc.to_fraction(r.device2host()[0])
    exact = sum(abs(orc.to_fraction(v)) for v in x)
    u = unit_roundoff(orc)
    assert abs(got - exact) <= n * u * exact
    ctx.close()


@pytest.mark.parametrize("N,n,incx", [(8, 1, 1), (8, 3000, 1), (16, 2049, 2), (32, 700, 1)])
def test_norm_inf_is_the_largest_magnitude(pkg, N, n, incx):
    ctx = pkg.Context(N, 0)
    orc = get_oracle(N, oracle.DEVICE)
    x = random_records(N, n * incx, orc.precision // 2, 1221)
    rng = np.random.RandomState(1222)
    x["exp"] += rng.randint(-3, 4, size=x.shape).astype(np.int32)
    if n > 10:
        x[4 * incx] = x[9 * incx]                     # a tie in magnitude
        x[4 * incx]["sign"] ^= 1
    dx, r = ctx.mp_array_from_host(x), _scalar(ctx)
    pkg.mp_norm(ctx, pkg.mblas_inf_norm, n, dx, incx, r)
    got = r.device2host()[0]
    vals = [abs(orc.to_fraction(v)) for v in x[::incx][:n]]
    assert int(got["sign"]) == 0 and orc.to_fraction(got) == max(vals)
    # the result is one of the elements, copied
    hits = [i for i, v in enumerate(vals) if v == max(vals)]
    assert any(diff_fields(np.array([got]), _abs(x[::incx][i:i + 1]), ("digits", "exp")).size == 0 for i in hits)
    if oracle.have_ref(N) and incx == 1:
        ref = oracle.RefLib(N, gpu=True).gpu_asum_norm(175, x, n, incx)
        assert orc.to_fraction(ref) == orc.to_fraction(got)
    ctx.close()


@pytest.mark.parametrize("N,m,n,lda", [(8, 33, 47, 33), (8, 300, 200, 310), (16, 64, 1025, 64), (32, 130, 70, 130)])
def test_ge_norm(pkg, N, m, n, lda):
    ctx = pkg.Context(N, 0)
    orc = get_oracle(N, oracle.DEVICE)
    A = random_records(N, lda * n, orc.precision // 4, 1231)
    dA, r = ctx.mp_array_from_host(A), _scalar(ctx)
    Am = A.reshape(n, lda)[:, :m]                      # [column][row]
    for norm in (pkg.mblas_one_norm, pkg.mblas_inf_norm):
        lines = [Am[j, :] for j in range(n)] if norm == pkg.mblas_one_norm else [Am[:, i] for i in range(m)]
        sums = np.array([_seq_sum(orc, _abs(np.ascontiguousarray(line))) for line in lines])
        vals = [orc.to_fraction(s) for s in sums]
        best = max(vals)
        for buf in (None, ctx.mp_array_init(max(m, n))):
            pkg.mp_ge_norm(ctx, norm, m, n, dA, lda, r, buf)
            got = r.device2host()[0]
            assert orc.to_fraction(got) == best, (norm, float(orc.to_fraction(got)), float(best))
            assert any(diff_fields(np.array([got]), sums[i:i + 1], ("digits", "sign", "exp")).size == 0 for i, v in enumerate(vals) if v == best)
        if oracle.have_ref(N) and lda == m:
            ref = oracle.RefLib(N, gpu=True).gpu_ge_norm(norm, m, n, A, lda)
            assert orc.to_fraction(ref) == best
    ctx.close()


def _random_csr(m, n, per_row, seed):
    rng = np.random.RandomState(seed)
    irp, ja = [0], []
    for i in range(m):
        cnt = int(rng.randint(0, per_row + 1))
        cols = sorted(rng.choice(n, size=min(cnt, n), replace=False).tolist())
        ja += cols
        irp.append(len(ja))
    return np.array(irp, dtype=np.int32), np.array(ja, dtype=np.int32)


@pytest.mark.parametrize("N,m,n,per_row,full", [(8, 60, 50, 7, False), (8, 200, 220, 12, True), (32, 40, 40, 5, True), (16, 90, 70, 9, False)])
def test_spmv_two_stage_matches_the_reference_sequence(pkg, N, m, n, per_row, full):
    ctx = pkg.Context(N, 0)
    orc = get_oracle(N, oracle.DEVICE)
    bits = orc.precision if full else orc.precision // 4
    irp, ja = _random_csr(m, n, per_row, 1241)
    nnz = int(irp[-1])
    vals = random_records(N, max(nnz, 1), bits, 1242)[:nnz]
    x = random_records(N, n, bits, 1243)
    # reference sequence per row: round(a x_j), sum = round(sum + product) from MP_ZERO
    want = orc.empty(m)
    for i in range(m):
        acc = orc.set_ints([0], [0], [0])
        for t in range(irp[i], irp[i + 1]):
            acc = orc.add(acc, orc.mul(vals[t:t + 1], x[ja[t]:ja[t] + 1]))
        want[i] = acc[0]
    dev = torch.device("cuda", 0)
    d_irp, d_ja = torch.as_tensor(irp, device=dev), torch.as_tensor(ja if nnz else np.zeros(1, np.int32), device=dev)
    dAs, dx, dy = ctx.mp_collection_from_host(vals if nnz else orc.empty(1)), ctx.mp_array_from_host(x), ctx.mp_array_init(m)
    pkg.mp_spmv_mpmtx_csr2st(ctx, m, n, nnz, d_irp, d_ja, dAs, dx, dy)
    got = dy.device2host()
    bad = diff_fields(got, want, ("digits", "sign", "exp"))
    assert bad.size == 0, "CSR: %d/%d rows differ, first %d\n%s\n%s" % (bad.size, m, bad[0], got[bad[0]], want[bad[0]])
    # ELLPACK: column-major m x maxnzr, padding ja = -1
    maxnzr = max(1, int(np.max(np.diff(irp))))
    eja = -np.ones((maxnzr, m), dtype=np.int32)
    evals = orc.empty((maxnzr, m))
    for i in range(m):
        for s, t in enumerate(range(irp[i], irp[i + 1])):
            eja[s, i] = ja[t]
            evals[s, i] = vals[t]
    d_eja = torch.as_tensor(eja.reshape(-1), device=dev)
    dEs, dy2 = ctx.mp_collection_from_host(evals.reshape(-1)), ctx.mp_array_init(m)
    pkg.mp_spmv_mpmtx_ell2st(ctx, m, n, maxnzr, d_eja, dEs, dx, dy2)
    bad = diff_fields(dy2.device2host(), want, ("digits", "sign", "exp"))
    assert bad.size == 0, "ELLPACK: %d/%d rows differ" % (bad.size, m)
    if oracle.have_ref(N) and nnz:
        ref = oracle.RefLib(N, gpu=True)
        r1 = ref.gpu_spmv_2st(0, m, n, nnz, irp, ja, vals, x)
        assert diff_fields(got, r1, ("digits", "sign", "exp")).size == 0
        r2 = ref.gpu_spmv_2st(1, m, n, maxnzr, None, eja.reshape(-1), evals.reshape(-1), x)
        assert diff_fields(got, r2, ("digits", "sign", "exp")).size == 0
    ctx.close()


@pytest.mark.parametrize("N", [8, 32, 64])
def test_double_conversions(pkg, N):
    ctx = pkg.Context(N, 0)
    orc = get_oracle(N, oracle.DEVICE)
    rng = np.random.RandomState(1251)
    vals = np.concatenate([rng.standard_normal(500) * 10.0 ** rng.randint(-30, 30, 500), [0.0, 1.0, -1.0, 0.5, 3.0, -2.0 ** 60, 2.0 ** -80, 1e300, -1e-300]])
    dev = torch.device("cuda", 0)
    src = torch.as_tensor(vals, device=dev)
    arr = ctx.mp_array_init(vals.size + 3)
    pkg.mp_array_set_d(ctx, arr, 3, src, vals.size)
    got = arr.device2host()[3:]
    M = orc.c["M"]
    for v, g in zip(vals, got):
        fr = Fraction(float(v))
        assert orc.to_fraction(g) == fr, (v, g)
        x = orc.to_int(g)
        assert x % 2 == 1 or x == 0                     # trailing zeros trimmed into the exponent (assign.cuh:69-76)
        if x:
            lo = Fraction(float(g["eval"]["frac"][0])) * Fraction(2) ** int(g["eval"]["exp"][0])
            up = Fraction(float(g["eval"]["frac"][1])) * Fraction(2) ** int(g["eval"]["exp"][1])
            assert lo <= Fraction(x, M) <= up
    # ... and back: exact for doubles
    back = torch.zeros(vals.size, dtype=torch.float64, device=dev)
    pkg.mp_array_get_d(ctx, back, arr, 3, vals.size)
    assert np.array_equal(back.cpu().numpy().view(np.int64), vals.view(np.int64))
    # wide significands: round to nearest even of the exact value
    recs = random_records(N, 400, orc.precision, 1252)
    rng2 = np.random.RandomState(1253)
    recs["exp"] += rng2.randint(-40, 40, size=recs.shape).astype(np.int32)
    d = ctx.mp_array_from_host(recs)
    out = torch.zeros(400, dtype=torch.float64, device=dev)
    pkg.mp_array_get_d(ctx, out, d, 0, 400)
    o = out.cpu().numpy()
    for r, v in zip(recs, o):
        fr = orc.to_fraction(r)
        want = fr.numerator / fr.denominator              # Python: correctly rounded (ties to even) division of integers
        assert struct.pack("<d", v) == struct.pack("<d", want), (v, want)
    ctx.close()


def _rne(num, den, prec):
    """(mantissa, exponent) of num / den rounded to nearest even at prec bits, trailing zeros trimmed (num, den > 0)"""
    L = num.bit_length() - den.bit_length()
    sh = prec + 2 - L
    n2, d2 = (num << sh, den) if sh >= 0 else (num, den << -sh)
    q, r = divmod(n2, d2)
    e = -sh
    d = q.bit_length() - prec
    if d > 0:
        half, rem = 1 << (d - 1), q & ((1 << d) - 1)
        q >>= d
        if rem > half or (rem == half and (r or (q & 1))):
            q += 1
        elif rem == half and not r and not (q & 1):
            pass
        e += d
    elif r:
        raise AssertionError("quotient shorter than the precision with a remainder")
    while q and not q & 1:
        q >>= 1
        e += 1
    return q, e


@pytest.mark.parametrize("N", [8, 16, 32, 64])
def test_div_is_correctly_rounded(pkg, N):
    ctx = pkg.Context(N, 0)
    orc = get_oracle(N, oracle.DEVICE)
    prec = orc.precision
    xs = random_records(N, 24, prec, 1261)
    ys = random_records(N, 24, prec // 3, 1262)
    xs[3] = orc.set_ints([1], [12], [5])[0]; ys[3] = orc.set_ints([0], [3], [-2])[0]       # exact quotient
    xs[4] = orc.set_ints([0], [1], [0])[0]; ys[4] = orc.set_ints([0], [3], [0])[0]        # 1 / 3
    xs[5] = orc.set_ints([0], [0], [0])[0]                                                   # 0 / y
    mods = orc.c["moduli"]
    r = ctx.mp_array_init(1)
    for x, y in zip(xs, ys):
        pkg.mp_div(ctx, r, ctx.mp_array_from_host(x.reshape(1)), ctx.mp_array_from_host(y.reshape(1)))
        g = r.device2host()[0]
        xi, yi = orc.to_int(x), orc.to_int(y)
        if xi == 0:
            assert not g["digits"].any() and g["eval"]["frac"][1] == 0
            continue
        q, e = _rne(xi, yi, prec)
        assert int(g["sign"]) == int(x["sign"]) ^ int(y["sign"])
        assert int(g["exp"]) == int(x["exp"]) - int(y["exp"]) + e, (g, e)
        assert [int(d) for d in g["digits"]] == [q % m for m in mods]
        lo = Fraction(float(g["eval"]["frac"][0])) * Fraction(2) ** int(g["eval"]["exp"][0])
        up = Fraction(float(g["eval"]["frac"][1])) * Fraction(2) ** int(g["eval"]["exp"][1])
        assert lo <= Fraction(q, orc.c["M"]) <= up
    ctx.close()


@pytest.mark.parametrize("N,precond", [(8, False), (32, False), (32, True)])
def test_conjugate_gradients(pkg, N, precond):
    """a 1-D beam-like SPD band matrix (the structure of the reference's fixture tests/sparse/matrices/LF10.mtx): in exact arithmetic CG ends after
    n steps; at 106 / 424 bits the residual must fall far below double precision and the solution must match the exact rational one"""
    ctx = pkg.Context(N, 0)
    orc = get_oracle(N, oracle.DEVICE)
    n = 24
    rng = np.random.RandomState(1271)
    A = np.zeros((n, n))
    for i in range(n):
        A[i, i] = 4.0 + rng.randint(0, 8) / 4.0
        if i + 1 < n:
            A[i, i + 1] = A[i + 1, i] = -1.0 - rng.randint(0, 4) / 8.0
        if i + 2 < n:
            A[i, i + 2] = A[i + 2, i] = 0.25
    bvec = rng.randint(-8, 9, size=n) / 4.0
    irp, ja, vals = [0], [], []
    for i in range(n):
        for j in range(n):
            if A[i, j] != 0:
                ja.append(j); vals.append(A[i, j])
        irp.append(len(ja))
    dev = torch.device("cuda", 0)
    d_irp = torch.as_tensor(np.array(irp, dtype=np.int32), device=dev)
    d_ja = torch.as_tensor(np.array(ja, dtype=np.int32), device=dev)
    d_vals = torch.as_tensor(np.array(vals, dtype=np.float64), device=dev)
    b = ctx.mp_array_init(n)
    pkg.mp_array_set_d(ctx, b, 0, torch.as_tensor(bvec, device=dev), n)
    x = ctx.mp_array_init(n)
    pkg.mp_array_set_d(ctx, x, 0, torch.zeros(n, dtype=torch.float64, device=dev), n)
    M = torch.as_tensor(1.0 / np.diag(A), device=dev) if precond else None
    tol = 2.0 ** -(orc.precision - 30)
    iters, res = pkg.mp_cg_csr(ctx, n, len(ja), d_irp, d_ja, d_vals, b, tol, 3 * n, x, M)
    assert 0 < iters <= 3 * n and res[-1] <= tol, (iters, res[-3:])
    # exact solution by rational Gaussian elimination
    Af = [[Fraction(float(v)) for v in row] for row in A]
    bf = [Fraction(float(v)) for v in bvec]
    for c in range(n):
        piv = Af[c][c]
        for rr in range(c + 1, n):
            f = Af[rr][c] / piv
            if f:
                Af[rr] = [a - f * p for a, p in zip(Af[rr], Af[c])]
                bf[rr] -= f * bf[c]
    sol = [Fraction(0)] * n
    for c in range(n - 1, -1, -1):
        sol[c] = (bf[c] - sum(Af[c][j] * sol[j] for j in range(c + 1, n))) / Af[c][c]
    got = [orc.to_fraction(v) for v in x.device2host()]
    err = max(abs(g - s) for g, s in zip(got, sol)) / max(abs(s) for s in sol)
    assert err < Fraction(2) ** -(orc.precision - 40), float(err)
    ctx.close()
