"""GPU parity of the scalar device routines against (a) the reference's own cuda:: functions compiled
into oracle/_ref and (b) the DEVICE flavour of the C oracle.  Bit-exact, interval evaluations included."""
import ctypes

import numpy as np
import pytest

import oracle
from util import diff_fields, get_oracle, random_records

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _dev(recs):
    t = torch.from_numpy(recs.view(np.uint8).copy()).cuda()
    return t


def _probe(pkg, ctx, op, x, y=None, bits=None):
    lib = ctx.lib
    dx = _dev(x)
    dy = _dev(y) if y is not None else None
    db = torch.from_numpy(np.ascontiguousarray(bits, dtype=np.int32)).cuda() if bits is not None else None
    dr = torch.zeros_like(dx)
    rc = lib.mpres_probe(ctx.h, op, ctypes.c_void_p(dr.data_ptr()), ctypes.c_void_p(dx.data_ptr()),
                         ctypes.c_void_p(dy.data_ptr()) if dy is not None else None,
                         ctypes.c_void_p(db.data_ptr()) if db is not None else None, ctypes.c_size_t(x.size), None)
    assert rc == 0
    torch.cuda.synchronize()
    return dr.cpu().numpy().view(x.dtype).reshape(x.shape)


def _encloses(orc, rec):
    from fractions import Fraction
    x = Fraction(orc.to_int(rec), orc.c["M"])
    lo = Fraction(float(rec["eval"]["frac"][0])) * Fraction(2) ** int(rec["eval"]["exp"][0])
    up = Fraction(float(rec["eval"]["frac"][1])) * Fraction(2) ** int(rec["eval"]["exp"][1])
    return lo <= x <= up


@pytest.mark.parametrize("N", [8, 16, 24, 32])
@pytest.mark.parametrize("full", [False, True])
def test_scalar_ops_bit_exact(pkg, N, full):
    ctx = pkg.Context(N, 0)
    orc = get_oracle(N, oracle.DEVICE)
    bits = orc.precision if full else orc.precision // 4
    n = 3000
    x = random_records(N, n, bits, 100 + N)
    y = random_records(N, n, bits, 200 + N)
    ref = oracle.RefLib(N, gpu=True) if oracle.have_ref(N) else None
    cases = [(0, x, y, None), (1, x, y, None)]
    prod = orc.mul(x, y)
    cases.append((1, prod, x, None))               # operands with different exponents / magnitudes
    cases.append((1, prod, orc.mul(y, x[::-1].copy()), None))
    cases.append((2, prod, None, None))
    cases.append((3, x, None, None))
    rb = np.random.default_rng(N).integers(1, max(2, bits), size=n).astype(np.int32)
    cases.append((4, x, None, rb))
    for op, a, b, bb in cases:
        got = _probe(pkg, ctx, op, a, b, bb)
        if op == 0:
            want = orc.mul(a, b)
        elif op == 1:
            want = orc.add(a, b)
        elif op in (2, 3):
            want = orc.eval(a, fast=(op == 3))
        else:
            want = orc.round(a, bb)
        bad = diff_fields(got, want, ("digits", "sign", "exp"))
        assert bad.size == 0, "op %d: %d/%d differ from the oracle, first %d:\n%s\n%s" % (op, bad.size, n, bad[0], got[bad[0]], want[bad[0]])
        # interval evaluations: the refinement step k = -(ceil(log2(upp)) + 1) (rns.cuh:912) depends on
        # the libm log2, so the CPU oracle can pick another (equally valid) k than the GPU in rare
        # cases (often, when refining rounding noise: values a few ulps wide sit on powers of two);
        # bit-exactness of eval is asserted against the reference KERNELS below
        bad = diff_fields(got, want)
        for i in bad[:10]:      # a different k must still give an enclosure of X / M
            assert _encloses(orc, got[i])
        if ref is not None:
            r = ref.gpu_probe(op, a, b, bb)
            bad = diff_fields(got, r)
            assert bad.size == 0, "op %d: %d/%d differ from the reference kernels, first %d:\n%s\n%s" % (op, bad.size, n, bad[0], got[bad[0]], r[bad[0]])
    ctx.close()


def test_cancellation_uses_mrc(pkg):
    """x + (-x*(1+tiny)) straddles zero in the interval evaluation -> sign from mixed-radix comparison"""
    N = 8
    ctx = pkg.Context(N, 0)
    orc = get_oracle(N, oracle.DEVICE)
    p = orc.precision
    base = [(1 << (p - 1)) + 12345 * i + 1 for i in range(64)]
    x = orc.set_ints([0] * 64, base, [-p] * 64)
    y = orc.set_ints([1] * 64, [b + (1 if i % 2 else -1) * 2 for i, b in enumerate(base)], [-p] * 64)
    got = _probe(pkg, ctx, 1, x, y)
    want = orc.add(x, y)
    assert diff_fields(got, want).size == 0
    if oracle.have_ref(N):
        r = oracle.RefLib(N, gpu=True).gpu_probe(1, x, y)
        assert diff_fields(got, r).size == 0
    for i in range(64):
        assert orc.to_fraction(got[i]) == orc.to_fraction(x[i]) + orc.to_fraction(y[i])
    ctx.close()
