"""GPU parity at the scale the benchmark runs at: the persistent small-modulus stage 2 with SEVERAL tiles per CTA (multi-tile
loop, mbarrier phase wrap across tiles, accumulator hand-off) entry for entry against the reference-order k-loop
(/root/reference/src/blas/gemm.cuh:39-58 is the ground truth of that mode), the device-side input generator
(mpres_array_set_binary) against the oracle, the mp_collection_t entry points and strided x on the fast GEMV / DOT path."""
import ctypes

import numpy as np
import pytest

import oracle
from oracle import gen
from util import diff_fields, get_oracle, random_records

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

SM_COUNT = 148


def _bulk(orc, count, bits, seed):
    return orc.random_records(count, bits, seed)


@pytest.mark.parametrize("N,shape,kind,spread", [
    (8, (2304, 2048, 256), "small", 0), (8, (2304, 2048, 256), "small_t128", 0), (8, (2304, 2048, 256), "small_k64", 6),
    (32, (1280, 1280, 128), "small", 0), (32, (1280, 1280, 128), "small_t128", 9), (32, (1280, 1280, 128), "small_k64", 0),
    (16, (1536, 1792, 384), "small", 4)])
def test_gemm_small_base_many_tiles_per_cta(pkg, N, shape, kind, spread):
    """>= 4 tiles per CTA of k_small_umma_p; every entry compared with REFERENCE_ORDER (digits, sign, exponent)"""
    m, n, k = shape
    ctx = pkg.Context(N, 0)
    orc = get_oracle(N, oracle.DEVICE)
    bits = orc.precision // 4
    A, B, C = _bulk(orc, m * k, bits, 901), _bulk(orc, k * n, bits, 902), _bulk(orc, m * n, bits, 903)
    if spread:
        rng = np.random.RandomState(904)
        A["exp"] += rng.randint(0, spread, size=A.shape).astype(np.int32)
        B["exp"] += rng.randint(0, spread, size=B.shape).astype(np.int32)
    alpha, beta = _bulk(orc, 1, bits, 905), _bulk(orc, 1, bits, 906)
    stage2 = {"small": pkg.STAGE2_SMALL, "small_t128": pkg.STAGE2_SMALL_T128, "small_k64": pkg.STAGE2_SMALL_K64}[kind]
    dA, dB = ctx.mp_array_from_host(A), ctx.mp_array_from_host(B)
    dal, dbe = ctx.mp_array_from_host(alpha), ctx.mp_array_from_host(beta)
    ctx.set_stage2_kernel(stage2)
    ctx.set_mode(pkg.MODE_AUTO)
    dC = ctx.mp_array_from_host(C)
    pkg.mp_gemm(ctx, 111, 111, m, n, k, dal, dA, m, dB, k, dbe, dC, m)
    got = dC.device2host()
    P, nin = ctx.last_small_base()
    assert P > 0 and ctx.last_fallback_count() == 0
    tj = 256 if kind == "small" else 128
    tiles = P * ((m + 255) // 256) * ((n + tj - 1) // tj)
    assert tiles >= 4 * SM_COUNT, "only %.1f tiles per CTA" % (tiles / SM_COUNT)
    ctx.set_mode(pkg.MODE_REFERENCE_ORDER)
    dC2 = ctx.mp_array_from_host(C)
    pkg.mp_gemm(ctx, 111, 111, m, n, k, dal, dA, m, dB, k, dbe, dC2, m)
    want = dC2.device2host()
    bad = diff_fields(got, want, ("digits", "sign", "exp"))
    assert bad.size == 0, "%d/%d entries differ (%.1f tiles per CTA), first %d\n%s\n%s" % (bad.size, m * n, tiles / SM_COUNT, bad[0], got[bad[0]], want[bad[0]])
    ctx.close()


@pytest.mark.parametrize("N", [8, 16, 32, 64])
@pytest.mark.parametrize("full", [False, True])
def test_set_binary_matches_oracle(pkg, N, full):
    """k_set_binary (the generator of every benchmark input) against the oracle's mp_set (mp_set_mpfr semantics,
    /root/reference/src/arith/assign.cuh:86-127): digits, sign, exponent and the interval evaluation"""
    ctx = pkg.Context(N, 0)
    orc = get_oracle(N, oracle.DEVICE)
    bits = orc.precision if full else orc.precision // 4
    count = 3000
    s, mant, e = gen.random_values(count, bits, 911 + N)
    # special cases: zero, a power of two (all trailing bits trimmed), a significand with many trailing zero bits
    mant[0], s[0], e[0] = 0, 0, 0
    mant[1], s[1], e[1] = 1 << (bits - 1), 1, -7
    mant[2], s[2], e[2] = ((1 << 20) | 1) << (bits - 21), 0, 5
    want = orc.set_ints(s, mant, e)
    nl = (bits + 31) // 32
    limbs = np.zeros((count, nl), dtype=np.uint32)
    for i, v in enumerate(mant):
        for w in range(nl):
            limbs[i, w] = (int(v) >> (32 * w)) & 0xFFFFFFFF
    d_s = torch.tensor(np.array(s, dtype=np.int32), device="cuda")
    d_e = torch.tensor(np.array(e, dtype=np.int32), device="cuda")
    d_l = torch.tensor(limbs.view(np.int32), device="cuda")
    arr = ctx.mp_array_init(count + 5)
    pkg._check(ctx.lib.mpres_array_set_binary(ctx.h, ctypes.byref(arr.s), ctypes.c_size_t(5), ctypes.c_void_p(d_s.data_ptr()), ctypes.c_void_p(d_e.data_ptr()),
                                              ctypes.c_void_p(d_l.data_ptr()), nl, ctypes.c_size_t(count), None), "mpres_array_set_binary")
    torch.cuda.synchronize()
    got = arr.device2host()[5:]
    bad = diff_fields(got, want, ("digits", "sign", "exp"))
    assert bad.size == 0, "%d records differ, first %d\n%s\n%s" % (bad.size, bad[0], got[bad[0]], want[bad[0]])
    # interval evaluations: those of the reference's own cuda::rns_eval_compute (src/rns.cuh:797-868) on the same digits, bit for bit
    # (the CPU oracle's refinement step can pick another, equally valid magnification: glibc and CUDA log2 differ in the last ulp)
    if oracle.have_ref(N):
        ref_ev = oracle.RefLib(N, gpu=True).gpu_probe(2, got)
        bad = diff_fields(got, ref_ev, ("eval",))
        assert bad.size == 0, "%d interval evaluations differ from the reference kernel, first %d\n%s\n%s" % (bad.size, bad[0], got[bad[0]], ref_ev[bad[0]])
    from fractions import Fraction
    M = orc.c["M"]
    for i in list(range(0, count, 97)) + [0, 1, 2]:
        X = Fraction(orc.to_int(got[i]), M)
        lo = Fraction(float(got[i]["eval"]["frac"][0])) * Fraction(2) ** int(got[i]["eval"]["exp"][0])
        up = Fraction(float(got[i]["eval"]["frac"][1])) * Fraction(2) ** int(got[i]["eval"]["exp"][1])
        assert lo <= X <= up, i
    ctx.close()


@pytest.mark.parametrize("N", [8, 32])
def test_gemm_and_gemv_collection_entry_points(pkg, N):
    """mpres_gemm_coll / mpres_gemv_coll: mp_collection_t operands with explicit lengths give the records of the mp_array_t calls"""
    ctx = pkg.Context(N, 0)
    orc = get_oracle(N, oracle.DEVICE)
    bits = orc.precision // 4
    m, n, k = 70, 45, 130
    A, B, C = random_records(N, m * k, bits, 921), random_records(N, k * n, bits, 922), random_records(N, m * n, bits, 923)
    alpha, beta = random_records(N, 1, bits, 924), random_records(N, 1, bits, 925)
    for mode in (pkg.MODE_AUTO, pkg.MODE_REFERENCE_ORDER):
        ctx.set_mode(mode)
        dA, dB, dC = ctx.mp_array_from_host(A), ctx.mp_array_from_host(B), ctx.mp_array_from_host(C)
        dal, dbe = ctx.mp_array_from_host(alpha), ctx.mp_array_from_host(beta)
        pkg.mp_gemm(ctx, 111, 111, m, n, k, dal, dA, m, dB, k, dbe, dC, m)
        want = dC.device2host()
        cA, cB, cC = ctx.mp_collection_from_host(A), ctx.mp_collection_from_host(B), ctx.mp_collection_from_host(C)
        cal, cbe = ctx.mp_collection_from_host(alpha), ctx.mp_collection_from_host(beta)
        pkg.mp_gemm(ctx, 111, 111, m, n, k, cal, cA, m, cB, k, cbe, cC, m)
        got = cC.device2host()
        assert diff_fields(got, want).size == 0, "gemm_coll mode %d" % mode
        for trans in (111, 112):
            lenx, leny = (k, m) if trans == 111 else (m, k)
            x, y = random_records(N, lenx, bits, 926), random_records(N, leny, bits, 927)
            dx, dy = ctx.mp_array_from_host(x), ctx.mp_array_from_host(y)
            pkg.mp_gemv(ctx, trans, m, k, dal, dA, m, dx, 1, dbe, dy, 1)
            want_y = dy.device2host()
            cx, cy = ctx.mp_collection_from_host(x), ctx.mp_collection_from_host(y)
            pkg.mp_gemv(ctx, trans, m, k, cal, cA, m, cx, 1, cbe, cy, 1)
            assert diff_fields(cy.device2host(), want_y).size == 0, "gemv_coll mode %d trans %d" % (mode, trans)
    ctx.close()


def _strided(vals, inc, zero):
    """BLAS storage of a logical vector with increment inc (negative: reversed, src/mpvector.cuh:68-70)"""
    n = len(vals)
    out = np.full((n - 1) * abs(inc) + 1, zero, dtype=vals.dtype)
    pos = np.arange(n) * inc if inc > 0 else (-n + np.arange(n) + 1) * inc
    out[pos] = vals
    return out


@pytest.mark.parametrize("N", [8, 16])
@pytest.mark.parametrize("incx,incy", [(2, 1), (-3, 1), (2, -3), (1, 2)])
def test_dot_and_gemv_fast_path_strided_x(pkg, N, incx, incy):
    """incx = 2, -3 on the single-pass exact-window kernels (k_mv_acc_n / k_mv_acc_t): same digits, sign, exponent as unit strides"""
    ctx = pkg.Context(N, 0)
    orc = get_oracle(N, oracle.DEVICE)
    bits = orc.precision // 4
    zero = orc.set_ints([0], [0], [0])[0]
    n = 7001
    x, y = random_records(N, n, bits, 931), random_records(N, n, bits, 932)
    ctx.set_mode(pkg.MODE_REFERENCE_ORDER)
    dx, dy, dr = ctx.mp_array_from_host(x), ctx.mp_array_from_host(y), ctx.mp_array_init(1)
    pkg.mp_dot(ctx, n, dx, 1, dy, 1, dr)
    want = dr.device2host()[0]
    ctx.set_mode(pkg.MODE_AUTO)
    sx, sy = ctx.mp_array_from_host(_strided(x, incx, zero)), ctx.mp_array_from_host(_strided(y, incy, zero))
    dr2 = ctx.mp_array_init(1)
    pkg.mp_dot(ctx, n, sx, incx, sy, incy, dr2)
    got = dr2.device2host()[0]
    assert ctx.last_fallback_count() == 0
    assert diff_fields(np.array([got]), np.array([want]), ("digits", "sign", "exp")).size == 0, (got, want)
    m, nn = 130, 90
    A = random_records(N, m * nn, bits, 933)
    alpha, beta = random_records(N, 1, bits, 934), random_records(N, 1, bits, 935)
    dA, dal, dbe = ctx.mp_array_from_host(A), ctx.mp_array_from_host(alpha), ctx.mp_array_from_host(beta)
    for trans in (111, 112):
        lenx, leny = (nn, m) if trans == 111 else (m, nn)
        xv, yv = random_records(N, lenx, bits, 936), random_records(N, leny, bits, 937)
        ctx.set_mode(pkg.MODE_REFERENCE_ORDER)
        d1, d2 = ctx.mp_array_from_host(xv), ctx.mp_array_from_host(yv)
        pkg.mp_gemv(ctx, trans, m, nn, dal, dA, m, d1, 1, dbe, d2, 1)
        want_y = d2.device2host()
        ctx.set_mode(pkg.MODE_AUTO)
        s1, s2 = ctx.mp_array_from_host(_strided(xv, incx, zero)), ctx.mp_array_from_host(_strided(yv, incy, zero))
        pkg.mp_gemv(ctx, trans, m, nn, dal, dA, m, s1, incx, dbe, s2, incy)
        assert ctx.last_fallback_count() == 0
        full = s2.device2host()
        pos = np.arange(leny) * incy if incy > 0 else (-leny + np.arange(leny) + 1) * incy
        assert diff_fields(full[pos], want_y, ("digits", "sign", "exp")).size == 0, (trans, incx, incy)
    ctx.close()
