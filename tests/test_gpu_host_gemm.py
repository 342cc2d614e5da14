"""mpres_gemm_host: mp_gemm over HOST mp_float_t[] operands with the transfers pipelined by column panels == the reference caller's
sequence (mp_array_host2device x 3, mp_gemm, mp_array_device2host) record for record, for every panel count, leading dimensions
with padding rows, transposed operands (transposed B runs as one panel), separate and in-place outputs."""
import numpy as np
import pytest

import oracle
from util import diff_fields, get_oracle, random_records

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _device_gemm(pkg, ctx, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc):
    dA, dB, dC = ctx.mp_array_from_host(A), ctx.mp_array_from_host(B), ctx.mp_array_from_host(C)
    dal, dbe = ctx.mp_array_from_host(alpha), ctx.mp_array_from_host(beta)
    pkg.mp_gemm(ctx, ta, tb, m, n, k, dal, dA, lda, dB, ldb, dbe, dC, ldc)
    return dC.device2host()


@pytest.mark.parametrize("N,shape,pads,ta,tb,panels", [(8, (70, 50, 40), (0, 0, 0), 111, 111, 0), (8, (70, 50, 40), (3, 5, 2), 111, 111, 3),
                                                        (32, (130, 600, 64), (0, 0, 0), 111, 111, 0), (32, (40, 37, 50), (1, 0, 7), 112, 111, 37),
                                                        (16, (33, 20, 70), (2, 2, 2), 111, 112, 4), (16, (33, 20, 70), (0, 1, 0), 112, 112, 0),
                                                        (24, (20, 9, 1), (0, 0, 0), 111, 111, 2)])
def test_gemm_host_matches_device_sequence(pkg, N, shape, pads, ta, tb, panels, monkeypatch):
    monkeypatch.setenv("MPRES_HOST_LEAN_MIN", "1")          # the lean upload of A / B also on these small operands
    ctx = pkg.Context(N, 0)
    orc = get_oracle(N, oracle.DEVICE)
    bits = orc.precision // 4
    m, n, k = shape
    ra, ca = (m, k) if ta == 111 else (k, m)
    rb, cb = (k, n) if tb == 111 else (n, k)
    lda, ldb, ldc = ra + pads[0], rb + pads[1], m + pads[2]
    A = random_records(N, lda * ca, bits, 401)
    B = random_records(N, ldb * cb, bits, 402)
    C = random_records(N, ldc * n, bits, 403)
    alpha, beta = random_records(N, 1, bits, 404), random_records(N, 1, bits, 405)
    want = _device_gemm(pkg, ctx, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc)
    out = np.zeros_like(C)
    pkg.mp_gemm_host(ctx, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, out=out, panels=panels)
    assert 0 < ctx.last_host_upload_residues() <= (N // 4 + 1)       # p/4-bit inputs: A and B crossed the link as lean records
    used = np.array([i + j * ldc for j in range(n) for i in range(m)])
    bad = diff_fields(out[used], want[used])
    assert bad.size == 0, "%d/%d entries differ, first %d" % (bad.size, m * n, bad[0])
    # in place, twice (the staging rings and device arrays are reused)
    for _ in range(2):
        Cio = C.copy()
        pkg.mp_gemm_host(ctx, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, Cio, ldc, panels=panels)
        assert diff_fields(Cio[used], want[used]).size == 0
        pad = np.array([i + j * ldc for j in range(n - 1) for i in range(m, ldc)], dtype=np.int64)
        if pad.size:
            assert diff_fields(Cio[pad], C[pad]).size == 0        # padding rows keep the caller's records
    # B already on the device (the multi-GPU recipe: 1/N of B per PCIe link, the rest gathered over NVLink); panels also with transposed B
    dB = ctx.mp_array_from_host(B)
    out[:] = 0
    pkg.mp_gemm_host_bdev(ctx, ta, tb, m, n, k, alpha, A, lda, dB, ldb, beta, C, ldc, out=out, panels=panels)
    bad = diff_fields(out[used], want[used])
    assert bad.size == 0, "device-resident B: %d/%d entries differ, first %d" % (bad.size, m * n, bad[0])
    # reference order through the same entry
    ctx.set_mode(pkg.MODE_REFERENCE_ORDER)
    want_ref = _device_gemm(pkg, ctx, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc)
    pkg.mp_gemm_host(ctx, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, out=out, panels=panels)
    assert diff_fields(out[used], want_ref[used]).size == 0
    ctx.close()


@pytest.mark.parametrize("N,where", [(8, "B_last_panel"), (32, "A"), (16, "B_first_panel")])
def test_gemm_host_lean_upload_falls_back(pkg, N, where, monkeypatch):
    """a few full-precision entries the sample does not see: the packing pass (or the call's own choice on the device) notices, the operands go up in
    full, and the records are those of the device sequence"""
    monkeypatch.setenv("MPRES_HOST_LEAN_MIN", "1")
    ctx = pkg.Context(N, 0)
    orc = get_oracle(N, oracle.DEVICE)
    bits = orc.precision // 4
    m, n, k = 96, 1024, 520
    A, B, C = random_records(N, m * k, bits, 421), random_records(N, k * n, bits, 422), random_records(N, m * n, bits, 423)
    alpha, beta = random_records(N, 1, bits, 424), random_records(N, 1, bits, 425)
    wide = random_records(N, 3, orc.precision // 2, 426)
    if where == "A":
        A[[5, 7777, m * k - 2]] = wide
    elif where == "B_last_panel":
        B[[k * n - 5, k * n - 900, k * (n - 3) + 1]] = wide
    else:
        B[[3, 11, 500]] = wide
    want = _device_gemm(pkg, ctx, 111, 111, m, n, k, alpha, A, m, B, k, beta, C, m)
    out = np.zeros_like(C)
    pkg.mp_gemm_host(ctx, 111, 111, m, n, k, alpha, A, m, B, k, beta, C, m, out=out, panels=4)
    assert diff_fields(out, want).size == 0
    Cio = C.copy()
    pkg.mp_gemm_host(ctx, 111, 111, m, n, k, alpha, A, m, B, k, beta, Cio, m, panels=4)      # in place: C must not have been consumed by a first attempt
    assert diff_fields(Cio, want).size == 0
    ctx.close()


def test_gemm_host_pinned_large_panels(pkg):
    """pinned torch buffers, several staging chunks per operand (> 64 MiB) and eight panels; silent returns and argument errors"""
    N, m, n, k = 8, 1024, 2048, 520
    ctx = pkg.Context(N, 0)
    orc = get_oracle(N, oracle.DEVICE)
    bits = orc.precision // 4
    rs = 4 * N + 40
    A, B, C = random_records(N, m * k, bits, 411), random_records(N, k * n, bits, 412), random_records(N, m * n, bits, 413)
    alpha, beta = random_records(N, 1, bits, 414), random_records(N, 1, bits, 415)
    want = _device_gemm(pkg, ctx, 111, 111, m, n, k, alpha, A, m, B, k, beta, C, m)
    def pin(recs):
        t = torch.empty(recs.size * rs, dtype=torch.uint8).pin_memory()
        t.numpy()[:] = recs.view(np.uint8).reshape(-1)
        return t
    hA, hB, hC, hal, hbe = pin(A), pin(B), pin(C), pin(alpha), pin(beta)
    hOut = torch.zeros(m * n * rs, dtype=torch.uint8).pin_memory()
    pkg.mp_gemm_host(ctx, 111, 111, m, n, k, hal, hA, m, hB, k, hbe, hC, m, out=hOut)
    got = hOut.numpy().view(orc.dtype)
    assert diff_fields(got, want).size == 0
    assert ctx.last_host_upload_residues() > 0
    before = hOut.clone()
    pkg.mp_gemm_host(ctx, 111, 111, 0, n, k, hal, hA, m, hB, k, hbe, hC, m, out=hOut)      # src/blas/gemm.cuh:75-78: silent return
    assert torch.equal(before, hOut)
    with pytest.raises(Exception):
        pkg.mp_gemm_host(ctx, 111, 111, m, n, k, hal, hA, m - 1, hB, k, hbe, hC, m, out=hOut)
    ctx.close()
