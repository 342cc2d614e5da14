"""Row-sharded mp_gemm (mpres_gemm_sharded) on two GPUs: two ranks as two threads of one process (peer access instead of CUDA IPC;
bench.py under torchrun exercises the IPC path), every rank's row block against the single-GPU call and the reference order.
Needs two visible devices: `gpurun --gpus 2 -- python -m pytest tests/test_gpu_sharded.py -m gpu`."""
import threading

import numpy as np
import pytest

import oracle
from util import diff_fields, get_oracle

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _row_block(recs, rows, cols, r0, r1):
    """rows [r0, r1) of a column-major rows x cols array of records, compact"""
    return np.ascontiguousarray(recs.reshape(cols, rows)[:, r0:r1]).reshape(-1)


@pytest.mark.parametrize("N,shape,world", [(8, (512, 512, 384), 2), (32, (384, 300, 200), 2), (16, (256, 1024, 640), 2)])
def test_gemm_sharded_matches_single_gpu(pkg, N, shape, world):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    m, n, k = shape
    orc = get_oracle(N, oracle.DEVICE)
    bits = orc.precision // 4
    A, B, C = orc.random_records(m * k, bits, 801), orc.random_records(k * n, bits, 802), orc.random_records(m * n, bits, 803)
    rng = np.random.RandomState(804)
    A["exp"] += rng.randint(0, 5, size=A.shape).astype(np.int32)
    alpha, beta = orc.random_records(1, bits, 805), orc.random_records(1, bits, 806)
    # single GPU: fast path and reference order
    ctx0 = pkg.Context(N, 0)
    want = {}
    for mode in (pkg.MODE_AUTO, pkg.MODE_REFERENCE_ORDER):
        ctx0.set_mode(mode)
        dA, dB, dC = ctx0.mp_array_from_host(A), ctx0.mp_array_from_host(B), ctx0.mp_array_from_host(C)
        dal, dbe = ctx0.mp_array_from_host(alpha), ctx0.mp_array_from_host(beta)
        pkg.mp_gemm(ctx0, 111, 111, m, n, k, dal, dA, m, dB, k, dbe, dC, m)
        want[mode] = dC.device2host()
    assert diff_fields(want[pkg.MODE_AUTO], want[pkg.MODE_REFERENCE_ORDER], ("digits", "sign", "exp")).size == 0
    ctx0.close()

    handles, results, errors = [None] * world, [None] * world, []
    bar = threading.Barrier(world)

    def rank_main(r):
        try:
            ctx = pkg.Context(N, r)
            r0, r1 = m * r // world, m * (r + 1) // world
            ml = r1 - r0
            dA = ctx.mp_array_from_host(_row_block(A, m, k, r0, r1))
            dB = ctx.mp_array_from_host(B)
            dal, dbe = ctx.mp_array_from_host(alpha), ctx.mp_array_from_host(beta)
            sh = pkg.Shard(ctx, r, world, n, k)
            handles[r] = sh.export()
            bar.wait()
            sh.connect(handles)
            bar.wait()
            out = []
            for rep in range(3):                    # several epochs: the receive buffers and flags are reused
                dC = ctx.mp_array_from_host(_row_block(C, m, n, r0, r1))
                sh.gemm(111, 111, ml, n, k, dal, dA, ml, dB, k, dbe, dC, ml)
                out.append(dC.device2host())
                assert ctx.last_fallback_count() == 0
                P, nin = ctx.last_small_base()
                assert P > 0
            results[r] = out
            bar.wait()
            sh.close()
            ctx.close()
        except Exception as e:                      # noqa: BLE001
            errors.append((r, repr(e)))
            try:
                bar.abort()
            except Exception:
                pass

    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not errors, errors
    for r in range(world):
        r0, r1 = m * r // world, m * (r + 1) // world
        ref = _row_block(want[pkg.MODE_AUTO], m, n, r0, r1)
        for rep, got in enumerate(results[r]):
            bad = diff_fields(got, ref)
            assert bad.size == 0, "rank %d call %d: %d/%d records differ from the single-GPU call, first %d\n%s\n%s" % (r, rep, bad.size, got.size, bad[0], got[bad[0]], ref[bad[0]])


def test_gemm_sharded_transposed_operands(pkg):
    """op(A) = A^T, op(B) = B^T through the flat layout (n / world = 256): the rank's rows of op(A) are columns of its stored block"""
    world, N = 2, 16
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    m, n, k = 192, 512, 320
    orc = get_oracle(N, oracle.DEVICE)
    bits = orc.precision // 4
    A, B, C = orc.random_records(m * k, bits, 811), orc.random_records(k * n, bits, 812), orc.random_records(m * n, bits, 813)
    alpha, beta = orc.random_records(1, bits, 814), orc.random_records(1, bits, 815)
    At = np.ascontiguousarray(A.reshape(k, m).T).reshape(-1)          # stored k x m (column-major): op(A) = At^T
    Bt = np.ascontiguousarray(B.reshape(n, k).T).reshape(-1)          # stored n x k
    ctx0 = pkg.Context(N, 0)
    dC = ctx0.mp_array_from_host(C)
    pkg.mp_gemm(ctx0, 111, 111, m, n, k, ctx0.mp_array_from_host(alpha), ctx0.mp_array_from_host(A), m, ctx0.mp_array_from_host(B), k,
                ctx0.mp_array_from_host(beta), dC, m)
    want = dC.device2host()
    ctx0.close()
    handles, results, errors = [None] * world, [None] * world, []
    bar = threading.Barrier(world)

    def rank_main(r):
        try:
            ctx = pkg.Context(N, r)
            r0, r1 = m * r // world, m * (r + 1) // world
            ml = r1 - r0
            # the rank's rows r0 .. r1 of op(A) = columns r0 .. r1 of the stored k x m array: a contiguous block with lda = k
            dA = ctx.mp_array_from_host(At.reshape(m, k)[r0:r1].reshape(-1))
            dB = ctx.mp_array_from_host(Bt)
            dal, dbe = ctx.mp_array_from_host(alpha), ctx.mp_array_from_host(beta)
            sh = pkg.Shard(ctx, r, world, n, k)
            handles[r] = sh.export()
            bar.wait()
            sh.connect(handles)
            bar.wait()
            dC = ctx.mp_array_from_host(_row_block(C, m, n, r0, r1))
            sh.gemm(112, 112, ml, n, k, dal, dA, k, dB, n, dbe, dC, ml)
            results[r] = dC.device2host()
            assert ctx.last_fallback_count() == 0
            bar.wait()
            sh.close()
            ctx.close()
        except Exception as e:                      # noqa: BLE001
            errors.append((r, repr(e)))
            try:
                bar.abort()
            except Exception:
                pass

    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not errors, errors
    for r in range(world):
        r0, r1 = m * r // world, m * (r + 1) // world
        bad = diff_fields(results[r], _row_block(want, m, n, r0, r1), ("digits", "sign", "exp"))
        assert bad.size == 0, "rank %d: %d records differ" % (r, bad.size)
