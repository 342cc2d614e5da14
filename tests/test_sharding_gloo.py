"""world_size-2 gloo runs of the multi-GPU host logic (mpres-blas_b200/parallel.py) with the C oracle
standing in for the CUDA kernels: the sharded results must equal the single-process results bit for bit
(p/4-bit inputs: nothing rounds, so the partition cannot change digits, sign or exponent)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

torch = pytest.importorskip("torch")


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import _pkg
    import oracle
    from util import diff_fields
    pkg = _pkg.load()
    from mpres_blas_b200 import parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    N = 8
    orc = oracle.Oracle(N, oracle.DEVICE)
    bits = orc.precision // 4
    m, n, k = 12, 7, 20
    A, B, C = orc.random_records(m * k, bits, 1), orc.random_records(k * n, bits, 2), orc.random_records(m * n, bits, 3)
    al, be = orc.random_records(1, bits, 4), orc.random_records(1, bits, 5)
    # --- GEMM: row blocks of A and C, B broadcast from rank 0 as raw SoA byte tensors
    lo, hi = parallel.row_block(m, world, rank)
    Am = A.reshape(k, m)[:, lo:hi].copy().reshape(-1)          # compact shard, lda = hi - lo
    Cm = C.reshape(n, m)[:, lo:hi].copy().reshape(-1)
    Bt = torch.from_numpy((B if rank == 0 else np.zeros_like(B)).view(np.uint8).copy())

    def local_gemm():
        Bl = np.frombuffer(Bt.numpy().tobytes(), dtype=orc.dtype)
        out, _ = orc.gemm(hi - lo, n, k, al, Am, Bl, be, Cm)
        return out
    mine = parallel.gemm_row_sharded(dist, [Bt], local_gemm)
    full, _ = orc.gemm(m, n, k, al, A, B, be, C)
    want = full.reshape(n, m)[:, lo:hi].reshape(-1)
    assert diff_fields(mine, want).size == 0
    # --- DOT: segments, packed partials all-gathered and reduced in rank order
    nvec = 1001
    x, y = orc.random_records(nvec, bits, 6), orc.random_records(nvec, bits, 7)
    s0, s1 = parallel.segment(nvec, world, rank)

    def local_partial():
        return torch.from_numpy(np.array([orc.dot_seq(x[s0:s1], y[s0:s1])]).view(np.uint8).copy())

    def reduce_partials(buf, count):
        recs = np.frombuffer(buf.numpy().tobytes(), dtype=orc.dtype)
        acc = orc.empty(1)
        for i in range(count):
            acc = orc.add(acc, recs[i:i + 1])
        return acc[0]
    r = parallel.dot_segment_sharded(dist, local_partial, reduce_partials)
    whole = orc.dot_seq(x, y)
    assert diff_fields(np.array([r]), np.array([whole]), ("digits", "sign", "exp")).size == 0
    # every rank must hold identical bits
    mine_b = torch.from_numpy(np.array([r]).view(np.uint8).copy())
    both = [torch.empty_like(mine_b) for _ in range(world)]
    dist.all_gather(both, mine_b)
    assert all(torch.equal(both[0], b) for b in both)
    # --- GEMV (T): row-sharded A and x, partial y per rank, gathered and summed in rank order
    xv, yv = orc.random_records(m, bits, 8), orc.random_records(k, bits, 9)
    Ak = orc.random_records(m * k, bits, 10)                      # m x k, y has k entries for trans = T
    zero_beta = orc.empty(1)
    zero_beta["digits"][:] = 0

    def local_partial_y():
        Al = Ak.reshape(k, m)[:, lo:hi].copy().reshape(-1)
        out = orc.gemv(112, hi - lo, k, al, Al, xv[lo:hi], zero_beta, orc.empty(k), block=1)
        return torch.from_numpy(out.view(np.uint8).copy())

    def reduce_columns(parts):
        acc = orc.mul(yv, np.repeat(be, k))
        for p_ in parts:
            acc = orc.add(acc, np.frombuffer(p_.numpy().tobytes(), dtype=orc.dtype))
        return acc
    yt = parallel.gemv_t_sharded(dist, local_partial_y, reduce_columns)
    want_y = orc.gemv(112, m, k, al, Ak, xv, be, yv, block=1)
    assert diff_fields(yt, want_y, ("digits", "sign", "exp")).size == 0
    # --- lean broadcast of B: only the first n_in residues, sign, exponent and the upper bounds travel
    lb = parallel.LeanBroadcast(dist, N)
    assert lb.agree(2 if rank == 0 else 3, "cpu") == 3
    cnt = k * n
    src_d = torch.from_numpy(np.ascontiguousarray(B["digits"]).reshape(-1).copy())
    src_s, src_e = torch.from_numpy(B["sign"].copy()), torch.from_numpy(B["exp"].copy())
    ev_np = np.concatenate([B["eval"][:, 0].copy().view(np.int64).reshape(-1), B["eval"][:, 1].copy().view(np.int64).reshape(-1)])
    src_ev = torch.from_numpy(ev_np.copy())
    if rank == 0:
        d, sg, ex, ev = src_d.clone(), src_s.clone(), src_e.clone(), src_ev.clone()
    else:
        d, sg, ex, ev = torch.full_like(src_d, -7), torch.zeros_like(src_s), torch.zeros_like(src_e), torch.full_like(src_ev, -7)
    lb.broadcast(d, sg, ex, ev)
    assert torch.equal(d.view(cnt, N)[:, :3], src_d.view(cnt, N)[:, :3]) and torch.equal(sg, src_s) and torch.equal(ex, src_e)
    assert torch.equal(ev[2 * cnt:], src_ev[2 * cnt:])
    if rank != 0:      # nothing else was moved
        assert bool((d.view(cnt, N)[:, 3:] == -7).all()) and bool((ev[: 2 * cnt] == -7).all())
    assert lb.nbytes(cnt) == cnt * (4 * 3 + 24)
    assert lb.verify(40, 3, 0) and not lb.verify(0, 0, 0) and not lb.verify(40, 4, 0) and not lb.verify(40, 3, 2)
    # --- end-to-end recipe: every rank holds its column block of B (k x n2, n2 % world == 0), the five SoA ranges are all-gathered
    n2 = 6
    B2 = orc.random_records(k * n2, bits, 11)
    cnt2, per = k * n2, k * n2 // world
    full_d = torch.from_numpy(np.ascontiguousarray(B2["digits"]).reshape(-1).copy())
    full_s, full_e = torch.from_numpy(B2["sign"].copy()), torch.from_numpy(B2["exp"].copy())
    full_ev = torch.from_numpy(np.concatenate([B2["eval"][:, 0].copy().view(np.int64).reshape(-1), B2["eval"][:, 1].copy().view(np.int64).reshape(-1)]).copy())
    mine_t = [torch.full_like(full_d, -7), torch.full_like(full_s, -7), torch.full_like(full_e, -7), torch.full_like(full_ev, -7)]
    for dst, src in zip(parallel.soa_ranges(*mine_t, N, cnt2, per * rank, per), parallel.soa_ranges(full_d, full_s, full_e, full_ev, N, cnt2, per * rank, per)):
        dst.copy_(src)                                             # "uploaded from host memory": only this rank's columns
    parallel.gather_column_shards(dist, parallel.soa_ranges(*mine_t, N, cnt2, 0, cnt2), parallel.soa_ranges(*mine_t, N, cnt2, per * rank, per))
    assert all(torch.equal(a, b) for a, b in zip(mine_t, [full_d, full_s, full_e, full_ev]))
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(tmp, "ok%d" % rank), "w").write("ok")


def test_sharded_paths_world2(tmp_path):
    import torch.multiprocessing as mp
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")


def test_partition_helpers(pkg):
    from mpres_blas_b200 import parallel
    for m in (1, 7, 4096, 4097):
        for w in (1, 2, 4, 8):
            blocks = [parallel.row_block(m, w, r) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == m
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
