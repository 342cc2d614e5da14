"""the reference's ELL two-stage SpMV wrapper after other moduli sets' reference libraries were used in the same process: kernel ms vs wall time"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import oracle
from util import get_oracle, random_records
from test_gpu_ops import _random_csr

torch.cuda.init()
for (N, m, n, per_row, full) in [(8, 60, 50, 7, False), (8, 200, 220, 12, True), (32, 40, 40, 5, True), (16, 90, 70, 9, False), (16, 90, 70, 9, False)]:
    orc = get_oracle(N, oracle.DEVICE); bits = orc.precision if full else orc.precision // 4
    irp, ja = _random_csr(m, n, per_row, 1241); nnz = int(irp[-1])
    vals = random_records(N, nnz, bits, 1242); x = random_records(N, n, bits, 1243)
    maxnzr = max(1, int(np.max(np.diff(irp))))
    eja = -np.ones((maxnzr, m), dtype=np.int32); evals = orc.empty((maxnzr, m))
    for i in range(m):
        for s, tt in enumerate(range(irp[i], irp[i + 1])):
            eja[s, i] = ja[tt]; evals[s, i] = vals[tt]
    ref = oracle.RefLib(N, gpu=True)
    ref.gpu_spmv_2st(0, m, n, nnz, irp, ja, vals, x)
    print("N=%d csr  kernels %.3f ms, call %.3f s" % (N, ref.last_kernel_ms, ref.last_wall_s), flush=True)
    ref.gpu_spmv_2st(1, m, n, maxnzr, None, eja.reshape(-1), evals.reshape(-1), x)
    print("N=%d ell  kernels %.3f ms, call %.3f s" % (N, ref.last_kernel_ms, ref.last_wall_s), flush=True)
