"""which phase of tests/test_gpu_ops.py::test_spmv_two_stage...[16-90-70-9-False] is slow (prints elapsed seconds per phase)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import _pkg, oracle
from util import get_oracle, random_records
from test_gpu_ops import _random_csr

pkg = _pkg.load()
N, m, n, per_row = (int(v) for v in sys.argv[1:5]) if len(sys.argv) > 4 else (16, 90, 70, 9)
t = time.time()
def lap(what):
    global t
    torch.cuda.synchronize(); print("%-28s %.2f s" % (what, time.time() - t), flush=True); t = time.time()
ctx = pkg.Context(N, 0); orc = get_oracle(N, oracle.DEVICE); bits = orc.precision // 4
irp, ja = _random_csr(m, n, per_row, 1241); nnz = int(irp[-1])
vals = random_records(N, nnz, bits, 1242); x = random_records(N, n, bits, 1243)
dev = torch.device("cuda", 0)
d_irp, d_ja = torch.as_tensor(irp, device=dev), torch.as_tensor(ja, device=dev)
dAs, dx, dy = ctx.mp_collection_from_host(vals), ctx.mp_array_from_host(x), ctx.mp_array_init(m)
lap("setup")
pkg.mp_spmv_mpmtx_csr2st(ctx, m, n, nnz, d_irp, d_ja, dAs, dx, dy); lap("ours csr2st")
pkg.mp_spmv_mpmtx_csr2st(ctx, m, n, nnz, d_irp, d_ja, dAs, dx, dy); lap("ours csr2st again")
ref = oracle.RefLib(N, gpu=True); lap("RefLib")
ref.gpu_spmv_2st(0, m, n, nnz, irp, ja, vals, x); lap("reference csr2st")
ref.gpu_spmv_2st(0, m, n, nnz, irp, ja, vals, x); lap("reference csr2st again")
maxnzr = max(1, int(np.max(np.diff(irp))))
eja = -np.ones((maxnzr, m), dtype=np.int32)
evals = orc.empty((maxnzr, m))
for i in range(m):
    for s, tt in enumerate(range(irp[i], irp[i + 1])):
        eja[s, i] = ja[tt]; evals[s, i] = vals[tt]
d_eja = torch.as_tensor(eja.reshape(-1), device=dev)
dEs, dy2 = ctx.mp_collection_from_host(evals.reshape(-1)), ctx.mp_array_init(m)
lap("ell setup")
pkg.mp_spmv_mpmtx_ell2st(ctx, m, n, maxnzr, d_eja, dEs, dx, dy2); lap("ours ell2st")
ref.gpu_spmv_2st(1, m, n, maxnzr, None, eja.reshape(-1), evals.reshape(-1), x); lap("reference ell2st")
ctx.close(); lap("close")
