"""Quick timing probe of the single-pass mp_gemv / mp_dot kernels (GPU box).  Not the bench: prints stage times."""
import sys, os, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import _pkg
pkg = _pkg.load()
from mpres_blas_b200 import torch_arrays as ta

def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

CFGS = [int(x) for x in os.environ.get('VEC_CFGS', '0,1,2').split(',')]

def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    m = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 16384
    nd = int(sys.argv[4]) if len(sys.argv) > 4 else 1 << 24
    ctx = pkg.Context(N, 0)
    bits = ctx.precision // 4
    rs = 4 * N + 40
    stream = torch.cuda.current_stream().cuda_stream
    out = {"N": N}
    # DOT
    x, y, r = ta.TorchMpArray(ctx, nd), ta.TorchMpArray(ctx, nd), ta.TorchMpArray(ctx, 1)
    ta.random_fill(ctx, x, bits, 1); ta.random_fill(ctx, y, bits, 2)
    ctx.set_profiling(True)
    for cfg in CFGS:
        ctx.set_vec_config(cfg)
        ms = timeit(lambda: pkg.mp_dot(ctx, nd, x, 1, y, 1, r, None, stream))
        st, _ = ctx.last_stage_ms()
        out["dot"] = {"cfg": cfg, "n": nd, "ms": ms, "stage_ms": st, "GBs_alg": 2 * nd * rs / ms / 1e6, "GBs_kernel_touched": 2 * nd * (rs - 16) / st[1] / 1e6, "fallback": ctx.last_fallback_count()}
        print(json.dumps(out["dot"]), flush=True)
    del x, y
    torch.cuda.empty_cache()
    A = ta.TorchMpArray(ctx, m * n)
    ta.random_fill(ctx, A, bits, 3)
    al, be = ta.TorchMpArray(ctx, 1), ta.TorchMpArray(ctx, 1)
    ta.random_fill(ctx, al, bits, 4); ta.random_fill(ctx, be, bits, 5)
    for trans, name in ((111, "gemv_n"), (112, "gemv_t")):
        lenx, leny = (n, m) if trans == 111 else (m, n)
        xv, yv = ta.TorchMpArray(ctx, lenx), ta.TorchMpArray(ctx, leny)
        ta.random_fill(ctx, xv, bits, 6); ta.random_fill(ctx, yv, bits, 7)
        for cfg in CFGS:
            ctx.set_vec_config(cfg)
            ms = timeit(lambda: pkg.mp_gemv(ctx, trans, m, n, al, A, m, xv, 1, be, yv, 1, None, None, stream), reps=3)
            st, _ = ctx.last_stage_ms()
            out[name] = {"cfg": cfg, "op": name, "m": m, "n": n, "ms": ms, "stage_ms": st, "GBs_alg": m * n * rs / ms / 1e6, "GBs_kernel_touched": m * n * (rs - 16) / st[1] / 1e6,
                         "fallback": ctx.last_fallback_count()}
            print(json.dumps(out[name]), flush=True)

main()
