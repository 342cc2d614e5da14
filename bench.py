#!/usr/bin/env python
"""bench.py -- throughput of the mp_gemm hot path (BASELINE.json metric) on 1..8 B200.

    python bench.py --gpus N --steps K --warmup W            our CUDA path
    python bench.py --impl reference --gpus N --steps K ...  the reference's CPU path (oracle/_ref)

One step = one C = alpha*A*B + beta*C over synthetic uniform(-1,1) matrices with p/4-bit significands
(the reference's own benchmark convention, tests/blas/performance/test_gemm_performance.cu:65-69).
Default workload: m = n = k = 4096 with the 32-moduli / 424-bit set (BASELINE config 3).  On N > 1 GPUs
A and C are split into N row blocks (one process per GPU), B is broadcast from rank 0 over NCCL inside
every timed step; total work is fixed ("strong" scaling).

Prints ONE JSON line (rank 0).  metric = MP-GFLOP/s = 2*m*n*k / seconds / 1e9 (one mp-flop = one
multiple-precision add or mul, tests/arith/peak/test_mp_arith_peak.cuh:86-87).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (m, n, k, moduli)
    "gemm4096_424bit": (4096, 4096, 4096, 32),
    "gemm1024_106bit": (1024, 1024, 1024, 8),
    "gemm2048_106bit": (2048, 2048, 2048, 8),
    "gemm2048_212bit": (2048, 2048, 2048, 16),
    "gemm2048_318bit": (2048, 2048, 2048, 24),
    "gemm2048_424bit": (2048, 2048, 2048, 32),
    "gemm2048_530bit": (2048, 2048, 2048, 40),
    "gemm2048_636bit": (2048, 2048, 2048, 48),
    "gemm2048_742bit": (2048, 2048, 2048, 56),
    "gemm2048_848bit": (2048, 2048, 2048, 64),
}
# HBM-bound configurations (BASELINE config 4): name: (op, m, n, moduli); for dot m is the vector length
VEC_WORKLOADS = {
    "gemv16384_212bit": ("gemv", 16384, 16384, 16),
    "gemvt16384_212bit": ("gemv_t", 16384, 16384, 16),
    "dot16m_212bit": ("dot", 1 << 24, 0, 16),
}
FALLBACK_HBM_GBS = 6650.0        # /opt/skills/guides/B200_PROFILING.md fallback (MEASURED_PEAKS.json absent)
# measured on this pool's B200 by tools/mma_bench.cu (profiles/r01_pipe_rates.json)
R_MAC_IMAD_WIDE = 8.54e12        # residue-MAC/s through IMAD.WIDE.U32: the INT32 roofline of SURVEY 8(d)
PEAK_INT8_LEGACY_MMA = 1.139e15  # int8 op/s (2 per MAC) through mma.sync m16n8k32 (IMMA.16832)
CPU_SAMPLE = (64, 256)             # block of C timed on the host cores (full k)
# dram__bytes_read.sum + dram__bytes_write.sum of one k_small_umma_p<128,256> launch of the default workload (ncu --set full,
# profiles/r01_ncu_full_gemm4096_424bit_final.txt): 1.385 GB + 0.650 GB; algorithmic: 40 x (2 x 16.8 MB operand planes + 16.8 MB result plane) = 2.01 GB
NCU_DRAM_BYTES_SMALL_UMMA = 2.035e9
FALLBACK_BF16_TFLOPS = 1590.0    # /opt/skills/guides/B200_PROFILING.md fallback (MEASURED_PEAKS.json absent)


def read_measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d.get("bf16_tflops", FALLBACK_BF16_TFLOPS)), "measured"
        except Exception:
            pass
    return FALLBACK_BF16_TFLOPS, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                for nme, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w": statistics.median(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_leg(N, m_s, n_s, k, bits, seed=7):
    """The reference's own CPU implementation (host mp_mul/mp_add over mp_float_t, compiled unmodified into
    oracle/_ref) -- or the C port (oracle/) where no _ref binary exists -- on a bounded sample: an
    m_s x n_s block of C with the full inner dimension k, all host threads."""
    import numpy as np
    import oracle
    from oracle import gen
    orc = oracle.Oracle(N, oracle.HOST)

    def recs(count, sd):
        return orc.random_records(count, bits, sd)
    A, B, C = recs(m_s * k, seed), recs(k * n_s, seed + 1), recs(m_s * n_s, seed + 2)
    al, be = recs(1, seed + 3), recs(1, seed + 4)
    if oracle.have_ref(N):
        ref = oracle.RefLib(N)
        t0 = time.perf_counter()
        _, nt = ref.host_gemm(m_s, n_s, k, al, A, B, be, C)
        dt = time.perf_counter() - t0
        kind = "reference"
    else:
        t0 = time.perf_counter()
        orc.gemm(m_s, n_s, k, al, A, B, be, C)
        dt = time.perf_counter() - t0
        nt = os.cpu_count()
        kind = "port"
    gflops = 2.0 * m_s * n_s * k / dt / 1e9
    return {"value": gflops, "unit": "MP-GFLOP/s", "cores": int(nt), "kind": kind, "seconds": dt,
            "sample": "%dx%d block of C with full k=%d (%d-bit inputs), %s host mp_mul+mp_add, OpenMP" % (m_s, n_s, k, bits,
                      "reference" if kind == "reference" else "oracle-port")}


def read_measured_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback"


def run_vec(args):
    """mp_gemv / mp_dot (HBM-bound).  One step = one call; GEMV (N) shards by row blocks of A and y (x replicated, no collective),
    GEMV (T) and DOT shard by rows / segments and all-gather packed partials that every rank reduces in RNS."""
    import ctypes
    op, m, n, N = VEC_WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import oracle
    from oracle import constants
    precision = constants.compute(oracle.moduli_sets()[N])["mp_precision"]
    bits = precision // 4
    rs = 4 * N + 40
    config = {"workload": args.workload, "op": "mp_" + op, "m": m, "n": n, "moduli": N, "precision_bits": precision, "input_significand_bits": bits,
              "sharding": "single GPU" if world == 1 else ("row blocks x%d" % world if op == "gemv" else "segments x%d, packed partials all-gathered (NCCL) and reduced in RNS on every rank" % world),
              "l2": "operands exceed the 126 MB L2; no flush needed"}
    if args.impl == "reference":
        if rank != 0:
            return
        orc = oracle.Oracle(N, oracle.HOST)
        ns = 1 << 21                                  # bounded sample: a 2^21-element dot product on the host cores
        vals = []
        for i in range(args.warmup + args.steps):
            x, y = orc.random_records(ns, bits, 11 + 2 * i), orc.random_records(ns, bits, 12 + 2 * i)
            t0 = time.perf_counter()
            if oracle.have_ref(N):
                _, nt = oracle.RefLib(N).host_dot_omp(x, y); kind = "reference"
            else:
                orc.dot_omp(x, y); nt = os.cpu_count(); kind = "port"
            dt = time.perf_counter() - t0
            if i >= args.warmup:
                vals.append(2.0 * ns / dt / 1e9)
        v = statistics.mean(vals)
        cb = {"value": v, "unit": "MP-GFLOP/s", "cores": int(nt), "kind": kind, "sample": "mp_dot of 2^21 elements (%d-bit inputs), %s host mp_mul+mp_add, OpenMP" % (bits, kind)}
        print(json.dumps({"impl": "reference", "metric": "mp_%s MP-GFLOP/s" % op.split("_")[0], "value": v, "unit": "MP-GFLOP/s", "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": 2.0 * ns / v / 1e6, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                          "dtype": "int32 residues (host mp_float_t arithmetic)", "data": "synthetic", "config": config, "cpu_baseline": cb,
                          "e2e": {"value": v, "unit": "MP-GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    import torch
    import _pkg
    pkg = _pkg.load()
    from mpres_blas_b200 import parallel, torch_arrays as ta
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU path in this library")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = pkg.Context(N, local_rank)
    stream = torch.cuda.current_stream().cuda_stream
    lib = ctx.lib
    assert m % world == 0
    ml = m // world
    part = torch.zeros(rs, dtype=torch.uint8, device="cuda")
    gathered = torch.zeros(rs * world, dtype=torch.uint8, device="cuda")
    if op == "dot":
        x, y, r = ta.TorchMpArray(ctx, ml), ta.TorchMpArray(ctx, ml), ta.TorchMpArray(ctx, 1)
        ta.random_fill(ctx, x, bits, 100 + rank); ta.random_fill(ctx, y, bits, 200 + rank)
        flops, alg_bytes = 2.0 * m, 2.0 * m * rs

        def step():
            if world == 1:
                pkg.mp_dot(ctx, ml, x, 1, y, 1, r, None, stream)
            else:
                pkg._check(lib.mpres_dot_partial(ctx.h, ml, ctypes.byref(x.s), 1, ctypes.byref(y.s), 1, ctypes.c_void_p(part.data_ptr()), ctypes.c_void_p(stream)), "mpres_dot_partial")
                dist.all_gather_into_tensor(gathered, part)
                pkg._check(lib.mpres_reduce_partials(ctx.h, ctypes.c_void_p(gathered.data_ptr()), world, ctypes.byref(r.s), ctypes.c_void_p(stream)), "mpres_reduce_partials")
        host_arrays, out_arr = [(x, ml), (y, ml)], (r, 1)
    else:
        tr = op == "gemv_t"
        A = ta.TorchMpArray(ctx, ml * n)
        lenx, leny = (ml, n) if tr else (n, ml)
        xv, yv, y0 = ta.TorchMpArray(ctx, lenx), ta.TorchMpArray(ctx, leny), ta.TorchMpArray(ctx, leny)
        al, be = ta.TorchMpArray(ctx, 1), ta.TorchMpArray(ctx, 1)
        ta.random_fill(ctx, A, bits, 300 + rank); ta.random_fill(ctx, xv, bits, 400 + (rank if tr else 0)); ta.random_fill(ctx, y0, bits, 500 + rank)
        ta.random_fill(ctx, al, bits, 41); ta.random_fill(ctx, be, bits, 42)
        flops, alg_bytes = 2.0 * m * n, float(m) * n * rs + (n + 2.0 * m) * rs
        sharded_t = tr and world > 1
        if sharded_t:
            # y = alpha A^T x + beta y with A and x split by rows: every rank forms t_r = alpha A_r^T x_r (a full-length vector), the t_r are
            # all-gathered (the four SoA arrays) and added to round(beta y) in rank order with mp_axpy on every rank (GEMV (T))
            one = ta.TorchMpArray(ctx, 1)
            pkg._check(lib.mpres_array_set_binary(ctx.h, ctypes.byref(one.s), ctypes.c_size_t(0), ctypes.c_void_p(torch.zeros(1, dtype=torch.int32, device="cuda").data_ptr()),
                                                  ctypes.c_void_p(torch.zeros(1, dtype=torch.int32, device="cuda").data_ptr()),
                                                  ctypes.c_void_p(torch.ones(1, dtype=torch.int32, device="cuda").data_ptr()), 1, ctypes.c_size_t(1), None), "mpres_array_set_binary")
            torch.cuda.synchronize()
            tpart, tzero = ta.TorchMpArray(ctx, n), ta.TorchMpArray(ctx, n)
            tall = [ta.TorchMpArray(ctx, n) for _ in range(world)]
            config["sharding"] = "row blocks x%d of A and x, partial y all-gathered (NCCL) and summed with mp_axpy in rank order on every rank" % world

        def step():
            for dst, src in zip(yv.tensors(), y0.tensors()):
                dst.copy_(src, non_blocking=True)
            if not sharded_t:
                pkg.mp_gemv(ctx, pkg.mblas_trans if tr else pkg.mblas_no_trans, ml, n, al, A, ml, xv, 1, be, yv, 1, None, None, stream)
                return
            for dst, src in zip(tpart.tensors(), tzero.tensors()):
                dst.copy_(src, non_blocking=True)
            pkg.mp_gemv(ctx, pkg.mblas_trans, ml, n, al, A, ml, xv, 1, be, tpart, 1, None, None, stream)     # beta * 0 = 0
            for f in range(4):
                dist.all_gather([t.tensors()[f] for t in tall], tpart.tensors()[f])
            pkg.mp_scal(ctx, n, be, yv, 1, stream)
            for t in tall:
                pkg.mp_axpy(ctx, n, one, t, 1, yv, 1, None, stream)
        host_arrays, out_arr = [(A, ml * n), (xv, lenx), (y0, leny)], (yv, leny)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(); torch.cuda.synchronize()
    for _ in range(args.warmup):
        step()
    barrier()
    ctx.set_profiling(True)
    sampler = ClockSampler(local_rank); sampler.start()
    launches0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    launches = ctx.launch_count - launches0
    fallback = ctx.last_fallback_count()
    try:
        stage_ms, _ = ctx.last_stage_ms()
    except Exception:
        stage_ms = None
    ctx.set_profiling(False)
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = flops / (ms_step * 1e-3) / 1e9
    e2e = None
    host_bytes = sum(c for _, c in host_arrays) * rs
    if not args.no_e2e and host_bytes <= 8e9:
        bufs = [torch.empty(c * rs, dtype=torch.uint8).pin_memory() for _, c in host_arrays]
        hout = torch.empty(out_arr[1] * rs, dtype=torch.uint8).pin_memory()
        for (arr, c), b in zip(host_arrays, bufs):
            arr.device2host_ptr(b.data_ptr(), c)

        def e2e_step():
            for (arr, c), b in zip(host_arrays, bufs):
                arr.host2device_ptr(b.data_ptr(), c)
            step()
            out_arr[0].device2host_ptr(hout.data_ptr(), out_arr[1])
        e2e_step(); barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        barrier()
        dt = (time.perf_counter() - t0) / args.e2e_steps
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); dt = float(t.item())
        e2e = {"value": flops / dt / 1e9, "unit": "MP-GFLOP/s", "h2d_bytes_per_step": int(host_bytes * world), "d2h_bytes_per_step": int(out_arr[1] * rs * world),
               "ms_per_step": dt * 1e3, "steps": args.e2e_steps, "path": "mpres_array_host2device(operands) + the call + mpres_array_device2host(result), pinned host AoS mp_float_t[]"}
    elif not args.no_e2e:
        e2e = {"value": None, "unit": "MP-GFLOP/s", "skipped": "host copy of the operands is %.1f GB per step (PCIe-bound by construction); run with a smaller workload" % (host_bytes / 1e9)}
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    hbm, src = read_measured_hbm()
    roof = None
    if stage_ms:
        tk = stage_ms[1] * 1e-3        # the single-pass accumulation kernel of this rank
        ach = alg_bytes / world / tk / 1e9
        roof = {"bound": "hbm", "kernel": "k_mv_acc_n" if op == "gemv" else "k_mv_acc_t", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
                "peak_source": "%s copy bandwidth" % src, "traffic": None, "avg_launch_ms": stage_ms[1], "algorithmic_bytes_per_launch": alg_bytes / world,
                "stage_ms": {"scale_vectors": stage_ms[0], "accumulate": stage_ms[1], "finalize": stage_ms[2]},
                "whole_call_frac": alg_bytes / world / (ms_step * 1e-3) / 1e9 / hbm,
                # the kernel never reads the lower interval bounds (16 of the 4N+40 bytes per element): the same time against the bytes it must touch
                "frac_of_touched_bytes": ach / hbm * (rs - 16) / rs,
                "note": "achieved counts the algorithmic bytes of SURVEY 8(d), (4N+40) per element; the kernel touches (4N+24), so frac can exceed 1"}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        orc = oracle.Oracle(N, oracle.HOST)
        ns = 1 << 21
        xh, yh = orc.random_records(ns, bits, 11), orc.random_records(ns, bits, 12)
        t0 = time.perf_counter()
        if oracle.have_ref(N):
            _, nt = oracle.RefLib(N).host_dot_omp(xh, yh); kind = "reference"
        else:
            orc.dot_omp(xh, yh); nt = os.cpu_count(); kind = "port"
        dt = time.perf_counter() - t0
        cpu = {"value": 2.0 * ns / dt / 1e9, "unit": "MP-GFLOP/s", "cores": int(nt), "kind": kind, "seconds": dt,
               "sample": "mp_dot of 2^21 elements (%d-bit inputs), %s host mp_mul+mp_add, OpenMP" % (bits, kind)}
    print(json.dumps({"metric": "mp_%s MP-GFLOP/s" % op.split("_")[0], "value": value, "unit": "MP-GFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                      "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                      "dtype": "int32 RNS residues, u64 lazy accumulation (f64 interval bounds)", "data": "synthetic", "config": config,
                      "gpu_launches": int(launches), "fallback_elements_last_step": int(fallback), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "e2e": e2e}))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="gemm4096_424bit", choices=sorted(WORKLOADS) + sorted(VEC_WORKLOADS))
    ap.add_argument("--mode", default="auto", choices=["auto", "reference_order", "fast"])
    ap.add_argument("--stage2", default="small", choices=["small", "small_t128", "small_k64", "small_tiled", "umma", "umma_unstacked", "mma_sync"], help="stage-2 kernel (A/B measurement)")
    ap.add_argument("--stage3", type=int, default=0, choices=[0, 1, 2, 3, 4], help="stage-3 kernel variant (mpres_set_stage3_kernel; A/B measurement)")
    ap.add_argument("--bcast", default="lean", choices=["lean", "full"], help="N > 1: what the per-step broadcast of B moves (lean: the fields the small-base path reads, verified on the device; full: all four SoA arrays)")
    ap.add_argument("--prefetch", action="store_true", help="N > 1, lean broadcast: issue the next step's broadcast of B on a second stream behind the current multiply "
                    "(measured: no gain on B200 -- NCCL's blocks find no room beside the multiply kernels, which fill every SM; kept for experiments)")
    ap.add_argument("--full-precision-inputs", action="store_true", help="p-bit significands instead of p/4")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--e2e-path", choices=["pipelined", "sequential"], default="pipelined",
                    help="GEMM end-to-end leg: mpres_gemm_host (one call, transfers overlapped) or the reference caller's call-by-call sequence")
    ap.add_argument("--e2e-panels", type=int, default=0, help="column panels of mpres_gemm_host (0 = the library's choice)")
    args = ap.parse_args()
    if args.impl == "reference":
        # torchrun pins OMP_NUM_THREADS to 1; the CPU arm is meant to use every host core (set before any OpenMP runtime loads)
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"     # keep NCCL's version banner off stdout: the bench prints ONE JSON line
    if args.workload in VEC_WORKLOADS:
        return run_vec(args)

    m, n, k, N = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import oracle  # only for the precision table and the CPU legs (never on the measured GPU path)
    from oracle import constants
    precision = constants.compute(oracle.moduli_sets()[N])["mp_precision"]
    bits = precision if args.full_precision_inputs else precision // 4
    config = {"workload": args.workload, "op": "mp_gemm", "m": m, "n": n, "k": k, "moduli": N, "precision_bits": precision,
              "input_significand_bits": bits, "sharding": "A,C row blocks x%d, B broadcast (NCCL) inside the step" % world if world > 1 else "single GPU",
              "l2": "operands (%.2f GB per matrix) exceed the 126 MB L2; no flush needed" % (m * k * (4 * N + 40) / 1e9), "mode": args.mode,
              "step": "C <- C0 (device copy of the pristine p/4-bit C), [B broadcast], C = alpha*A*B + beta*C"}

    if args.impl == "reference":
        if rank != 0:
            return
        ms_s, ns_s = CPU_SAMPLE
        vals, last = [], None
        for i in range(args.warmup + args.steps):
            last = cpu_reference_leg(N, ms_s, ns_s, k, bits, seed=7 + i)
            if i >= args.warmup:
                vals.append(last)
        v = statistics.mean(x["value"] for x in vals) if vals else last["value"]
        secs = statistics.mean(x["seconds"] for x in vals) if vals else last["seconds"]
        cb = dict(last); cb["value"] = v
        print(json.dumps({"impl": "reference", "metric": "mp_gemm MP-GFLOP/s", "value": v, "unit": "MP-GFLOP/s", "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs * 1e3, "higher_is_better": True,
                          "scaling": "strong", "vs_baseline": None, "dtype": "int32 residues (host mp_float_t arithmetic)",
                          "data": "synthetic", "config": config, "cpu_baseline": cb,
                          "e2e": {"value": v, "unit": "MP-GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    import _pkg
    pkg = _pkg.load()
    from mpres_blas_b200 import parallel, torch_arrays as ta
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU path in this library")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = pkg.Context(N, local_rank)
    ctx.set_mode({"auto": pkg.MODE_AUTO, "reference_order": pkg.MODE_REFERENCE_ORDER, "fast": pkg.MODE_FAST}[args.mode])
    ctx.set_stage2_kernel({"small": pkg.STAGE2_SMALL, "small_tiled": pkg.STAGE2_SMALL_TILED, "small_k64": pkg.STAGE2_SMALL_K64, "small_t128": pkg.STAGE2_SMALL_T128, "umma": pkg.STAGE2_UMMA, "umma_unstacked": pkg.STAGE2_UMMA_UNSTACKED, "mma_sync": pkg.STAGE2_MMA_SYNC}[args.stage2])
    config["stage2_kernel"] = args.stage2
    if world > 1:
        config["broadcast"] = args.bcast
    ctx.set_stage3_kernel(args.stage3)
    config["stage3_kernel"] = args.stage3
    assert m % world == 0
    mr = m // world                       # rows of A and C owned by this rank
    A = ta.TorchMpArray(ctx, mr * k)
    B = ta.TorchMpArray(ctx, k * n)
    C = ta.TorchMpArray(ctx, mr * n)
    C0 = ta.TorchMpArray(ctx, mr * n)     # pristine p/4-bit C, copied into C at the start of every step
    alpha, beta = ta.TorchMpArray(ctx, 1), ta.TorchMpArray(ctx, 1)
    ta.random_fill(ctx, A, bits, 1000 + rank)
    ta.random_fill(ctx, C0, bits, 2000 + rank)
    ta.random_fill(ctx, alpha, bits, 31)
    ta.random_fill(ctx, beta, bits, 32)
    if rank == 0:
        ta.random_fill(ctx, B, bits, 33)
    stream = torch.cuda.current_stream().cuda_stream

    # mp_gemm updates C in place, so every step gets its own pristine p/4-bit C: a ring of pre-filled copies (HBM has the room:
    # 2.8 GB each at config 3); only when the run has more steps than buffers is a buffer restored (device copy) before reuse
    free_b, _ = torch.cuda.mem_get_info()
    c_bytes = C0.nbytes()
    n_buf = int(max(1, min(args.steps + args.warmup, 24, (free_b - (16 << 30)) // max(1, c_bytes))))
    ring = [C] + [ta.TorchMpArray(ctx, mr * n) for _ in range(n_buf - 1)]
    for Cb in ring:
        for dst, src in zip(Cb.tensors(), C0.tensors()):
            dst.copy_(src)
    state = {"i": 0}
    config["step"] = "[B broadcast], C_i = alpha*A*B + beta*C_i on a pristine p/4-bit C_i (ring of %d pre-filled device buffers)" % n_buf

    lean = parallel.LeanBroadcast(dist, N) if (world > 1 and args.bcast == "lean") else None
    lean_state = {"on": False, "repeats": 0}
    # Prefetch (--prefetch): two B buffers; the broadcast of step i+1 runs on a second stream while step i multiplies,
    # so a step costs max(broadcast, multiply) instead of their sum.  Every timed step still broadcasts B inside the timed region; the first
    # timed step is not prefetched from the warm-up.
    overlap = lean is not None and args.prefetch
    Bbuf = [B]
    if overlap:
        B1 = ta.TorchMpArray(ctx, k * n)
        for dst, src in zip(B1.tensors(), B.tensors()):
            dst.copy_(src)                                   # on rank 0 both buffers hold B (the source of every broadcast)
        Bbuf.append(B1)
    side = torch.cuda.Stream(priority=-1) if overlap else None      # high priority: its blocks are placed as soon as an SM has room
    ev_b, ev_g = {}, {}
    total_steps = args.warmup + args.steps
    config["broadcast_prefetch"] = bool(overlap)

    def issue_bcast(i):
        Bi = Bbuf[i % len(Bbuf)]
        with torch.cuda.stream(side):
            if (i - 2) in ev_g:
                side.wait_event(ev_g.pop(i - 2))             # the multiply that read this buffer two steps ago has finished
            lean.broadcast(Bi.digits, Bi.sign, Bi.exp, Bi.eval)
            e = torch.cuda.Event(); e.record(side)
            ev_b[i] = e

    def step():
        i = state["i"]; state["i"] = i + 1
        Cb = ring[i % n_buf]
        if i >= n_buf:                                    # buffer reuse: restore the pristine C first (device-to-device)
            for dst, src in zip(Cb.tensors(), C0.tensors()):
                dst.copy_(src, non_blocking=True)
        Bi = Bbuf[i % len(Bbuf)] if (overlap and lean_state["on"]) else B
        gemm = lambda: pkg.mp_gemm(ctx, pkg.mblas_no_trans, pkg.mblas_no_trans, mr, n, k, alpha, A, mr, Bi, k, beta, Cb, mr, None, stream)
        if lean is not None and lean_state["on"]:
            # only the fields of B the small-base fast path reads travel; a device-side check guards it (parallel.LeanBroadcast)
            if overlap:
                main = torch.cuda.current_stream()
                if i not in ev_b:
                    issue_bcast(i)
                if i + 1 < total_steps and i + 1 != args.warmup:
                    issue_bcast(i + 1)
                main.wait_event(ev_b.pop(i))
                gemm()
                e = torch.cuda.Event(); e.record(main); ev_g[i] = e
            else:
                lean.broadcast(B.digits, B.sign, B.exp, B.eval)
                gemm()
            P_used, nin_used = ctx.last_small_base()
            ok = torch.tensor([1 if lean.verify(P_used, nin_used, ctx.last_fallback_count()) else 0], dtype=torch.int32, device="cuda")
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 1:
                return
            lean_state["repeats"] += 1                   # the lean copy was not enough somewhere: repeat with the complete B
            for dst, src in zip(Cb.tensors(), C0.tensors()):
                dst.copy_(src, non_blocking=True)
            gemm = lambda: pkg.mp_gemm(ctx, pkg.mblas_no_trans, pkg.mblas_no_trans, mr, n, k, alpha, A, mr, B, k, beta, Cb, mr, None, stream)
        parallel.gemm_row_sharded(dist, B.tensors(), gemm)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for w in range(args.warmup):
        step()
        if lean is not None and w == 0:
            # the first (fully replicated) call tells how many residues per entry the input conversion reads on this rank
            P_used, nin_used = ctx.last_small_base()
            nin_all = lean.agree(nin_used if P_used > 0 else 0, "cuda")
            lean_state["on"] = nin_all > 0
    barrier()
    ctx.set_profiling(True)
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    use_profiler_range = os.environ.get("MPRES_BENCH_PROFILER_RANGE") == "1"   # ncu --profile-from-start off
    if use_profiler_range:
        torch.cuda.profiler.start()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    if use_profiler_range:
        torch.cuda.profiler.stop()
    clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    launches = ctx.launch_count - launches0
    fallback = ctx.last_fallback_count()
    slow_listed = ctx.last_slow_count()
    base_size = ctx.last_base_size()          # moduli stages 1-2 actually ran on (reduced-base fast path)
    small_P, small_nin = ctx.last_small_base()   # one-byte moduli of the small-modulus stage 2 (0: not used)
    try:
        stage_ms, s2_launches = ctx.last_stage_ms()
    except Exception:
        stage_ms, s2_launches = None, 0
    ctx.set_profiling(False)
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = 2.0 * m * n * k / (ms_step * 1e-3) / 1e9

    # ---- end-to-end through the C-ABI with host buffers ------------------------------------------------
    e2e = None
    if not args.no_e2e:
        rs = 4 * N + 40
        hA = torch.empty(mr * k * rs, dtype=torch.uint8).pin_memory()
        hB = torch.empty(k * n * rs, dtype=torch.uint8).pin_memory()
        hC = torch.empty(mr * n * rs, dtype=torch.uint8).pin_memory()
        hOut = torch.empty(mr * n * rs, dtype=torch.uint8).pin_memory()
        hal, hbe = torch.empty(rs, dtype=torch.uint8).pin_memory(), torch.empty(rs, dtype=torch.uint8).pin_memory()
        if world > 1:
            for t in B.tensors():
                dist.broadcast(t, src=0)
        A.device2host_ptr(hA.data_ptr(), mr * k); B.device2host_ptr(hB.data_ptr(), k * n)
        C0.device2host_ptr(hC.data_ptr(), mr * n)
        alpha.device2host_ptr(hal.data_ptr(), 1); beta.device2host_ptr(hbe.data_ptr(), 1)

        def e2e_step():
            if args.e2e_path == "pipelined" and world > 1 and n % world == 0:
                # every rank pulls 1/N of B through its own PCIe link, the ranks gather the rest over NVLink (five contiguous ranges of the
                # SoA arrays), then one call over the host A / C row blocks with B resident (mpres_gemm_host_bdev)
                cols = n // world
                B.host2device_ptr_at(k * cols * rank, hB.data_ptr() + k * cols * rank * rs, k * cols)
                parallel.gather_column_shards(dist, B.slices(0, k * n), B.slices(k * cols * rank, k * cols))
                torch.cuda.synchronize()
                pkg.mp_gemm_host_bdev(ctx, pkg.mblas_no_trans, pkg.mblas_no_trans, mr, n, k, hal, hA, mr, B, k, hbe, hC, mr, out=hOut, panels=args.e2e_panels)
                return
            if args.e2e_path == "pipelined":
                # one call over the host buffers: uploads, compute and download overlapped by column panels (mpres_gemm_host)
                pkg.mp_gemm_host(ctx, pkg.mblas_no_trans, pkg.mblas_no_trans, mr, n, k, hal, hA, mr, hB, k, hbe, hC, mr, out=hOut, panels=args.e2e_panels)
                return
            # the reference caller's sequence, call by call (tests/blas/test_gemm.cu)
            A.host2device_ptr(hA.data_ptr(), mr * k)
            B.host2device_ptr(hB.data_ptr(), k * n)
            C.host2device_ptr(hC.data_ptr(), mr * n)
            alpha.host2device_ptr(hal.data_ptr(), 1); beta.host2device_ptr(hbe.data_ptr(), 1)
            pkg.mp_gemm(ctx, pkg.mblas_no_trans, pkg.mblas_no_trans, mr, n, k, alpha, A, mr, B, k, beta, C, mr, None, stream)
            C.device2host_ptr(hOut.data_ptr(), mr * n)
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        barrier()
        dt = (time.perf_counter() - t0) / args.e2e_steps
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        b_up = k * n // world if (args.e2e_path == "pipelined" and world > 1 and n % world == 0) else k * n
        h2d = (mr * k + b_up + mr * n + 2) * rs
        d2h = mr * n * rs
        e2e = {"value": 2.0 * m * n * k / dt / 1e9, "unit": "MP-GFLOP/s", "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
               "ms_per_step": dt * 1e3, "steps": args.e2e_steps,
               "path": ("per rank: mpres_array_host2device(1/N of B) + all-gather of B over NVLink + mpres_gemm_host_bdev(alpha, A_r, B, beta, C_r -> out_r), pinned host AoS mp_float_t[]"
                        if (args.e2e_path == "pipelined" and world > 1 and n % world == 0) else
                        "mpres_gemm_host(alpha, A, B, beta, C -> out): pinned host AoS mp_float_t[], PCIe transfers pipelined with the compute by column panels"
                        if args.e2e_path == "pipelined" else
                        "mpres_array_host2device(A,B,C,alpha,beta) + mpres_gemm + mpres_array_device2host(C), pinned host AoS mp_float_t[]")}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (stage 2: k_limb_umma, tcgen05.mma kind::i8) ---------------------
    bf16_peak, peak_src = read_measured_peaks()
    roof, int32_roof = None, None
    if stage_ms and s2_launches:
        t2 = stage_ms[1] * 1e-3                       # all stage-2 launches of one mp_gemm call on this rank
        nb = base_size if 0 < base_size <= N else N
        if small_P > 0:
            limb_macs = 1.0 * mr * n * k * small_P    # one u8 x u8 MAC per one-byte modulus
        else:
            limb_macs = 16.0 * mr * n * k * nb        # int8 MACs the kernel executes: 16 limb products per residue MAC, nb moduli
        ops = 2.0 * limb_macs
        achieved = ops / t2 / 1e12
        peak = 2.0 * bf16_peak                        # dense int8 = 2 x dense bf16 on the same tensor cores
        kname = {"small": "k_small_umma_p (persistent, tcgen05.mma kind::i8 per one-byte modulus, TMA ring, two TMEM accumulators)" if small_P > 0 else "k_limb_umma<stacked> (small base not selected)",
                 "small_tiled": "k_small_umma (one tile per CTA)" if small_P > 0 else "k_limb_umma<stacked> (small base not selected)",
                 "small_t128": "k_small_umma_p<128,128> (persistent, 128 x 256 tiles, double-buffered accumulator)" if small_P > 0 else "k_limb_umma<stacked> (small base not selected)",
                 "small_k64": "k_small_umma_p<64> (persistent, 64-byte operand rows)" if small_P > 0 else "k_limb_umma<stacked> (small base not selected)",
                 "umma": "k_limb_umma<stacked> (tcgen05.mma kind::i8, TMA, TMEM)", "umma_unstacked": "k_limb_umma<unstacked>",
                 "mma_sync": "k_limb_gemm<0>+<1> (legacy mma.sync IMMA)"}[args.stage2]
        roof = {"bound": "tensor", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "TOP/s (int8)",
                "frac": achieved / peak, "peak_source": "2 x bf16 %s peak of %s TFLOP/s (int8 dense = 2 x bf16 dense)" % (peak_src, bf16_peak),
                "traffic": (NCU_DRAM_BYTES_SMALL_UMMA if (small_P == 40 and args.workload == "gemm4096_424bit" and world == 1 and args.stage2 == "small") else None),
                "launches_per_step": s2_launches, "avg_launch_ms": stage_ms[1] / s2_launches,
                "algorithmic_ops_per_launch": ops / s2_launches, "moduli_in_stage2": small_P if small_P > 0 else nb, "moduli_total": N,
                "small_base": {"one_byte_moduli": small_P, "input_residues_read": small_nin},
                "frac_of_nominal_int8_peak_4500": achieved / 4500.0,
                "stage_ms": {"stage1_align": stage_ms[0], "stage2_limb_gemm": stage_ms[1], "stage3_extend_normalise_epilogue": stage_ms[2]}}
        int32_roof = {"definition": "SURVEY 8(d): m*n*k*N residue-MACs / t / R_mac, R_mac = measured IMAD.WIDE.U32 rate (all N moduli counted: "
                                    "the reduced base is an algorithmic saving)",
                      "R_mac_per_s": R_MAC_IMAD_WIDE, "residue_macs_per_s_stage2": mr * n * k * N / t2,
                      "frac_stage2": mr * n * k * N / t2 / R_MAC_IMAD_WIDE,
                      "frac_whole_step": (m / world) * n * k * N / (ms_step * 1e-3) / R_MAC_IMAD_WIDE}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference_leg(N, CPU_SAMPLE[0], CPU_SAMPLE[1], k, bits)
    line = {"metric": "mp_gemm MP-GFLOP/s", "value": value, "unit": "MP-GFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u8 limbs of int32 RNS residues, s32 accumulate (f64 interval bounds)", "data": "synthetic", "config": config,
            "gpu_launches": int(launches), "fallback_elements_last_step": int(fallback), "stage3_listed_elements_last_step": int(slow_listed), "reduced_base_moduli": int(base_size), "small_base_moduli": int(small_P), "clocks": clocks,
            "broadcast_bytes_per_step": (int(lean.nbytes(k * n)) if (lean is not None and lean_state["on"]) else (B.nbytes() if world > 1 else 0)),
            "lean_broadcast_repeats": lean_state["repeats"],
            "roofline": roof, "int32_roofline": int32_roof, "cpu_baseline": cpu, "e2e": e2e}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
