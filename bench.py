#!/usr/bin/env python
"""bench.py -- throughput of the mp_gemm hot path (BASELINE.json metric) on 1..8 B200.

    python bench.py --gpus N --steps K --warmup W            our CUDA path
    python bench.py --impl reference --gpus N --steps K ...  the reference's CPU path (oracle/_ref)

One step = one C = alpha*A*B + beta*C over synthetic uniform(-1,1) matrices with p/4-bit significands
(the reference's own benchmark convention, tests/blas/performance/test_gemm_performance.cu:65-69).
Default workload: m = n = k = 4096 with the 32-moduli / 424-bit set (BASELINE config 3).  On N > 1 GPUs
A and C are split into N row blocks (one process per GPU), every rank holds B; inside every timed step each
rank converts ONE column block of B and the one-byte planes are exchanged over NVLink (mpres_gemm_sharded);
total work is fixed ("strong" scaling).

Prints ONE JSON line (rank 0).  metric = MP-GFLOP/s = 2*m*n*k / seconds / 1e9 (one mp-flop = one
multiple-precision add or mul, tests/arith/peak/test_mp_arith_peak.cuh:86-87).  Without --workload the line also
carries `sub_results`: the other BASELINE configurations (config 2: 1024^3 at 106 bit with p/4- and p-bit inputs,
config 4: GEMV 16384^2 / DOT 2^24 at 212 bit, config 5: the 2048^3 precision sweep) measured the same way.
"""
import argparse
import ctypes
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
DEFAULT_WORKLOAD = "gemm4096_424bit"
FALLBACK_HBM_GBS = 6650.0        # /opt/skills/guides/B200_PROFILING.md fallback (MEASURED_PEAKS.json absent)
FALLBACK_BF16_TFLOPS = 1590.0    # /opt/skills/guides/B200_PROFILING.md fallback (MEASURED_PEAKS.json absent)
# measured on this pool's B200 by tools/mma_bench.cu (profiles/r01_pipe_rates.jsonl)
R_MAC_IMAD_WIDE = 8.54e12        # residue-MAC/s through IMAD.WIDE.U32: the INT32 roofline of SURVEY 8(d)
CPU_SAMPLE = (64, 256)           # block of C timed on the host cores (full k)
VERIFY_ROWS, VERIFY_COLS = 64, 32


def read_peaks():
    """(hbm GB/s, bf16 TFLOP/s burst, source) from the driver-written MEASURED_PEAKS.json, else the recipe's fallbacks"""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d.get("hbm_gbs", FALLBACK_HBM_GBS)), float(d.get("bf16_tflops", FALLBACK_BF16_TFLOPS)), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, FALLBACK_BF16_TFLOPS, "fallback (B200_PROFILING.md)"


def read_int8_peak(bf16_peak):
    """dense int8 tensor peak in TOP/s: measured (tools/int8_peak.cu: cuBLAS s8 x s8 -> s32 8192^3, profiles/r02_int8_peak.json) when that
    file exists, else 2 x the measured bf16 burst peak (int8 dense = 2 x bf16 dense on the same tensor cores)"""
    p = os.path.join(ROOT, "profiles", "r02_int8_peak.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["int8_tops"]), "measured: %s" % d.get("how", "profiles/r02_int8_peak.json")
        except Exception:
            pass
    return 2.0 * bf16_peak, "derived: 2 x bf16 burst peak %.1f TFLOP/s" % bf16_peak


def read_ncu_traffic(workload, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, parsed from the committed ncu --set full capture of this
    workload (profiles/r02_ncu_traffic.json, written by tools/ncu_summary.py from the raw CSV of the same bench command)"""
    p = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
    if not os.path.exists(p):
        return None
    try:
        d = json.load(open(p))
        if d.get("workload") != workload:
            return None
        for name, v in d.get("kernels", {}).items():
            if name.split("(")[0].split("<")[0] == kernel.split("(")[0].split("<")[0]:
                return float(v["dram_bytes_read"]) + float(v["dram_bytes_write"])
    except Exception:
        pass
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs.  One query process per rank streams lines for the whole
    run (its first line takes ~0.1 s, longer than a short timed region) and every line is stamped on arrival; `window(t0, t1)` summarises
    the samples taken between two host times.  bench.py brackets the timed region with it; when that region is shorter than the sampling
    period and holds no sample, the same loop is repeated UNTIMED for a few periods and its samples are reported (`window` says so)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index, period_ms=25):
        self.idx, self.rows, self.proc, self.period = gpu_index, [], None, period_ms

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", str(self.period)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def close(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
            self.proc = None

    def window(self, t0, t1, what="timed region"):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.002 * self.period)            # a line arriving just after t1 was sampled inside the window
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, r in list(self.rows):
            if not (t0 <= t <= t1 + 0.001 * self.period):
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                for nme, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w": statistics.median(pw) if pw else None, "samples": len(sm), "window": what, "reasons": sorted(reasons)}


def clocks_under_load(env, t0, t1, step, ms_step, steps):
    """clocks of the timed region [t0, t1]; if it held fewer than three samples, of an untimed repeat of the same loop (the same number of steps on every rank)"""
    sampler = env.sampler
    c = sampler.window(t0, t1)
    none = env.max_over_ranks(0.0 if c.get("samples", 0) >= 3 else 1.0) > 0   # any rank with fewer than three samples: all repeat (the sharded step is collective)
    if none and sampler.proc:
        reps = max(steps, int(12 * sampler.period / max(ms_step, 1e-3)) + 1)
        env.barrier()
        r0 = time.time()
        for _ in range(reps):
            step()
        env.barrier()
        c = sampler.window(r0, time.time(), "untimed repeat of the timed loop (%d steps, same load): the timed region (%.1f ms) is shorter than the %d ms sampling period"
                           % (reps, ms_step * steps, sampler.period))
    return c


# ---- CPU legs (the reference's own host code through oracle/_ref, or the C port where no _ref binary exists) -------------------------

def cpu_reference_leg(N, m_s, n_s, k, bits, seed=7, mpfr=False):
    """The reference's own CPU implementation (host mp_mul/mp_add over mp_float_t, compiled unmodified into oracle/_ref) -- or the C
    port (oracle/) where no _ref binary exists -- on a bounded sample: an m_s x n_s block of C with the full inner dimension k, all
    host threads.  mpfr: also the reference's MPFR loop at the working precision (tests/blas/v2/gemm/test_mpfr_gemm.cuh:29-60)."""
    import oracle
    orc = oracle.Oracle(N, oracle.HOST)

    def recs(count, sd):
        return orc.random_records(count, bits, sd)
    A, B, C = recs(m_s * k, seed), recs(k * n_s, seed + 1), recs(m_s * n_s, seed + 2)
    al, be = recs(1, seed + 3), recs(1, seed + 4)
    extra = {}
    if oracle.have_ref(N):
        ref = oracle.RefLib(N)
        t0 = time.perf_counter()
        _, nt = ref.host_gemm(m_s, n_s, k, al, A, B, be, C)
        dt = time.perf_counter() - t0
        kind = "reference"
        if mpfr:
            try:
                secs, nt2 = ref.mpfr_gemm_timed(m_s, n_s, k, al, A, B, be, C, orc.precision)
                extra["mpfr"] = {"value": 2.0 * m_s * n_s * k / secs / 1e9, "unit": "MP-GFLOP/s", "cores": int(nt2), "precision_bits": orc.precision, "seconds": secs,
                                 "what": "the reference's MPFR GEMM loop (tests/blas/v2/gemm/test_mpfr_gemm.cuh:29-60) on the same block, operands pre-converted"}
            except Exception as e:      # an older _ref binary without the wrapper
                extra["mpfr"] = {"unavailable": repr(e)}
    else:
        t0 = time.perf_counter()
        orc.gemm(m_s, n_s, k, al, A, B, be, C)
        dt = time.perf_counter() - t0
        nt = os.cpu_count()
        kind = "port"
    gflops = 2.0 * m_s * n_s * k / dt / 1e9
    out = {"value": gflops, "unit": "MP-GFLOP/s", "cores": int(nt), "kind": kind, "seconds": dt,
           "sample": "%dx%d block of C with full k=%d (%d-bit inputs), %s host mp_mul+mp_add, OpenMP" % (m_s, n_s, k, bits,
                     "reference" if kind == "reference" else "oracle-port")}
    out.update(extra)
    return out


def cpu_dot_leg(N, bits, ns=1 << 21, seed=11):
    import oracle
    orc = oracle.Oracle(N, oracle.HOST)
    xh, yh = orc.random_records(ns, bits, seed), orc.random_records(ns, bits, seed + 1)
    t0 = time.perf_counter()
    if oracle.have_ref(N):
        _, nt = oracle.RefLib(N).host_dot_omp(xh, yh); kind = "reference"
    else:
        _, nt = orc.dot_omp(xh, yh); kind = "port"
    dt = time.perf_counter() - t0
    return {"value": 2.0 * ns / dt / 1e9, "unit": "MP-GFLOP/s", "cores": int(nt), "kind": kind, "seconds": dt,
            "sample": "mp_dot of 2^21 elements (%d-bit inputs), %s host mp_mul+mp_add, OpenMP" % (bits, kind)}


# ---- helpers over torch-owned mp_array_t views --------------------------------------------------------------------------------------

def _sub_matrix(ta, ctx, src, rows_total, cols_total, rows, cols):
    """compact copy (len(rows) x len(cols), column-major) of entries (rows, cols) of a column-major rows_total x cols_total TorchMpArray"""
    import torch
    N = ctx.N
    dst = ta.TorchMpArray(ctx, len(rows) * len(cols))
    r = torch.as_tensor(rows, device=src.digits.device, dtype=torch.long)
    c = torch.as_tensor(cols, device=src.digits.device, dtype=torch.long)
    idx = (c[:, None] * rows_total + r[None, :]).reshape(-1)          # column-major positions in src
    ln = max(1, src.size)
    dst.digits.copy_(src.digits.view(ln, N)[idx].reshape(-1))
    dst.sign.copy_(src.sign[idx]); dst.exp.copy_(src.exp[idx])
    ev = src.eval.view(2, ln, 2)
    dst.eval.copy_(ev[:, idx, :].reshape(-1))
    return dst


def _fractions(ctx, recs):
    """exact values of host mp_float_t records as Fractions (CRT over the moduli of the context; plain Python integers)"""
    from fractions import Fraction
    mods = [int(v) for v in ctx.constant(0, "int32", ctx.N)]
    M = 1
    for q in mods:
        M *= q
    w = [(M // q) * pow(M // q, -1, q) for q in mods]
    out = []
    for r in recs:
        x = sum(int(d) * wi for d, wi in zip(r["digits"], w)) % M
        v = Fraction(x) * Fraction(2) ** int(r["exp"])
        out.append(-v if int(r["sign"]) else v)
    return out


def _tolerance_check(ctx, k, alpha, beta, As, Bs, C0s, got, ref, nr, nc, take=8):
    """p-bit inputs (single-rounding stage 3): entries (i < take, j < take) of the sampled block against exact rational arithmetic under the
    reference's own error model |err| <= gamma_(k+3) (|alpha| sum |a||b| + |beta c|), u = 4 / sqrt(M)
    (tests/blas/accuracy/test_dot_accuracy.cu:41-72).  Returns (entries, failures, worst ratio of ours, worst ratio of the reference-order loop)."""
    from fractions import Fraction
    import math
    mods = [int(v) for v in ctx.constant(0, "int32", ctx.N)]
    M = 1
    for q in mods:
        M *= q
    u = Fraction(4, math.isqrt(M))
    gam = (k + 3) * u / (1 - (k + 3) * u)
    tr, tc = min(take, nr), min(take, nc)
    hA, hB = As.device2host().reshape(k, nr), Bs.device2host().reshape(nc, k)
    fa = [_fractions(ctx, hA[:, i]) for i in range(tr)]
    fb = [_fractions(ctx, hB[j, :]) for j in range(tc)]
    al, be = _fractions(ctx, alpha.device2host())[0], _fractions(ctx, beta.device2host())[0]
    hC0, hG, hR = C0s.device2host().reshape(nc, nr), got.device2host().reshape(nc, nr), ref.device2host().reshape(nc, nr)
    bad, worst_g, worst_r = 0, 0.0, 0.0
    for j in range(tc):
        c0 = _fractions(ctx, hC0[j, :tr]); g = _fractions(ctx, hG[j, :tr]); r = _fractions(ctx, hR[j, :tr])
        for i in range(tr):
            exact = al * sum(x * y for x, y in zip(fa[i], fb[j])) + be * c0[i]
            bound = gam * (abs(al) * sum(abs(x * y) for x, y in zip(fa[i], fb[j])) + abs(be * c0[i]))
            eg, er = abs(g[i] - exact), abs(r[i] - exact)
            if eg > bound:
                bad += 1
            if bound:
                worst_g = max(worst_g, float(eg / bound)); worst_r = max(worst_r, float(er / bound))
    return tr * tc, bad, worst_g, worst_r


def _equal_des(a, b):
    """number of entries whose digits, sign or exponent differ between two TorchMpArrays of equal size"""
    N = a.ctx.N
    bad = (a.digits.view(-1, N) != b.digits.view(-1, N)).any(dim=1) | (a.sign != b.sign) | (a.exp != b.exp)
    return int(bad.sum().item())


def kernel_roofline(name, ms, dims, hbm, int8_peak, int8_src, peak_src, workload):
    """roofline block of one kernel of the fast mp_gemm path from its algorithmic work per launch"""
    m, n, k, N, P, nin, W = dims
    rec = 4 * N + 40
    alg = {
        "k_norm_fast": ("hbm", m * n * (4 * N + 2 * rec)),
        "k_ext_norm_small": ("hbm", m * n * (P + 2 * rec)),
        "k_ext_small": ("hbm", m * n * (P + 4 * N)),
        "k_align_small(A)": ("hbm", m * k * (4 * nin + 24 + P + 2)),
        "k_align_small(B)": ("hbm", k * (n // W) * (4 * nin + 24 + P + 2)),
        "k_outer_info": ("hbm", (m * k + k * (n // W)) * 20),
        "k_small_umma_p": ("tensor", 2.0 * P * m * n * k),
    }
    if name not in alg or ms <= 0:
        return {"kernel": name, "avg_launch_ms": ms, "bound": None, "note": "no algorithmic-work model for this kernel"}
    bound, work = alg[name]
    t = ms * 1e-3
    if bound == "hbm":
        ach = work / t / 1e9
        return {"bound": "hbm", "kernel": name, "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "peak_source": peak_src + " copy bandwidth",
                "traffic": read_ncu_traffic(workload, name), "avg_launch_ms": ms, "algorithmic_bytes_per_launch": work}
    ach = work / t / 1e12
    return {"bound": "tensor", "kernel": name, "achieved": ach, "peak": int8_peak, "unit": "TOP/s (int8)", "frac": ach / int8_peak, "peak_source": int8_src,
            "traffic": read_ncu_traffic(workload, name), "avg_launch_ms": ms, "algorithmic_ops_per_launch": work, "frac_of_nominal_int8_peak_4500": ach / 4500.0}


class Env:
    """process-wide state shared by the workloads of one bench invocation"""

    def __init__(self, args):
        self.args = args
        self.rank = int(os.environ.get("RANK", "0")); self.world = int(os.environ.get("WORLD_SIZE", "1")); self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.dist = None
        self.torch = None
        self.ctxs = {}

    def start_gpu(self):
        import torch
        import _pkg
        self.torch = torch
        self.pkg = _pkg.load()
        from mpres_blas_b200 import parallel, torch_arrays
        self.parallel, self.ta = parallel, torch_arrays
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: there is no CPU path in this library")
        torch.cuda.set_device(self.local_rank)
        self.sampler = ClockSampler(self.local_rank).start()
        import atexit
        atexit.register(self.sampler.close)
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist = dist

    def ctx(self, N):
        if N not in self.ctxs:
            self.ctxs[N] = self.pkg.Context(N, self.local_rank)
        return self.ctxs[N]

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if self.dist is None:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, v):
        if self.dist is None:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def finish(self):
        if self.dist is not None:
            self.dist.destroy_process_group()


def precision_of(N):
    import oracle  # only for the precision table and the CPU legs (never on the measured GPU path)
    from oracle import constants
    return constants.compute(oracle.moduli_sets()[N])["mp_precision"]


# ---- mp_gemv / mp_dot ---------------------------------------------------------------------------------------------------------------

def run_vec(env, workload, steps, warmup, want_e2e=True, want_cpu=True, e2e_steps=2):
    """mp_gemv / mp_dot (HBM-bound).  One step = one call; GEMV (N) shards by row blocks of A and y (x replicated, no collective),
    GEMV (T) and DOT shard by rows / segments and all-gather packed partials that every rank reduces in RNS."""
    torch, pkg, ta, dist = env.torch, env.pkg, env.ta, env.dist
    rank, world = env.rank, env.world
    op, m, n, N = VEC_WORKLOADS[workload]
    precision = precision_of(N)
    bits = precision // 4
    rs = 4 * N + 40
    config = {"workload": workload, "op": "mp_" + op, "m": m, "n": n, "moduli": N, "precision_bits": precision, "input_significand_bits": bits,
              "sharding": "single GPU" if world == 1 else ("row blocks x%d" % world if op == "gemv" else "segments x%d, packed partials all-gathered (NCCL) and reduced in RNS on every rank" % world),
              "l2": "operands exceed the 126 MB L2; no flush needed"}
    ctx = env.ctx(N)
    ctx.set_mode(pkg.MODE_AUTO)
    stream = torch.cuda.current_stream().cuda_stream
    lib = ctx.lib
    assert m % world == 0
    ml = m // world
    part = torch.zeros(rs, dtype=torch.uint8, device="cuda")
    gathered = torch.zeros(rs * world, dtype=torch.uint8, device="cuda")
    verify = None
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

        def verify():
            # the whole product again in reference order (mp_mul, mp_add, rounding after each: src/blas/dot.cuh:84-107 semantics)
            if world > 1:
                return None
            r2 = ta.TorchMpArray(ctx, 1)
            ctx.set_mode(pkg.MODE_REFERENCE_ORDER)
            pkg.mp_dot(ctx, ml, x, 1, y, 1, r2, None, stream)
            ctx.set_mode(pkg.MODE_AUTO)
            torch.cuda.synchronize()
            return {"verified_entries": 1, "verified_mismatches": _equal_des(r, r2), "against": "the whole dot product in REFERENCE_ORDER mode (digits, sign, exponent)"}
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

        def verify():
            # VERIFY_ROWS outputs of this rank recomputed in reference order on the gathered rows / columns of A
            if sharded_t:
                return None
            g = torch.Generator(device="cpu"); g.manual_seed(777 + rank)
            pick = sorted(torch.randperm(leny, generator=g)[:VERIFY_ROWS].tolist())
            if tr:
                As = _sub_matrix(ta, ctx, A, ml, n, list(range(ml)), pick)      # ml x V: the picked columns
                ys, yw = _sub_matrix(ta, ctx, y0, leny, 1, pick, [0]), _sub_matrix(ta, ctx, yv, leny, 1, pick, [0])
                ctx.set_mode(pkg.MODE_REFERENCE_ORDER)
                pkg.mp_gemv(ctx, pkg.mblas_trans, ml, len(pick), al, As, ml, xv, 1, be, ys, 1, None, None, stream)
            else:
                As = _sub_matrix(ta, ctx, A, ml, n, pick, list(range(n)))       # V x n: the picked rows
                ys, yw = _sub_matrix(ta, ctx, y0, leny, 1, pick, [0]), _sub_matrix(ta, ctx, yv, leny, 1, pick, [0])
                ctx.set_mode(pkg.MODE_REFERENCE_ORDER)
                pkg.mp_gemv(ctx, pkg.mblas_no_trans, len(pick), n, al, As, len(pick), xv, 1, be, ys, 1, None, None, stream)
            ctx.set_mode(pkg.MODE_AUTO)
            torch.cuda.synchronize()
            bad = env.sum_over_ranks(_equal_des(ys, yw))
            return {"verified_entries": len(pick) * world, "verified_mismatches": int(bad),
                    "against": "%d sampled outputs per rank recomputed in REFERENCE_ORDER mode (digits, sign, exponent)" % len(pick)}

    ctx.set_profiling(True)
    for _ in range(max(3, warmup)):
        step()
    env.barrier()
    launches0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    env.barrier()
    t_host0 = time.time()
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    env.barrier()
    t_host1 = time.time()
    ms_total = e0.elapsed_time(e1)
    launches = ctx.launch_count - launches0
    fallback = ctx.last_fallback_count()
    try:
        stage_ms, _ = ctx.last_stage_ms()
    except Exception:
        stage_ms = None
    ctx.set_profiling(False)
    ms_step = env.max_over_ranks(ms_total) / steps
    clocks = clocks_under_load(env, t_host0, t_host1, step, ms_step, steps)
    value = flops / (ms_step * 1e-3) / 1e9
    ver = verify() if verify else None
    e2e = None
    host_bytes = sum(c for _, c in host_arrays) * rs
    if want_e2e and host_bytes <= 8e9:
        bufs = [torch.empty(c * rs, dtype=torch.uint8).pin_memory() for _, c in host_arrays]
        hout = torch.empty(out_arr[1] * rs, dtype=torch.uint8).pin_memory()
        for (arr, c), b in zip(host_arrays, bufs):
            arr.device2host_ptr(b.data_ptr(), c)

        def e2e_step():
            for (arr, c), b in zip(host_arrays, bufs):
                arr.host2device_ptr(b.data_ptr(), c)
            step()
            out_arr[0].device2host_ptr(hout.data_ptr(), out_arr[1])
        e2e_step(); env.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        env.barrier()
        dt = env.max_over_ranks((time.perf_counter() - t0) / e2e_steps)
        e2e = {"value": flops / dt / 1e9, "unit": "MP-GFLOP/s", "h2d_bytes_per_step": int(host_bytes * world), "d2h_bytes_per_step": int(out_arr[1] * rs * world),
               "ms_per_step": dt * 1e3, "steps": e2e_steps, "path": "mpres_array_host2device(operands) + the call + mpres_array_device2host(result), pinned host AoS mp_float_t[]"}
    elif want_e2e:
        e2e = {"value": None, "unit": "MP-GFLOP/s", "skipped": "host copy of the operands is %.1f GB per step (PCIe-bound by construction); run with a smaller workload" % (host_bytes / 1e9)}
    hbm, bf16, src = read_peaks()
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
    cpu = cpu_dot_leg(N, bits) if (world == 1 and want_cpu and rank == 0) else None
    line = {"metric": "mp_%s MP-GFLOP/s" % op.split("_")[0], "value": value, "unit": "MP-GFLOP/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "int32 RNS residues, u64 lazy accumulation (f64 interval bounds)", "data": "synthetic", "config": config,
            "gpu_launches": int(launches), "fallback_elements_last_step": int(fallback), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "e2e": e2e}
    if ver:
        line.update(ver)
    return line


# ---- mp_gemm ------------------------------------------------------------------------------------------------------------------------

def run_gemm(env, workload, steps, warmup, full_precision=False, want_e2e=True, want_cpu=True, bcast_inside=False, ref_gpu=False, e2e_steps=2):
    args = env.args
    torch, pkg, ta, dist, parallel = env.torch, env.pkg, env.ta, env.dist, env.parallel
    rank, world = env.rank, env.world
    m, n, k, N = WORKLOADS[workload]
    precision = precision_of(N)
    bits = precision if full_precision else precision // 4
    sharded = world > 1 and not bcast_inside and n % world == 0 and m % world == 0
    config = {"workload": workload, "op": "mp_gemm", "m": m, "n": n, "k": k, "moduli": N, "precision_bits": precision,
              "input_significand_bits": bits, "l2": "operands (%.2f GB per matrix) exceed the 126 MB L2; no flush needed" % (m * k * (4 * N + 40) / 1e9), "mode": args.mode}
    if world == 1:
        config["sharding"] = "single GPU"
    elif sharded:
        config["sharding"] = ("A,C row blocks x%d; B resident on every rank (replicated once, outside the step); inside every step each rank converts its n/%d column "
                              "block of B and the packages (one-byte planes, shift planes, windows) are exchanged over NVLink by the copy engines while the tensor kernel "
                              "multiplies the panels that have arrived (mpres_gemm_sharded)" % (world, world))
    else:
        config["sharding"] = "A,C row blocks x%d; B lives on rank 0 and is broadcast (NCCL, all four SoA arrays) inside every step, then every rank runs mpres_gemm on its row block" % world
    ctx = env.ctx(N)
    ctx.set_mode({"auto": pkg.MODE_AUTO, "reference_order": pkg.MODE_REFERENCE_ORDER, "fast": pkg.MODE_FAST}[args.mode])
    ctx.set_stage2_kernel({"small": pkg.STAGE2_SMALL, "small_tiled": pkg.STAGE2_SMALL_TILED, "small_k64": pkg.STAGE2_SMALL_K64, "small_t128": pkg.STAGE2_SMALL_T128, "umma": pkg.STAGE2_UMMA, "umma_unstacked": pkg.STAGE2_UMMA_UNSTACKED, "mma_sync": pkg.STAGE2_MMA_SYNC}[args.stage2])
    ctx.set_stage3_kernel(args.stage3)
    config["stage2_kernel"], config["stage3_kernel"] = args.stage2, args.stage3
    assert m % world == 0
    mr = m // world                       # rows of A and C owned by this rank
    A = ta.TorchMpArray(ctx, mr * k)
    B = ta.TorchMpArray(ctx, k * n)
    C = ta.TorchMpArray(ctx, mr * n)
    C0 = ta.TorchMpArray(ctx, mr * n)     # pristine C, copied into C at the start of every step
    alpha, beta = ta.TorchMpArray(ctx, 1), ta.TorchMpArray(ctx, 1)
    ta.random_fill(ctx, A, bits, 1000 + rank)
    ta.random_fill(ctx, C0, bits, 2000 + rank)
    ta.random_fill(ctx, alpha, bits, 31)
    ta.random_fill(ctx, beta, bits, 32)
    if rank == 0 or sharded:
        ta.random_fill(ctx, B, bits, 33)      # the same seed on every rank: B replicated without a transfer
    stream = torch.cuda.current_stream().cuda_stream
    shard = None
    if sharded:
        shard = pkg.Shard(ctx, rank, world, n, k)
        handles = [None] * world
        dist.all_gather_object(handles, shard.export())
        shard.connect(handles)
        dist.barrier()

    # mp_gemm updates C in place, so every step gets its own pristine C: a ring of pre-filled copies (HBM has the room:
    # 2.8 GB each at config 3); only when the run has more steps than buffers is a buffer restored (device copy) before reuse
    free_b, _ = torch.cuda.mem_get_info()
    c_bytes = C0.nbytes()
    n_buf = int(max(1, min(steps + max(3, warmup), 24, (free_b - (24 << 30)) // max(1, c_bytes))))
    ring = [C] + [ta.TorchMpArray(ctx, mr * n) for _ in range(n_buf - 1)]
    for Cb in ring:
        for dst, src in zip(Cb.tensors(), C0.tensors()):
            dst.copy_(src)
    state = {"i": 0, "last": C}
    config["step"] = "%sC_i = alpha*A*B + beta*C_i on a pristine C_i (ring of %d pre-filled device buffers)" % ("" if world == 1 or sharded else "[B broadcast] ", n_buf)

    def step():
        i = state["i"]; state["i"] = i + 1
        Cb = ring[i % n_buf]
        state["last"] = Cb
        if i >= n_buf:                                    # buffer reuse: restore the pristine C first (device-to-device)
            for dst, src in zip(Cb.tensors(), C0.tensors()):
                dst.copy_(src, non_blocking=True)
        if shard is not None:
            shard.gemm(pkg.mblas_no_trans, pkg.mblas_no_trans, mr, n, k, alpha, A, mr, B, k, beta, Cb, mr, stream)
            return
        if world > 1:
            parallel.broadcast_arrays(dist, B.tensors(), 0)
        pkg.mp_gemm(ctx, pkg.mblas_no_trans, pkg.mblas_no_trans, mr, n, k, alpha, A, mr, B, k, beta, Cb, mr, None, stream)

    ctx.set_profiling(True)                               # (on during the warm-up too: the same load as the timed steps)
    for _ in range(max(3, warmup)):
        step()
    env.barrier()
    launches0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    env.barrier()
    use_profiler_range = os.environ.get("MPRES_BENCH_PROFILER_RANGE") == "1"   # ncu --profile-from-start off
    if use_profiler_range:
        torch.cuda.profiler.start()
    t_host0 = time.time()
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    env.barrier()
    t_host1 = time.time()
    if use_profiler_range:
        torch.cuda.profiler.stop()
    ms_total = e0.elapsed_time(e1)
    launches = ctx.launch_count - launches0
    fallback = ctx.last_fallback_count()
    slow_listed = ctx.last_slow_count()
    base_size = ctx.last_base_size()          # moduli stages 1-2 actually ran on (limb-plane path)
    small_P, small_nin = ctx.last_small_base()   # one-byte moduli of the small-modulus stage 2 (0: not used)
    try:
        stage_ms, s2_launches = ctx.last_stage_ms()
    except Exception:
        stage_ms, s2_launches = None, 0
    kern_ms = ctx.last_kernel_ms()
    ctx.set_profiling(False)
    ms_step = env.max_over_ranks(ms_total) / steps
    clocks = clocks_under_load(env, t_host0, t_host1, step, ms_step, steps)
    value = 2.0 * m * n * k / (ms_step * 1e-3) / 1e9

    # ---- sampled check of the last step's C against the reference-order k-loop (src/blas/gemm.cuh:39-58 semantics) on this rank's rows ----
    verified = None
    if args.mode == "auto" and not args.no_verify:
        g = torch.Generator(device="cpu"); g.manual_seed(4242 + rank)
        rows = sorted(torch.randperm(mr, generator=g)[:VERIFY_ROWS].tolist())
        cols = sorted(torch.randperm(n, generator=g)[:VERIFY_COLS].tolist())
        binary = ctx.last_binary_rounding()
        As = _sub_matrix(ta, ctx, A, mr, k, rows, list(range(k)))
        Bs = _sub_matrix(ta, ctx, B, k, n, list(range(k)), cols)
        Cs = _sub_matrix(ta, ctx, C0, mr, n, rows, cols)
        C0s = _sub_matrix(ta, ctx, C0, mr, n, rows, cols) if binary else None
        Cw = _sub_matrix(ta, ctx, state["last"], mr, n, rows, cols)
        ctx.set_mode(pkg.MODE_REFERENCE_ORDER)
        pkg.mp_gemm(ctx, pkg.mblas_no_trans, pkg.mblas_no_trans, len(rows), len(cols), k, alpha, As, len(rows), Bs, k, beta, Cs, len(rows), None, stream)
        ctx.set_mode(pkg.MODE_AUTO)
        torch.cuda.synchronize()
        bad = int(env.sum_over_ranks(_equal_des(Cs, Cw)))
        note = "" if not full_precision else " -- p-bit inputs: every element runs the reference-order k-loop or the single-rounding path; accuracy is pinned by tests"
        verified = {"verified_entries": len(rows) * len(cols) * world, "verified_mismatches": bad,
                    "verified_against": "%d x %d sampled entries of the last step's C per rank, recomputed by the reference-order k-loop (mp_mul, mp_add, rounding after each) "
                                        "on the gathered rows of A / columns of B: digits, sign, exponent%s" % (len(rows), len(cols), note)}
        if binary:
            # the exact sums were rounded ONCE in binary (csrc/kernels_bin.cuh): not the reference's bits, so the check is the reference's error model
            cnt, fails, wg, wr = _tolerance_check(ctx, k, alpha, beta, As, Bs, C0s, Cw, Cs, len(rows), len(cols))
            verified = {"verified_entries": cnt, "verified_mismatches": fails, "bitwise_differences_from_reference_order": bad,
                        "worst_error_over_bound": wg, "worst_error_over_bound_reference_order": wr,
                        "verified_against": "single-rounding stage 3 (p-bit inputs): %d sampled entries of the last step's C against exact rational arithmetic, |err| <= "
                                            "gamma_(k+3) (|alpha| sum |a||b| + |beta c|) with u = 4 / sqrt(M) (the reference's model, tests/blas/accuracy/test_dot_accuracy.cu:41-72); "
                                            "worst error / bound reported for this library and for the reference-order k-loop on the same entries" % cnt}
        del As, Bs, Cs, Cw, C0s

    # ---- end-to-end through the C-ABI with host buffers ------------------------------------------------
    e2e = None
    if want_e2e:
        rs = 4 * N + 40
        hA = torch.empty(mr * k * rs, dtype=torch.uint8).pin_memory()
        hB = torch.empty(k * n * rs, dtype=torch.uint8).pin_memory()
        hC = torch.empty(mr * n * rs, dtype=torch.uint8).pin_memory()
        hOut = torch.empty(mr * n * rs, dtype=torch.uint8).pin_memory()
        hal, hbe = torch.empty(rs, dtype=torch.uint8).pin_memory(), torch.empty(rs, dtype=torch.uint8).pin_memory()
        if world > 1 and not sharded:
            for t in B.tensors():
                dist.broadcast(t, src=0)
        A.device2host_ptr(hA.data_ptr(), mr * k); B.device2host_ptr(hB.data_ptr(), k * n)
        C0.device2host_ptr(hC.data_ptr(), mr * n)
        alpha.device2host_ptr(hal.data_ptr(), 1); beta.device2host_ptr(hbe.data_ptr(), 1)
        split_b = args.e2e_path == "pipelined" and world > 1 and n % world == 0

        def e2e_step():
            if split_b:
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
        env.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        env.barrier()
        dt = env.max_over_ranks((time.perf_counter() - t0) / e2e_steps)
        b_up = k * n // world if split_b else k * n
        # what crossed the PCIe link: mpres_gemm_host cuts the records of A (and of B when it uploads B) down to the fields the fast path reads
        lean = ctx.last_host_upload_residues() if args.e2e_path == "pipelined" else 0
        ls = (4 * lean + 24) if lean > 0 else rs
        h2d = mr * k * ls + b_up * (rs if split_b else ls) + (mr * n + 2) * rs
        d2h = mr * n * rs
        e2e = {"value": 2.0 * m * n * k / dt / 1e9, "unit": "MP-GFLOP/s", "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
               "ms_per_step": dt * 1e3, "steps": e2e_steps, "operand_residues_uploaded": lean if lean > 0 else N,
               "host_pack": ("A and B cross the link as %d-byte records (first %d residues, sign, exponent, upper interval bound of the %d-byte mp_float_t), cut down by the "
                             "host cores inside the call; C travels in full both ways" % (ls, lean, rs)) if lean > 0 else "full records",
               "path": ("per rank: mpres_array_host2device(1/N of B) + all-gather of B over NVLink + mpres_gemm_host_bdev(alpha, A_r, B, beta, C_r -> out_r), pinned host AoS mp_float_t[]"
                        if split_b else
                        "mpres_gemm_host(alpha, A, B, beta, C -> out): pinned host AoS mp_float_t[], PCIe transfers pipelined with the compute by column panels"
                        if args.e2e_path == "pipelined" else
                        "mpres_array_host2device(A,B,C,alpha,beta) + mpres_gemm + mpres_array_device2host(C), pinned host AoS mp_float_t[]")}
        del hA, hB, hC, hOut

    # ---- reference GPU kernels on the same device and inputs (the reference's own v1 mp_gemm, launch configuration of
    #      tests/blas/performance/test_gemm_performance.cu:53-57), single GPU, small configurations only ----
    ref_gpu_block = None
    if ref_gpu and world == 1 and rank == 0:
        try:
            import oracle
            if oracle.have_ref(N):
                ref = oracle.RefLib(N, gpu=True)
                # p-bit inputs: the reference kernel rounds after every product (minutes at 1024^3); a 128 x 128 block of C with the full k is
                # timed instead and scaled by its share of the entries
                ms_, ns_ = (128, 128) if full_precision else (m, n)
                rows, cols = list(range(ms_)), list(range(ns_))
                hA = _sub_matrix(ta, ctx, A, mr, k, rows, list(range(k))).device2host()
                hB = _sub_matrix(ta, ctx, B, k, n, list(range(k)), cols).device2host()
                hC = _sub_matrix(ta, ctx, C0, mr, n, rows, cols).device2host()
                hal, hbe = alpha.device2host(), beta.device2host()
                _, _, ms_ref = ref.gpu_gemm(ms_, ns_, k, hal, hA, hB, hbe, hC, repeat=1 if full_precision else 2)
                scale = (m * n) / float(ms_ * ns_)
                ms_ref *= scale
                ref_gpu_block = {"ms_per_call": ms_ref, "value": 2.0 * m * n * k / (ms_ref * 1e-3) / 1e9, "unit": "MP-GFLOP/s",
                                 "what": "reference v1 cuda::mp_gemm<32,1,128,64,16> (unmodified, oracle/_ref) on this GPU, same inputs" +
                                         ("" if scale == 1 else "; timed on a %d x %d block of C with the full k and scaled by %g" % (ms_, ns_, scale)),
                                 "speedup_of_this_library": ms_ref / ms_step}
        except Exception as e:      # noqa: BLE001
            ref_gpu_block = {"unavailable": repr(e)}

    if shard is not None:
        env.barrier()
        shard.close()
    line = None
    if rank == 0:
        hbm, bf16_peak, peak_src = read_peaks()
        int8_peak, int8_src = read_int8_peak(bf16_peak)
        roof, int32_roof, whole, per_kernel = None, None, None, None
        if kern_ms:
            agg = {}
            for name, ms in kern_ms:
                agg[name] = agg.get(name, 0.0) + ms
            per_kernel = {kname: round(v, 4) for kname, v in agg.items() if kname not in ("end",)}
            modelled = {kname: v for kname, v in agg.items() if kname in ("k_norm_fast", "k_ext_norm_small", "k_ext_small", "k_align_small(A)", "k_align_small(B)", "k_small_umma_p", "k_outer_info")}
            if modelled and small_P > 0:
                dom = max(modelled, key=modelled.get)
                dims = (mr, n, k, N, small_P, small_nin, world if sharded else 1)
                launches_of = sum(1 for kname, _ in kern_ms if kname == dom)
                roof = kernel_roofline(dom, modelled[dom] / max(1, launches_of), dims, hbm, int8_peak, int8_src, peak_src, workload)
                roof["why_this_kernel"] = "the kernel with the largest share of the step among those timed by CUDA events on the call's stream (per_kernel_ms, last step)"
                roof["tensor_kernel"] = kernel_roofline("k_small_umma_p", agg.get("k_small_umma_p", 0.0), dims, hbm, int8_peak, int8_src, peak_src, workload)
        if stage_ms and s2_launches:
            t2 = stage_ms[1] * 1e-3                       # all stage-2 launches of one mp_gemm call on this rank
            nb = base_size if 0 < base_size <= N else N
            if roof is None:
                limb_macs = 16.0 * mr * n * k * nb        # int8 MACs the limb-plane kernel executes: 16 limb products per residue MAC, nb moduli
                ops = 2.0 * limb_macs
                roof = {"bound": "tensor", "kernel": "k_limb_umma<stacked> (small base not selected)", "achieved": ops / t2 / 1e12, "peak": int8_peak, "unit": "TOP/s (int8)",
                        "frac": ops / t2 / 1e12 / int8_peak, "peak_source": int8_src, "traffic": None, "avg_launch_ms": stage_ms[1] / s2_launches,
                        "algorithmic_ops_per_launch": ops / s2_launches, "moduli_in_stage2": nb}
            roof["stage_ms"] = {"stage1_align": stage_ms[0], "stage2_multiply": stage_ms[1], "stage3_extend_normalise_epilogue": stage_ms[2]}
            int32_roof = {"definition": "SURVEY 8(d): m*n*k*N residue-MACs / t / R_mac, R_mac = measured IMAD.WIDE.U32 rate (all N moduli counted)",
                          "R_mac_per_s": R_MAC_IMAD_WIDE, "frac_stage2": mr * n * k * N / t2 / R_MAC_IMAD_WIDE,
                          "frac_whole_step": (m / world) * n * k * N / (ms_step * 1e-3) / R_MAC_IMAD_WIDE}
        rec = 4 * N + 40
        alg_bytes = (mr * k + k * n + 2 * mr * n) * rec
        t_step = ms_step * 1e-3
        whole = {"algorithmic_bytes_per_step_per_gpu": alg_bytes, "algorithmic_GBps": alg_bytes / t_step / 1e9, "frac_hbm": alg_bytes / t_step / 1e9 / hbm,
                 "int8_ops_per_step_per_gpu": 2.0 * small_P * mr * n * k if small_P > 0 else None,
                 "frac_int8": (2.0 * small_P * mr * n * k / t_step / 1e12 / int8_peak) if small_P > 0 else None,
                 "note": "A, B in; C in and out (4N+40 bytes per entry) and the int8 operations of stage 2, both over the whole step time"}
        cpu = None
        if world == 1 and want_cpu:
            cpu = cpu_reference_leg(N, CPU_SAMPLE[0], CPU_SAMPLE[1], k, bits, mpfr=True)
        line = {"metric": "mp_gemm MP-GFLOP/s", "value": value, "unit": "MP-GFLOP/s", "n_gpus": world, "steps": steps, "warmup": warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "u8 residues modulo one-byte moduli, s32 accumulate (int32 RNS residues, f64 interval bounds outside stage 2)", "data": "synthetic", "config": config,
                "gpu_launches": int(launches), "fallback_elements_last_step": int(fallback), "stage3_listed_elements_last_step": int(slow_listed),
                "limb_base_moduli": int(base_size), "small_base_moduli": int(small_P), "small_base_input_residues": int(small_nin), "clocks": clocks,
                "roofline": roof, "whole_step": whole, "per_kernel_ms": per_kernel, "int32_roofline": int32_roof, "cpu_baseline": cpu, "e2e": e2e}
        if verified:
            line.update(verified)
        if ref_gpu_block:
            line["reference_gpu_kernels"] = ref_gpu_block
    # free this workload's device memory before the next one
    del ring, A, B, C, C0
    torch.cuda.empty_cache()
    return line


def reference_arm(env, workload):
    """--impl reference: the reference's own CPU implementation on the box's host cores (rank 0 only)"""
    args = env.args
    if env.rank != 0:
        return
    if workload in VEC_WORKLOADS:
        op, m, n, N = VEC_WORKLOADS[workload]
        precision = precision_of(N)
        bits = precision // 4
        vals, last = [], None
        for i in range(args.warmup + args.steps):
            last = cpu_dot_leg(N, bits, seed=11 + 2 * i)
            if i >= args.warmup:
                vals.append(last["value"])
        v = statistics.mean(vals) if vals else last["value"]
        cb = dict(last); cb["value"] = v
        config = {"workload": workload, "op": "mp_" + op, "m": m, "n": n, "moduli": N, "precision_bits": precision, "input_significand_bits": bits}
        print(json.dumps({"impl": "reference", "metric": "mp_%s MP-GFLOP/s" % op.split("_")[0], "value": v, "unit": "MP-GFLOP/s", "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": 2.0 * (1 << 21) / v / 1e6, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                          "dtype": "int32 residues (host mp_float_t arithmetic)", "data": "synthetic", "config": config, "cpu_baseline": cb,
                          "e2e": {"value": v, "unit": "MP-GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    m, n, k, N = WORKLOADS[workload]
    precision = precision_of(N)
    bits = precision if args.full_precision_inputs else precision // 4
    config = {"workload": workload, "op": "mp_gemm", "m": m, "n": n, "k": k, "moduli": N, "precision_bits": precision, "input_significand_bits": bits}
    ms_s, ns_s = CPU_SAMPLE
    vals, last = [], None
    for i in range(args.warmup + args.steps):
        last = cpu_reference_leg(N, ms_s, ns_s, k, bits, seed=7 + i, mpfr=(i == args.warmup + args.steps - 1))
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


def slim(line, name):
    """a sub-result: the measured numbers of another configuration without the bulky blocks"""
    if line is None:
        return None
    keep = ("metric", "value", "unit", "ms_per_step", "n_gpus", "steps", "gpu_launches", "fallback_elements_last_step", "small_base_moduli", "limb_base_moduli",
            "verified_entries", "verified_mismatches", "reference_gpu_kernels", "per_kernel_ms")
    out = {"name": name}
    out.update({kk: line[kk] for kk in keep if kk in line})
    out["config"] = {kk: line["config"][kk] for kk in ("workload", "op", "m", "n", "k", "moduli", "precision_bits", "input_significand_bits", "sharding") if kk in line["config"]}
    if line.get("roofline"):
        out["roofline"] = {kk: line["roofline"].get(kk) for kk in ("bound", "kernel", "achieved", "peak", "unit", "frac")}
    if line.get("e2e") and line["e2e"].get("value"):
        out["e2e"] = {kk: line["e2e"][kk] for kk in ("value", "unit", "ms_per_step", "h2d_bytes_per_step", "d2h_bytes_per_step") if kk in line["e2e"]}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS) + sorted(VEC_WORKLOADS))
    ap.add_argument("--mode", default="auto", choices=["auto", "reference_order", "fast"])
    ap.add_argument("--stage2", default="small", choices=["small", "small_t128", "small_k64", "small_tiled", "umma", "umma_unstacked", "mma_sync"], help="stage-2 kernel (A/B measurement)")
    ap.add_argument("--stage3", type=int, default=0, choices=[0, 1, 2, 3, 4], help="stage-3 kernel variant (mpres_set_stage3_kernel; A/B measurement)")
    ap.add_argument("--bcast", default="packed", choices=["packed", "full"], help="N > 1: packed = B resident on every rank, per-step exchange of the converted column blocks "
                    "(mpres_gemm_sharded); full = B on rank 0, NCCL broadcast of all four SoA arrays inside every step")
    ap.add_argument("--full-precision-inputs", action="store_true", help="p-bit significands instead of p/4")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-verify", action="store_true")
    ap.add_argument("--no-sub", action="store_true", help="skip the sub_results (the other BASELINE configurations) of the default invocation")
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
    env = Env(args)
    workload = args.workload or DEFAULT_WORKLOAD
    if args.impl == "reference":
        return reference_arm(env, workload)
    env.start_gpu()
    default_invocation = args.workload is None and not args.no_sub and not args.full_precision_inputs and args.mode == "auto"
    if workload in VEC_WORKLOADS:
        line = run_vec(env, workload, args.steps, args.warmup, want_e2e=not args.no_e2e, want_cpu=not args.no_cpu_baseline, e2e_steps=args.e2e_steps)
    else:
        line = run_gemm(env, workload, args.steps, args.warmup, full_precision=args.full_precision_inputs, want_e2e=not args.no_e2e,
                        want_cpu=not args.no_cpu_baseline, bcast_inside=(args.bcast == "full"), e2e_steps=args.e2e_steps)
    subs = []
    if default_invocation:
        # the other BASELINE configurations, measured the same way in the same process (short runs; the headline above is unaffected)
        def sub(name, fn):
            try:
                r = fn()
                if env.rank == 0:
                    subs.append(slim(r, name))
            except Exception as e:      # noqa: BLE001
                if env.rank == 0:
                    subs.append({"name": name, "error": repr(e)})
        if env.world > 1:
            sub("config3_B_on_rank0_broadcast_inside_step", lambda: run_gemm(env, DEFAULT_WORKLOAD, 3, 3, want_e2e=False, want_cpu=False, bcast_inside=True))
        if env.world == 1:       # (configs 2 and 5 are single-GPU configurations: 1024 / 2048 columns leave a rank of 8 only one or two tensor tiles)
            sub("config2_gemm1024_106bit", lambda: run_gemm(env, "gemm1024_106bit", 10, 3, want_e2e=True, want_cpu=False, ref_gpu=True))
            sub("config2_gemm1024_106bit_full_precision_inputs", lambda: run_gemm(env, "gemm1024_106bit", 5, 3, full_precision=True, want_e2e=False, want_cpu=False, ref_gpu=True))
            sub("config3_gemm4096_424bit_full_precision_inputs", lambda: run_gemm(env, DEFAULT_WORKLOAD, 3, 3, full_precision=True, want_e2e=False, want_cpu=False))
        for w in ("gemv16384_212bit", "gemvt16384_212bit", "dot16m_212bit"):
            sub("config4_" + w, lambda w=w: run_vec(env, w, 5, 3, want_e2e=False, want_cpu=False))
        for w in ("gemm2048_106bit", "gemm2048_212bit", "gemm2048_318bit", "gemm2048_424bit", "gemm2048_530bit", "gemm2048_636bit", "gemm2048_742bit", "gemm2048_848bit"):
            if env.world == 1:
                sub("config5_" + w, lambda w=w: run_gemm(env, w, 5, 3, want_e2e=False, want_cpu=False))
    if env.rank == 0:
        if subs:
            line["sub_results"] = subs
        print(json.dumps(line))
    env.finish()


if __name__ == "__main__":
    main()
