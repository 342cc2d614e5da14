"""Host-side logic of bench.py that needs no GPU: the clock sampler's windows, the untimed repeat when the timed region holds too few samples,
the traffic lookup in the committed ncu summary, and the reference arm's JSON line (the CPU leg the driver runs with --impl reference)."""
import json
import os
import subprocess
import sys
import time

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def _row(sm, mx=1965, pw=400.0, reasons=("Not Active",) * 4):
    return ["0", str(sm), str(mx), str(pw)] + list(reasons)


class _FakeProc:
    pass


def _sampler(rows):
    s = bench.ClockSampler(0, period_ms=1)
    s.proc = _FakeProc()
    s.rows = rows
    return s


def test_clock_window_takes_only_the_samples_inside():
    t = time.time()
    s = _sampler([(t - 5.0, _row(1000)), (t - 1.0, _row(1950)), (t - 0.9, _row(1965)), (t - 0.8, _row(1965, reasons=("Active", "Not Active", "Not Active", "Active"))),
                  (t + 5.0, _row(500))])
    c = s.window(t - 1.05, t - 0.75)
    assert c["samples"] == 3 and c["sm_mhz"] == 1965.0 and c["sm_max_mhz"] == 1965.0
    assert c["reasons"] == ["hw_slowdown", "sw_power_cap"] and c["window"] == "timed region"
    assert s.window(t - 4.0, t - 3.0)["samples"] == 0


def test_clock_sampler_without_nvidia_smi_reports_it():
    s = bench.ClockSampler(0)
    s.proc = None
    c = s.window(0.0, time.time())
    assert c["sm_mhz"] is None and c["reasons"] == ["nvidia-smi unavailable"]


class _FakeEnv:
    def __init__(self, sampler):
        self.sampler, self.barriers = sampler, 0

    def max_over_ranks(self, v):
        return v

    def barrier(self):
        self.barriers += 1


def test_short_timed_region_is_repeated_untimed_until_it_holds_samples():
    s = _sampler([])
    env = _FakeEnv(s)
    calls = []

    def step():
        calls.append(1)
        s.rows.append((time.time(), _row(1965)))

    t0 = time.time()
    c = bench.clocks_under_load(env, t0 - 0.010, t0 - 0.005, step, ms_step=2.0, steps=5)
    assert len(calls) >= 5 and c["samples"] == len(calls) and c["sm_mhz"] == 1965.0
    assert c["window"].startswith("untimed repeat of the timed loop") and env.barriers == 2
    # a region with enough samples is reported as it is, nothing is re-run
    t1 = time.time()
    s2 = _sampler([(t1 - 0.3 + 0.05 * i, _row(1900 + i)) for i in range(4)])
    c2 = bench.clocks_under_load(_FakeEnv(s2), t1 - 0.31, t1 - 0.1, lambda: calls.append(2), ms_step=50.0, steps=4)
    assert c2["samples"] == 4 and c2["window"] == "timed region" and 2 not in calls


def test_traffic_comes_from_the_committed_ncu_summary():
    p = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
    d = json.load(open(p))
    assert d["workload"] == bench.DEFAULT_WORKLOAD
    k = d["kernels"]["k_norm_fast"]
    assert bench.read_ncu_traffic(bench.DEFAULT_WORKLOAD, "k_norm_fast") == float(k["dram_bytes_read"]) + float(k["dram_bytes_write"])
    assert bench.read_ncu_traffic(bench.DEFAULT_WORKLOAD, "k_small_umma_p<128,256>") is not None     # template arguments are ignored
    assert bench.read_ncu_traffic("gemm1024_106bit", "k_norm_fast") is None                            # another workload: no figure, not a wrong one


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libmpres_ref_N32.so")), reason="oracle/_ref not built")
def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["impl"] == "reference" and line["metric"] == "mp_gemm MP-GFLOP/s" and line["higher_is_better"] is True
    assert line["config"]["workload"] == bench.DEFAULT_WORKLOAD
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["value"] == line["value"] > 0
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
