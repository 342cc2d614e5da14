"""include/mpres_compat.cuh: a caller written against the reference's own names (rns_const_init, cuda::mp_array_*, cuda::mp_gemm<...>,
cuda::mp_dot<...>) builds against the shim and libmpres_b200.so (CPU: compile + link) and, on the GPU, gives the oracle's records."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
SRC = os.path.join(ROOT, "tests", "compat", "compat_caller.cu")


def _build(out):
    libdir = os.path.join(ROOT, "mpres-blas_b200")
    cmd = [NVCC, "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-I", os.path.join(ROOT, "include"), SRC, "-L", libdir, "-lmpres_b200",
           "-Xlinker", "-rpath", "-Xlinker", libdir, "-o", out]
    subprocess.check_call(cmd)


def test_reference_style_caller_builds_against_the_shim(pkg, tmp_path):
    pkg.load_library()                      # the library must exist (no fallback)
    exe = str(tmp_path / "compat_caller")
    _build(exe)
    assert os.path.exists(exe)


@pytest.mark.gpu
def test_reference_style_caller_matches_the_oracle(pkg, tmp_path):
    import oracle
    from util import diff_fields, get_oracle, random_records
    exe = str(tmp_path / "compat_caller")
    _build(exe)
    N, m, n, k = 8, 37, 22, 150
    orc = get_oracle(N, oracle.DEVICE)
    bits = orc.precision // 4
    al, be = random_records(N, 1, bits, 901), random_records(N, 1, bits, 902)
    A, B, C = random_records(N, m * k, bits, 903), random_records(N, k * n, bits, 904), random_records(N, m * n, bits, 905)
    x, y = random_records(N, k, bits, 906), random_records(N, k, bits, 907)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        f.write(np.array([m, n, k], dtype=np.int32).tobytes())
        for arr in (al, be, A, B, C, x, y):
            f.write(np.ascontiguousarray(arr).tobytes())
    out = subprocess.run([exe, fin, fout], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert "MP_PRECISION %d" % orc.precision in out.stdout
    raw = np.fromfile(fout, dtype=orc.dtype)
    gotC, gotr = raw[: m * n], raw[m * n:]
    wantC, _ = orc.gemm(m, n, k, al, A, B, be, C)
    assert diff_fields(gotC, wantC, ("digits", "sign", "exp")).size == 0
    wantr = orc.dot_seq(x, y)
    assert diff_fields(gotr, np.array([wantr], dtype=orc.dtype), ("digits", "sign", "exp")).size == 0
