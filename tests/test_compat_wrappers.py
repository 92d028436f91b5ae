"""The reference's own wrapper_*.cpp harnesses, compiled UNMODIFIED from /root/reference against a
stand-in Halide.h (tests/compat/include) and linked with libtiramisu_b200.so: the drop-in boundary
of INTEGRATION.md, exercised with the reference's code on both sides of it.

* not gpu: (only where /root/reference exists) every selected harness compiles and links.
* gpu    : the binaries built here by __graft_entry__.build() travel to the GPU box with the repo
           snapshot; each is run there and must report success by its own criteria."""
import os
import subprocess

import numpy as np
import pytest

from tests.compat import build_wrappers as bw

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.skipif(not bw.available(), reason="needs /root/reference and the built library")
def test_reference_wrappers_compile_and_link_unmodified():
    res = bw.build()
    failed = {k: v[1] for k, v in res.items() if v[0] != 0}
    assert not failed, failed
    assert len(res) == len(bw.WRAPPERS)


def _run(name, args=(), cwd=None, timeout=600, env=None):
    exe = os.path.join(bw.BIN, name)
    if not os.path.exists(exe):
        pytest.skip(f"{exe} not built (built by __graft_entry__.build() where /root/reference exists)")
    r = subprocess.run([exe, *args], cwd=cwd, capture_output=True, text=True, timeout=timeout,
                       env=dict(os.environ, **(env or {})))
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-3000:]
    assert "failed" not in out.lower(), out[-3000:]
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["wrapper_test_162", "wrapper_test_164", "wrapper_test_167", "wrapper_test_168",
                                  "wrapper_test_197"])
def test_reference_feature_test_wrappers(name, tmp_path):
    """tests/wrapper_test_16x.cpp: cublas_gemm / gpu_tile / transpose / allocate tests; each compares
    with its own scalar reference through compare_buffers (exact)."""
    out = _run(name, cwd=tmp_path)
    assert "succeeded" in out.lower() or "passed" in out.lower(), out[-2000:]


@pytest.mark.gpu
def test_reference_heat2d_wrapper(tmp_path):
    """benchmarks/halide/wrapper_heat2d.cpp: 10000x10000, NB_TESTS runs, compare_buffers_approximately
    against heat2d_ref (here: the oracle)."""
    out = _run("wrapper_heat2d", cwd=tmp_path)
    assert "heat2d" in out


@pytest.mark.gpu
def test_reference_heat2dgpu_wrapper(tmp_path):
    out = _run("wrapper_heat2dgpu", cwd=tmp_path)
    assert "heat2dgpu" in out


@pytest.mark.gpu
def test_reference_sgemm_gpu_wrapper_matches_cublas(tmp_path):
    """benchmarks/linear_algebra/blas/level3/sgemm/gpu/wrapper.cpp with check_correctness=1: 3072^3,
    rand()%10-5 operands, exact comparison against cuBLAS (the harness's own reference)."""
    out = _run("sgemm_gpu_wrapper", args=("1", "1"), cwd=tmp_path)
    assert "cuBLAS and Tiramisu outputs match" in out, out[-2000:]


@pytest.mark.gpu
@pytest.mark.parametrize("mode,min_pct", [("1", 99.9), ("0", 97.0)])
def test_reference_lstm_wrapper(tmp_path, mode, min_pct):
    """benchmarks/DNN/blocks/LSTM/cpu/wrapper.cpp (F=32, B=8, L=4, T=100): the harness compares its
    output file with mkldnn_result.txt at |diff| < 0.02 and prints the percentage of matching
    entries; the MKL-DNN program cannot run here, so the file is produced by the oracle on the
    harness's own srand(0) inputs.  The unscaled +-1 weights make the recurrence chaotic over 100
    steps, so the TF32 path (mode 0) is held to 97 % of entries and the exact fp32 path
    (TB200_LSTM_MODE=1) to 99.9 % (the reference's own GPU harness carries a "result mismatch"
    TODO for this block, LSTM/gpu/wrapper.cpp:77-88)."""
    from oracle import oracle
    from tests import refdata
    W, b, x = refdata.lstm_inputs(32, 8, 4, 100)
    want = oracle.lstm(W, b, x)                      # [T][B][F]
    with open(tmp_path / "mkldnn_result.txt", "w") as f:
        for v in want.reshape(-1):
            f.write(f"{float(v):.9g}\n")
    out = _run("lstm_wrapper", cwd=tmp_path, env={"TB200_LSTM_MODE": mode})
    pct = [ln for ln in out.splitlines() if "Percentage of correctness" in ln]
    assert pct and float(pct[0].split()[-1]) >= min_pct, out[-2000:]


@pytest.mark.gpu
def test_reference_sgemm_cpu_wrapper(tmp_path):
    """benchmarks/linear_algebra/blas/level3/sgemm/cpu/sgemm_wrapper.cpp: 16 sizes x 20 runs, A=B=C=1,
    a=b=3 => 3K+3, compare_buffers against cblas_sgemm (~1 minute: the CPU reference side dominates)."""
    out = _run("sgemm_wrapper", cwd=tmp_path, timeout=3000)
    assert out.count("Test sgemm succeeded") == 16, out[-3000:]
