#!/usr/bin/env python
"""Compile the reference's own wrapper_*.cpp harnesses UNMODIFIED (from where they lie under
/root/reference) against tests/compat/include (a stand-in Halide.h) and link them with
libtiramisu_b200.so — the drop-in claim of INTEGRATION.md, exercised.  TEST INFRASTRUCTURE.

Runs only where /root/reference exists (the build container); binaries go to tests/compat/_bin/
(git-ignored, shipped to the GPU box by gpurun) and are run there by tests/test_compat_wrappers.py.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("TIRAMISU_REFERENCE", "/root/reference")
BIN = os.path.join(HERE, "_bin")
LIBDIR = os.path.join(ROOT, "tiramisu_b200", "lib")
ORACLE = os.path.join(ROOT, "oracle")
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")

# name -> (wrapper source relative to the reference, extra include dirs relative to the reference, extra link flags)
WRAPPERS = {
    "wrapper_heat2d": ("benchmarks/halide/wrapper_heat2d.cpp", ["benchmarks/halide"], []),
    "wrapper_heat2dgpu": ("benchmarks/halide/wrapper_heat2dgpu.cpp", ["benchmarks/halide"], []),
    # MPI program: one process per rank (tests/compat/include/mpi.h maps its MPI calls onto tiramisu_MPI_*)
    "wrapper_heat2d_dist": ("benchmarks/halide/wrapper_heat2d_dist.cpp", ["benchmarks/halide"], ["-DTOTAL_ITERATIONS=2"]),
    "sgemm_wrapper": ("benchmarks/linear_algebra/blas/level3/sgemm/cpu/sgemm_wrapper.cpp",
                      ["benchmarks/linear_algebra/blas/level3/sgemm/cpu"], []),
    "sgemm_gpu_wrapper": ("benchmarks/linear_algebra/blas/level3/sgemm/gpu/wrapper.cpp",
                          ["benchmarks/linear_algebra/blas/level3/sgemm/gpu"],
                          [f"-L{CUDA}/lib64", "-lcublas", "-lcudart"]),
    "lstm_wrapper": ("benchmarks/DNN/blocks/LSTM/cpu/wrapper.cpp", ["benchmarks/DNN/blocks/LSTM/cpu"], []),
    "wrapper_test_162": ("tests/wrapper_test_162.cpp", ["tests"], []),
    "wrapper_test_164": ("tests/wrapper_test_164.cpp", ["tests"], []),
    "wrapper_test_167": ("tests/wrapper_test_167.cpp", ["tests"], []),
    "wrapper_test_168": ("tests/wrapper_test_168.cpp", ["tests"], []),
    "wrapper_test_197": ("tests/wrapper_test_197.cpp", ["tests"], []),
}


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "benchmarks")) and os.path.exists(os.path.join(LIBDIR, "libtiramisu_b200.so"))


def build(names=None, verbose=False) -> dict:
    os.makedirs(BIN, exist_ok=True)
    env = dict(os.environ)
    env.pop("CC", None); env.pop("CXX", None)
    common = ["g++", "-std=c++17", "-O2", "-mavx", "-fopenmp", "-w",
              "-I", os.path.join(HERE, "include"), "-I", os.path.join(REF, "3rdParty/Halide/src/runtime"),
              "-I", os.path.join(REF, "include"), "-I", os.path.join(REF, "benchmarks"), "-I", os.path.join(CUDA, "include")]
    support = [os.path.join(HERE, "compat_support.cpp"), os.path.join(REF, "src/tiramisu_utils.cpp"),
               os.path.join(REF, "src/tiramisu_debug.cpp")]
    out = {}
    for name, (src, incs, extra) in WRAPPERS.items():
        if names and name not in names:
            continue
        exe = os.path.join(BIN, name)
        cmd = common + sum((["-I", os.path.join(REF, i)] for i in incs), []) + [os.path.join(REF, src)] + support + [
            "-o", exe, f"-L{LIBDIR}", "-ltiramisu_b200", f"-L{ORACLE}", "-ltiramisu_oracle",
            "-Wl,-rpath,$ORIGIN/../../../tiramisu_b200/lib", "-Wl,-rpath,$ORIGIN/../../../oracle",
            f"-Wl,-rpath,{CUDA}/lib64", "-Wl,--allow-shlib-undefined", "-lpthread", "-ldl"] + extra
        r = subprocess.run(cmd, capture_output=True, text=True, env=env)
        out[name] = (r.returncode, (r.stdout + r.stderr)[-2000:])
        if verbose:
            print(name, "OK" if r.returncode == 0 else "FAILED\n" + out[name][1])
    return out


if __name__ == "__main__":
    if not available():
        print("reference checkout or libtiramisu_b200.so missing; nothing built")
        sys.exit(0)
    res = build(sys.argv[1:] or None, verbose=True)
    sys.exit(1 if any(rc for rc, _ in res.values()) else 0)
