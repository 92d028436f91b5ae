#!/usr/bin/env python
"""Build the repo's own front-end test generators (tests/frontend_cases/*_gen.cpp): compile each
against tiramisu_b200/frontend, run it (codegen lowers to sm_100a and calls nvcc), and link the
object into tests/frontend_cases/_build/lib<name>.so next to libtiramisu_b200.so so the tests can
call the generated function through ctypes.  TEST INFRASTRUCTURE; needs g++ and nvcc, no GPU and no
reference checkout."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "_build")
LIBDIR = os.path.join(ROOT, "tiramisu_b200", "lib")
FRONT_INC = os.path.join(ROOT, "tiramisu_b200", "frontend", "include")

# library name -> (generator source, argv[1] for the generator or None, lowering mode or None)
CASES = {
    "heat2d_gen": ("heat2d_gen.cpp", "heat2d_gen", None),                  # statement pair -> the runtime's heat2d kernel
    "heat2d_gen_generic": ("heat2d_gen.cpp", "heat2d_gen_generic", "generic"),   # the two kernels the lowering emits itself
    "blurxy_u8_gen": ("blurxy_u8_gen.cpp", None, None),
    "gemm_sym_gen": ("gemm_sym_gen.cpp", "gemm_sym_gen", None),
    "gemm_sym_generic": ("gemm_sym_gen.cpp", "gemm_sym_generic", "generic"),
    # CPU schedules (no gpu level, codegen without the gpu flag): re-targeted to sm_100a by the lowering
    "blur_f32_cpu_gen": ("blur_f32_cpu_gen.cpp", None, None),
    "matvec_cpu_gen": ("matvec_cpu_gen.cpp", "matvec_cpu_gen", None),                      # warp-shuffle reduction
    "matvec_cpu_generic": ("matvec_cpu_gen.cpp", "matvec_cpu_generic", "generic"),         # one row per thread, statement order
    "sgemm_cpu_sched_gen": ("sgemm_cpu_sched_gen.cpp", "sgemm_cpu_sched_gen", "generic"),   # three SIMT kernels, bit-exact
    "fused_reduce_cpu_gen": ("fused_reduce_cpu_gen.cpp", "fused_reduce_cpu_gen", None),    # fused i/j/k nest: i, j -> grid, k in the thread
    # advisor findings of round 1: a store index shared by two iterators; a fused run with unequal fuse depths
    "polymul_cpu_gen": ("polymul_cpu_gen.cpp", "polymul_cpu_gen", None),
    "prefix_fuse_cpu_gen": ("prefix_fuse_cpu_gen.cpp", "prefix_fuse_cpu_gen", None),
    "sgemm_cpu_sched_tc": ("sgemm_cpu_sched_gen.cpp", "sgemm_cpu_sched_tc", None),           # statement triple -> one tcgen05 GEMM
    # the register-blocking idioms of the reference's DNN generators / the old-style API of its sgemm generator
    "regblock_cpu_gen": ("regblock_cpu_gen.cpp", "regblock_cpu_gen", None),
    "pieces_cpu_gen": ("pieces_cpu_gen.cpp", "pieces_cpu_gen", None),
    # distributed functions (rank scope + tiramisu_MPI_* on device addresses), one build per rank count the GPU boxes offer
    **{f"dist_blur_gen_{k}": ("dist_blur_gen.cpp", f"dist_blur_gen_{k}", None) for k in (2, 4, 8)},     # gen_communication halo
    **{f"dist_xfer_gen_{k}": ("dist_xfer_gen.cpp", f"dist_xfer_gen_{k}", None) for k in (2, 4, 8)},     # explicit create_xfer
}


def _env(lowering=None):
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    if lowering:
        env["TIRAMISU_B200_LOWERING"] = lowering
    return env


def build(verbose=False) -> dict:
    os.makedirs(OUT, exist_ok=True)
    res = {}
    for name, (src, arg, lowering) in CASES.items():
        gen = os.path.join(OUT, name + "_generator")
        steps = [
            (["g++", "-std=c++17", "-O1", "-w", "-I", FRONT_INC, os.path.join(HERE, src), "-o", gen,
              f"-L{LIBDIR}", "-ltiramisu_b200_frontend", f"-Wl,-rpath,{LIBDIR}"], None),
            ([gen] + ([arg] if arg else []), OUT),
            (["g++", "-shared", "-o", os.path.join(OUT, f"lib{name}.so"), os.path.join(OUT, name + ".o"),
              f"-L{LIBDIR}", "-ltiramisu_b200", "-L/usr/local/cuda/lib64", "-lcudart",
              "-Wl,-rpath,$ORIGIN/../../../tiramisu_b200/lib", "-Wl,-rpath,/usr/local/cuda/lib64"], None),
        ]
        rc, log = 0, ""
        for cmd, cwd in steps:
            r = subprocess.run(cmd, capture_output=True, text=True, env=_env(lowering), cwd=cwd)
            log += r.stdout + r.stderr
            if r.returncode:
                rc = r.returncode
                break
        res[name] = (rc, log[-3000:])
        if verbose:
            print(name, "OK" if rc == 0 else "FAILED\n" + log[-3000:], flush=True)
    return res


def lib_path(name: str) -> str:
    return os.path.join(OUT, f"lib{name}.so")


def generated_source(name: str) -> str:
    return os.path.join(OUT, name + ".o.b200.cu")


if __name__ == "__main__":
    r = build(verbose=True)
    sys.exit(1 if any(rc for rc, _ in r.values()) else 0)
