"""Builds tiramisu_b200/lib/libtiramisu_b200.so (sm_100a only) with nvcc, in-tree.

    python -m tiramisu_b200.build [-v] [--force]

nvcc cross-compiles without a GPU.  Objects are rebuilt when a source or header is newer.
"""
from __future__ import annotations

import concurrent.futures as cf
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libtiramisu_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-fvisibility=hidden", "-Xptxas", "-v",
          "-I", os.path.join(ROOT, "include"), "-I", CSRC, "--expt-relaxed-constexpr"]


def _newest(paths):
    return max(os.path.getmtime(p) for p in paths)


def build(verbose: bool = False, force: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdrs = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(ROOT, "include", "**", "*.h"), recursive=True)
    hdr_time = _newest(hdrs)
    env = dict(os.environ)
    env.pop("CC", None)      # the image's $CC wrapper is not a usable nvcc host compiler
    env.pop("CXX", None)

    def compile_one(src):
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        if (not force and os.path.exists(obj) and os.path.getmtime(obj) > os.path.getmtime(src)
                and os.path.getmtime(obj) > hdr_time):
            return obj, ""
        cmd = [NVCC, *ARCH, *CFLAGS, "-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True, env=env)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj, r.stderr

    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(compile_one, srcs))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            if log:
                sys.stderr.write(log)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < _newest(objs):
        cmd = [NVCC, *ARCH, "-shared", "-o", LIB, *objs, "-ldl", "-lpthread"]
        r = subprocess.run(cmd, capture_output=True, text=True, env=env)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    # one-function ABI shims: the reference's LSTM/gpu_lib and LSTM/cpu_lib wrappers link against a symbol `lstm` of
    # another arity than LSTM/cpu's (csrc/abi/*.c)
    for name in ("lstm_gpu_lib", "lstm_cpu_lib"):
        src = os.path.join(CSRC, "abi", name + ".c")
        so = os.path.join(LIBDIR, f"libtiramisu_b200_{name}.so")
        if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(LIB)):
            r = subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-fvisibility=hidden", src, "-o", so, f"-L{LIBDIR}", "-ltiramisu_b200",
                                "-Wl,-rpath,$ORIGIN"], capture_output=True, text=True, env=env)
            if r.returncode != 0:
                raise RuntimeError(f"ABI shim {name} failed:\n{r.stdout}\n{r.stderr}")
    return LIB


FRONTEND = os.path.join(HERE, "frontend")
FRONTEND_LIB = os.path.join(LIBDIR, "libtiramisu_b200_frontend.so")


def build_frontend(force: bool = False) -> str:
    """The generator-time front-end + sm_100a lowering (tiramisu_b200/frontend): plain C++, no CUDA
    in the library itself — the code it generates is compiled by nvcc when a generator runs."""
    os.makedirs(LIBDIR, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(FRONTEND, "src", "*.cpp")))
    hdrs = glob.glob(os.path.join(FRONTEND, "include", "tiramisu", "*.h"))
    if (not force and os.path.exists(FRONTEND_LIB) and os.path.getmtime(FRONTEND_LIB) > _newest(srcs + hdrs)):
        return FRONTEND_LIB
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    cmd = ["g++", "-std=c++17", "-O2", "-g", "-fPIC", "-shared", "-Wall", "-Wno-unused-function",
           "-I", os.path.join(FRONTEND, "include"), *srcs, "-o", FRONTEND_LIB]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if r.returncode != 0:
        raise RuntimeError(f"front-end build failed:\n{r.stdout}\n{r.stderr}")
    return FRONTEND_LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="--force" in sys.argv))
    print(build_frontend(force="--force" in sys.argv))
