#!/usr/bin/env python
"""Sweep the reference's own feature tests through the front-end: for every tests/test_N.cpp under
/root/reference, compile the generator UNMODIFIED against tiramisu_b200/frontend, run it (lowering + nvcc),
and link tests/wrapper_test_N.cpp with the object.  TEST INFRASTRUCTURE (how much of the
reference's API surface the rebuilt front-end carries).  Most of these tests are CPU schedules with integer
data: they go through the CPU-schedule re-targeting of the lowering.

Outcome per test (tests/compat/_gen/reftests/status.json, not committed):
  "built"                      generator compiled, lowered, wrapper linked -> tests/compat/_bin/reftest_N
  "generator does not compile" the test uses API the front-end does not have (skew, compute_at, xfer, ...)
  "generator stopped: <msg>"   the front-end / lowering refused the program (non-box domain, ...)
  "wrapper does not link"
The binaries are run on the GPU box by tests/test_frontend.py against the pinned list
tests/compat/reference_tests_expected.json.
"""
from __future__ import annotations

import concurrent.futures as cf
import glob
import json
import os
import re
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("TIRAMISU_REFERENCE", "/root/reference")
WORK = os.path.join(HERE, "_gen", "reftests")
BIN = os.path.join(HERE, "_bin")
LIBDIR = os.path.join(ROOT, "tiramisu_b200", "lib")
FRONT_INC = os.path.join(ROOT, "tiramisu_b200", "frontend", "include")
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")
# already covered one by one (with their generic / pattern variants) by build_generators.py
SKIP = {"162", "164", "165", "166", "167", "168", "169", "170", "171", "197"}


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "tests")) and os.path.exists(os.path.join(LIBDIR, "libtiramisu_b200_frontend.so"))


def _env():
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    return env


def build_one(n: str) -> tuple[str, str]:
    gsrc = os.path.join(REF, "tests", f"test_{n}.cpp")
    wsrc = os.path.join(REF, "tests", f"wrapper_test_{n}.cpp")
    if not os.path.exists(wsrc):
        return n, "no wrapper"
    work = os.path.join(WORK, n)
    shutil.rmtree(work, ignore_errors=True)
    os.makedirs(os.path.join(work, "build"), exist_ok=True)
    env = _env()
    gen = os.path.join(work, "generator")
    r = subprocess.run(["g++", "-std=c++17", "-O0", "-w", "-I", FRONT_INC, "-I", os.path.join(FRONT_INC, "compat"), "-I", os.path.join(REF, "tests"),
                        gsrc, "-o", gen, f"-L{LIBDIR}", "-ltiramisu_b200_frontend", f"-Wl,-rpath,{LIBDIR}"], capture_output=True, text=True, env=env)
    if r.returncode:
        m = re.search(r"error: (.*)", r.stderr)
        return n, "generator does not compile: " + (m.group(1)[:160] if m else "?")
    try:
        r = subprocess.run([gen], capture_output=True, text=True, env=env, cwd=work, timeout=300)
    except subprocess.TimeoutExpired:
        return n, "generator stopped: timeout"
    objs = glob.glob(os.path.join(work, "build", "*.o")) + glob.glob(os.path.join(work, "*.o"))
    if r.returncode or not objs:
        msg = [ln for ln in (r.stderr + r.stdout).splitlines() if "tiramisu_b200" in ln or "error" in ln.lower()]
        return n, "generator stopped: " + (msg[-1][:200] if msg else f"rc={r.returncode}")
    exe = os.path.join(BIN, f"reftest_{n}")
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-w", "-I", os.path.join(HERE, "include"), "-I", os.path.join(REF, "3rdParty/Halide/src/runtime"),
                        "-I", os.path.join(REF, "include"), "-I", os.path.join(REF, "tests"), "-I", os.path.join(CUDA, "include"),
                        wsrc, *objs, os.path.join(HERE, "compat_support.cpp"), os.path.join(REF, "src/tiramisu_utils.cpp"),
                        os.path.join(REF, "src/tiramisu_debug.cpp"), "-o", exe, f"-L{LIBDIR}", "-ltiramisu_b200",
                        f"-L{os.path.join(ROOT, 'oracle')}", "-ltiramisu_oracle", "-Wl,-rpath,$ORIGIN/../../../tiramisu_b200/lib",
                        "-Wl,-rpath,$ORIGIN/../../../oracle", f"-L{CUDA}/lib64", "-lcudart", f"-Wl,-rpath,{CUDA}/lib64",
                        "-Wl,--allow-shlib-undefined", "-lpthread", "-ldl"], capture_output=True, text=True, env=env)
    if r.returncode:
        m = re.search(r"(error|undefined reference)[^\n]*", r.stderr)
        return n, "wrapper does not link: " + (m.group(0)[:160] if m else "?")
    return n, "built"


def build(verbose=False) -> dict:
    os.makedirs(WORK, exist_ok=True)
    os.makedirs(BIN, exist_ok=True)
    ids = sorted({re.match(r"test_(\w+)\.cpp", os.path.basename(p)).group(1) for p in glob.glob(os.path.join(REF, "tests", "test_*.cpp"))} - SKIP,
                 key=lambda s: (len(s), s))
    status = {}
    with cf.ThreadPoolExecutor(max_workers=6) as ex:
        for n, st in ex.map(build_one, ids):
            status[n] = st
            if verbose:
                print(n, st, flush=True)
    with open(os.path.join(WORK, "status.json"), "w") as f:
        json.dump(status, f, indent=1, sort_keys=True)
    return status


if __name__ == "__main__":
    if not available():
        print("reference checkout or the front-end library is missing; nothing built")
        sys.exit(0)
    st = build(verbose="-v" in sys.argv)
    kinds = {}
    for v in st.values():
        kinds[v.split(":")[0]] = kinds.get(v.split(":")[0], 0) + 1
    print(kinds)
