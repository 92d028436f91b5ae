#!/usr/bin/env python
"""Host emulation of GENERATED sources — TEST INFRASTRUCTURE, never part of the product.

The build container has no GPU, so a regression in the lowering (an index expression, a guard, the order of host
loops, a device-mirror copy) would only show up at the next GPU run.  This module takes the `x.o.b200.cu` a
generator run left behind, rewrites its `k<<<grid, block, 0, stream>>>(args)` launches into
`tb_emu_launch(k, grid, block, args)` (tests/emulation/include/cuda_runtime.h: the kernel body run for every
(block, thread) index, one after the other), compiles the result with g++ against host stand-ins for the runtime
ABI (runtime_emu.cpp) and runs it:

  * reference tests (tests/compat/build_reference_tests.py): linked with the reference's own wrapper_test_N.cpp,
    pass = the wrapper's own comparison;
  * the repo's frontend cases: built as a shared object and called through ctypes by tests/test_frontend_emulation.py.

Kernels that need SIMT semantics (`__syncthreads`, `__shared__`, `__shfl*`) are not emulatable by a sequential
thread loop and are reported as such.  The GPU tests (-m gpu) stay the parity gate.
"""
from __future__ import annotations

import concurrent.futures as cf
import glob
import json
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("TIRAMISU_REFERENCE", "/root/reference")
COMPAT = os.path.join(ROOT, "tests", "compat")
OUT = os.path.join(HERE, "_build")
ORACLE = os.path.join(ROOT, "oracle")

LAUNCH = re.compile(r"(\w+)<<<grid, block, 0, \(cudaStream_t\) tiramisu_b200_get_stream\(\)>>>\(")
SIMT = re.compile(r"__syncthreads|__shared__|__shfl|__ballot|__activemask|cooperative_groups")


class NotEmulatable(Exception):
    pass


def rewrite(src_path: str, dst_path: str) -> int:
    """Rewrite the kernel launches of a generated source; returns how many were rewritten."""
    text = open(src_path).read()
    if SIMT.search(text):
        raise NotEmulatable("kernel needs SIMT semantics (barrier / shared memory / shuffle)")
    text, n = LAUNCH.subn(r"tb_emu_launch(\1, grid, block, ", text)
    if "<<<" in text:
        raise NotEmulatable("a launch the rewriter does not know")
    os.makedirs(os.path.dirname(dst_path), exist_ok=True)
    with open(dst_path, "w") as f:
        f.write("// host-emulation copy made by tests/emulation/emulate.py — test infrastructure\n" + text)
    return n


def _gxx(args: list[str]) -> subprocess.CompletedProcess:
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    return subprocess.run(["g++", "-std=c++17", "-O1", "-w", "-ffp-contract=off", *args], capture_output=True, text=True, env=env)


def runtime_object() -> str:
    obj = os.path.join(OUT, "runtime_emu.o")
    src = os.path.join(HERE, "runtime_emu.cpp")
    if not os.path.exists(obj) or os.path.getmtime(obj) < os.path.getmtime(src):
        os.makedirs(OUT, exist_ok=True)
        r = _gxx(["-fPIC", "-c", src, "-o", obj])
        if r.returncode:
            raise RuntimeError(r.stderr)
    return obj


def build_shared(cu_paths: list[str], name: str, extra_sources: tuple = ()) -> str:
    """Generated source(s) -> tests/emulation/_build/<name>.emu.so (runtime stand-ins and oracle linked in)."""
    rt = runtime_object()
    srcs = []
    for p in cu_paths:
        dst = os.path.join(OUT, name, os.path.basename(p).replace(".cu", ".emu.cpp"))
        rewrite(p, dst)
        srcs.append(dst)
    so = os.path.join(OUT, f"{name}.emu.so")
    # -Bsymbolic: the stand-ins linked in here win over libtiramisu_b200.so's symbols of the same name, which an
    # earlier test of the same process may have loaded RTLD_GLOBAL
    r = _gxx(["-shared", "-fPIC", "-Wl,-Bsymbolic", "-I", os.path.join(HERE, "include"), *srcs, *extra_sources, rt, "-o", so, f"-L{ORACLE}", "-ltiramisu_oracle",
              f"-Wl,-rpath,{ORACLE}"])
    if r.returncode:
        raise RuntimeError(r.stderr[-2000:])
    return so


def regenerate(n: str, env_extra: dict) -> str:
    """Run reference test N's already-built generator again with extra environment (a lowering option) in a scratch
    directory; returns that directory (its build/*.b200.cu is what the option produces)."""
    gen = os.path.join(COMPAT, "_gen", "reftests", n, "generator")
    tag = "_".join(f"{k}-{v}" for k, v in sorted(env_extra.items()))
    work = os.path.join(OUT, "regen", tag, n)
    os.makedirs(os.path.join(work, "build"), exist_ok=True)
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    env.update(env_extra)
    r = subprocess.run([gen], capture_output=True, text=True, env=env, cwd=work, timeout=300)
    if r.returncode:
        raise RuntimeError(f"generator {n} failed: " + (r.stderr + r.stdout)[-400:])
    return work


def reference_test(n: str, gen_dir: str | None = None) -> tuple[str, str]:
    """Emulated build + run of reference test N; needs the generator's output: tests/compat/_gen/reftests/N (the
    sweep) or tests/compat/_gen/test_N (build_generators.py: the gpu-scheduled tests 162-171, 197)."""
    if gen_dir is None:
        gen_dir = os.path.join(COMPAT, "_gen", "reftests", n)
        if not os.path.isdir(gen_dir):
            gen_dir = os.path.join(COMPAT, "_gen", f"test_{n}")
    cus = glob.glob(os.path.join(gen_dir, "build", "*.b200.cu")) + glob.glob(os.path.join(gen_dir, "*.b200.cu"))
    if not cus:
        return n, "no generated source"
    try:
        srcs = []
        for p in cus:
            sub = os.path.join("reftests", n) if gen_dir.startswith(COMPAT) else os.path.join(os.path.relpath(gen_dir, OUT), "emu")
            dst = os.path.join(OUT, sub, os.path.basename(p).replace(".cu", ".emu.cpp"))
            rewrite(p, dst)
            srcs.append(dst)
    except NotEmulatable as e:
        return n, f"not emulatable: {e}"
    exe = os.path.join(os.path.dirname(srcs[0]), "run")
    r = _gxx(["-I", os.path.join(HERE, "include"), "-I", os.path.join(COMPAT, "include"), "-I", os.path.join(REF, "3rdParty/Halide/src/runtime"),
              "-I", os.path.join(REF, "include"), "-I", os.path.join(REF, "tests"), os.path.join(REF, "tests", f"wrapper_test_{n}.cpp"), *srcs,
              os.path.join(COMPAT, "compat_support.cpp"), os.path.join(REF, "src/tiramisu_utils.cpp"), os.path.join(REF, "src/tiramisu_debug.cpp"),
              runtime_object(), "-o", exe, f"-L{ORACLE}", "-ltiramisu_oracle", f"-Wl,-rpath,{ORACLE}", "-lpthread", "-ldl"])
    if r.returncode:
        m = re.search(r"(error|undefined reference)[^\n]*", r.stderr)
        return n, "does not build: " + (m.group(0)[:200] if m else r.stderr[-200:])
    try:
        r = subprocess.run([exe], capture_output=True, text=True, timeout=300, cwd=os.path.dirname(exe))
    except subprocess.TimeoutExpired:
        return n, "timeout"
    out = r.stdout + r.stderr
    if r.returncode == 0 and "Test failed" not in out and "rror" not in out:
        return n, "pass"
    tail = [ln for ln in out.splitlines() if ln.strip()]
    return n, f"FAIL rc={r.returncode}: " + (tail[-1][:160] if tail else "")


def reference_dist_test(n: str, ranks: int = 10, gen_dir: str | None = None) -> tuple[str, str]:
    """Emulated build of DISTRIBUTED reference test N (-DWITH_MPI, tests/emulation/mpi_emu.cpp as the message layer) and a
    run of the reference's own wrapper as `ranks` processes; the wrapper's compare_buffers decides."""
    import tempfile
    if gen_dir is None:
        gen_dir = os.path.join(COMPAT, "_gen", "reftests", n)
    cus = glob.glob(os.path.join(gen_dir, "build", "*.b200.cu"))
    if not cus:
        return n, "no generated source"
    srcs = []
    try:
        for p in cus:
            dst = os.path.join(OUT, "reftests_dist", n, os.path.basename(p).replace(".cu", ".emu.cpp"))
            rewrite(p, dst)
            srcs.append(dst)
    except NotEmulatable as e:
        return n, f"not emulatable: {e}"
    exe = os.path.join(os.path.dirname(srcs[0]), "run")
    r = _gxx(["-DWITH_MPI", "-I", os.path.join(HERE, "include"), "-I", os.path.join(COMPAT, "include"), "-I", os.path.join(REF, "3rdParty/Halide/src/runtime"),
              "-I", os.path.join(REF, "include"), "-I", os.path.join(REF, "tests"), os.path.join(REF, "tests", f"wrapper_test_{n}.cpp"), *srcs,
              os.path.join(COMPAT, "compat_support.cpp"), os.path.join(REF, "src/tiramisu_utils.cpp"), os.path.join(REF, "src/tiramisu_debug.cpp"),
              runtime_object(), os.path.join(HERE, "mpi_emu.cpp"), "-o", exe, f"-L{ORACLE}", "-ltiramisu_oracle", f"-Wl,-rpath,{ORACLE}", "-lpthread", "-ldl"])
    if r.returncode:
        m = re.search(r"(error|undefined reference)[^\n]*", r.stderr)
        return n, "does not build: " + (m.group(0)[:300] if m else r.stderr[-300:])
    with tempfile.TemporaryDirectory(prefix="tb200_mpi_emu_") as box:
        procs = []
        for rk in range(ranks):
            env = dict(os.environ, RANK=str(rk), WORLD_SIZE=str(ranks), TB200_EMU_MPI_DIR=box)
            procs.append(subprocess.Popen([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env, cwd=os.path.dirname(exe)))
        outs, bad = [], []
        for rk, pr in enumerate(procs):
            try:
                o, _ = pr.communicate(timeout=180)
            except subprocess.TimeoutExpired:
                pr.kill()
                o, _ = pr.communicate()
                o += "\n[timeout]"
            outs.append(o)
            if pr.returncode != 0 or "Test failed" in o or "rror" in o or "[timeout]" in o:
                tail = [ln for ln in o.splitlines() if ln.strip()]
                bad.append(f"rank {rk} rc={pr.returncode}: " + (tail[-1][:160] if tail else ""))
    return n, "pass" if not bad else "FAIL " + "; ".join(bad[:3])


def dist_wrapper_test(tag: str, cu_paths: list[str], wrapper: str, include_dirs: list[str], ranks: int, timeout: int = 600) -> str:
    """Emulated build of distributed generated sources with one of the reference's own -DWITH_MPI wrappers (path under the
    reference checkout) and a run as `ranks` processes over tests/emulation/mpi_emu.cpp; the wrapper's compare_buffers decides."""
    import tempfile
    srcs = []
    for p in cu_paths:
        dst = os.path.join(OUT, "dist_wrappers", tag, os.path.basename(p).replace(".cu", ".emu.cpp"))
        rewrite(p, dst)
        srcs.append(dst)
    exe = os.path.join(os.path.dirname(srcs[0]), "run")
    inc = sum((["-I", os.path.join(REF, d)] for d in include_dirs), [])
    r = _gxx(["-DWITH_MPI", "-I", os.path.join(HERE, "include"), "-I", os.path.join(COMPAT, "include"), "-I", os.path.join(REF, "3rdParty/Halide/src/runtime"),
              "-I", os.path.join(REF, "include"), *inc, os.path.join(REF, wrapper), *srcs,
              os.path.join(COMPAT, "compat_support.cpp"), os.path.join(REF, "src/tiramisu_utils.cpp"), os.path.join(REF, "src/tiramisu_debug.cpp"),
              runtime_object(), os.path.join(HERE, "mpi_emu.cpp"), "-o", exe, f"-L{ORACLE}", "-ltiramisu_oracle", f"-Wl,-rpath,{ORACLE}", "-lpthread", "-ldl"])
    if r.returncode:
        m = re.search(r"(error|undefined reference)[^\n]*", r.stderr)
        return "does not build: " + (m.group(0)[:300] if m else r.stderr[-300:])
    with tempfile.TemporaryDirectory(prefix="tb200_mpi_emu_") as box:
        procs = []
        for rk in range(ranks):
            env = dict(os.environ, RANK=str(rk), WORLD_SIZE=str(ranks), TB200_EMU_MPI_DIR=box)
            procs.append(subprocess.Popen([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env, cwd=box))
        bad = []
        for rk, pr in enumerate(procs):
            try:
                o, _ = pr.communicate(timeout=timeout)
            except subprocess.TimeoutExpired:
                pr.kill()
                o, _ = pr.communicate()
                o += "\n[timeout]"
            if pr.returncode != 0 or "Test failed" in o or "rror" in o or "[timeout]" in o or "ifferen" in o:
                tail = [ln for ln in o.splitlines() if ln.strip()]
                bad.append(f"rank {rk} rc={pr.returncode}: " + (tail[-1][:200] if tail else ""))
    return "pass" if not bad else "FAIL " + "; ".join(bad[:3])


def sweep(ids: list[str], workers: int = 8) -> dict:
    runtime_object()
    with cf.ThreadPoolExecutor(max_workers=workers) as ex:
        return dict(ex.map(reference_test, ids))


if __name__ == "__main__":
    expected = json.load(open(os.path.join(COMPAT, "reference_tests_expected.json")))
    ids = sys.argv[1:] or expected["pass"] + expected["host_emulation_only"]["tests"]
    res = sweep(ids)
    bad = {k: v for k, v in res.items() if v != "pass"}
    print(json.dumps(bad, indent=1))
    print(f"{len(res) - len(bad)}/{len(res)} pass in host emulation")
