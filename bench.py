#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on BASELINE.json's config, one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Headline workload (BASELINE.json configs[1]): heat2d + 3x3 blur stencil, fp32, 16384 x 16384 grid.
A step = one heat2d sweep (heat2d_tiramisu) + one fused 3x3 blur sweep (blurxy_f32_tiramisu) over
the grid.  metric = stencil GB/s = algorithmic bytes / time, algorithmic bytes per step =
2*N^2*4 (heat2d: read once, write once) + 4*(N^2 + (N-2)^2) (blur).
  value : inputs resident in HBM, kernels launched through the device-resident C-ABI entry.
  e2e   : same step through the host-buffer C-ABI call (pinned host memory; H2D + D2H inside).
The other BASELINE configs — sgemm 1024^3 (configs[0]), conv+ReLU+maxpool (configs[2]), the LSTM stack
(configs[3]) and the sgemm sweep (configs[4]) — ride in the same line under "configs", each with its own
value / roofline / cpu_baseline / e2e objects (tools/bench_configs.py), sharded over the ranks at N > 1.
Inputs (1 GiB each) are far larger than the 126 MB L2, so no explicit L2 flush is needed.
stdout carries exactly one JSON line: file descriptor 1 is pointed at stderr for the whole run (NCCL's
NCCL_DEBUG log, library banners) and the line is written to the saved descriptor at the end.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GRID = int(os.environ.get("TB200_BENCH_GRID", 16384))
METRIC = "stencil_GBps"
HEAT_NCU = "r02_heat2d_ncu_full.txt"          # committed `ncu --set full` summaries of the two sweep kernels
BLUR_NCU = "r02_blur_f32_ncu_full.txt"
UNIT = "GB/s"


_REAL_STDOUT = None


def claim_stdout() -> None:
    """Everything any library prints to fd 1 (NCCL_DEBUG=INFO goes to stdout) lands on stderr; the JSON line
    goes to the original stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict) -> None:
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def bench_config(world: int) -> dict:
    """The `config` object, identical in both arms (ours / --impl reference) for a given N."""
    return {"workload": f"heat2d + 3x3 blur fp32, {GRID}x{GRID} grid (BASELINE configs[1])",
            "sharding": f"rows/{world}",
            "l2": "inputs (1 GiB / n_gpus per rank) >> L2, no flush"}


def alg_bytes(n_rows: int, n_cols: int) -> dict:
    heat = 2 * n_rows * n_cols * 4
    blur = 4 * (n_rows * n_cols + (n_rows - 2) * (n_cols - 2))
    return {"heat2d": heat, "blur": blur, "step": heat + blur}


def measured_peaks() -> tuple[dict, str]:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def ncu_traffic(kernel_file: str):
    """dram__bytes_read.sum + dram__bytes_write.sum of the committed `ncu --set full` capture of the
    dominant kernel (profiles/), per launch, in bytes; None if the summary is missing."""
    path = os.path.join(ROOT, "profiles", kernel_file)
    if not os.path.exists(path):
        return None
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = 0.0
    for line in open(path):
        parts = line.split()
        if len(parts) == 3 and parts[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum") and parts[2] in unit:
            tot += float(parts[1]) * unit[parts[2]]
    return tot or None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md).  nvidia-smi needs
    a second or more to start on an 8-GPU box — longer than the timed region itself — so the sampler is
    started early, `mark()` brackets the span during which the GPU is kept under the benchmark's load, and
    only samples inside marked spans count."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int = 0):
        self.index, self.rows, self.proc, self.spans = index, [], None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.monotonic(), [c.strip() for c in line.split(",")]))

    def wait_started(self, timeout: float = 8.0) -> bool:
        t0 = time.monotonic()
        while self.proc and not self.rows and time.monotonic() - t0 < timeout:
            time.sleep(0.02)
        return bool(self.rows)

    def mark(self, t0: float, t1: float) -> None:
        self.spans.append((t0, t1))

    def samples_in_spans(self) -> int:
        return sum(1 for ts, _ in self.rows if any(a <= ts <= b for a, b in self.spans))

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self) -> dict:
        rows = [r for ts, r in self.rows if any(a <= ts <= b for a, b in self.spans)] if self.spans else [r for _, r in self.rows]
        sm = [float(r[0]) for r in rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v == "Active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "samples_total": len(self.rows)}


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args) -> None:
    """The reference's own CPU implementation of the path, timed on this box's host cores.  The
    reference x86 codegen cannot be built here (ISL+Halide+LLVM), so this is the oracle
    port (kind 'port') on ALL host cores (set explicitly: torchrun exports OMP_NUM_THREADS=1 to its
    ranks); each step is a bounded sample: the full heat2d+blur sweep on a GRID x GRID/4-row slab.
    Under torchrun rank 0 alone runs; the other ranks exit 0 without work."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    claim_stdout()
    from oracle import oracle
    cores = oracle.set_num_threads(None)
    rows = max(64, GRID // 4)
    rng = np.random.default_rng(1)
    a = rng.random((rows, GRID), dtype=np.float32)
    oh, ob = np.zeros_like(a), np.zeros((rows - 2, GRID - 2), np.float32)     # preallocated and touched
    for _ in range(max(1, min(args.warmup, 2))):
        oracle.heat2d(a, "cpu", oh); oracle.blur_f32(a, ob)
    ts = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        oracle.heat2d(a, "cpu", oh); oracle.blur_f32(a, ob)
        ts.append(time.perf_counter() - t0)
    dt = float(np.mean(ts))
    val = alg_bytes(rows, GRID)["step"] / dt / 1e9
    # one-core number on a smaller slab (BASELINE.md section 3 asks for both)
    oracle.set_num_threads(1)
    r1 = max(64, GRID // 64)
    a1 = a[:r1].copy()
    oracle.heat2d(a1, "cpu", oh[:r1]); oracle.blur_f32(a1, ob[:r1 - 2])
    t0 = time.perf_counter()
    oracle.heat2d(a1, "cpu", oh[:r1]); oracle.blur_f32(a1, ob[:r1 - 2])
    one = alg_bytes(r1, GRID)["step"] / (time.perf_counter() - t0) / 1e9
    oracle.set_num_threads(None)
    configs = None
    if not args.no_others:
        try:
            from tools import bench_configs
            configs = bench_configs.run_cpu_reference(args.steps)
        except Exception as e:  # context only
            configs = {"error": repr(e)}
    line = {"metric": METRIC, "value": val, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(args.gpus),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{rows} of {GRID} rows x {GRID} cols per step, {args.steps} steps (mean; median "
                                       f"{alg_bytes(rows, GRID)['step'] / float(np.median(ts)) / 1e9:.1f})",
                             "one_core": {"value": one, "unit": UNIT, "cores": 1, "sample": f"{r1} of {GRID} rows"},
                             "host_cores": oracle.host_cores(), "omp_num_threads_env": os.environ.get("OMP_NUM_THREADS")},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "configs": configs}
    emit(line)


# ------------------------------------------------------------------------------------------------ our arm
def time_steps(torch, stream, fn, steps, warmup, world):
    for _ in range(warmup):
        fn()
    stream.synchronize()
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t.item())
    return ms / steps


def per_kernel_ms(torch, stream, fn, iters=10):
    evs = []
    for _ in range(3):
        fn()
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); fn(); b.record(stream)
        evs.append((a, b))
    torch.cuda.synchronize()
    t = sorted(x.elapsed_time(y) for x, y in evs)
    return float(np.mean(t)), float(t[0])


def run_ours(args) -> None:
    # stdout stays the one JSON line; NCCL's log (NCCL_DEBUG, which the driver may set to check the
    # communicator's rank count) goes to stderr with everything else libraries print
    claim_stdout()
    if os.environ.get("TB200_NCCL_DEBUG"):
        os.environ["NCCL_DEBUG"] = os.environ["TB200_NCCL_DEBUG"]
    import torch
    from tiramisu_b200 import _lib, api

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = _lib.lib()
    stream = torch.cuda.Stream()
    api.set_stream(stream.cuda_stream)
    peaks, peak_src = measured_peaks()

    # ---- shard: rows [r0, r1) of the GRID x GRID grid (strong scaling, BASELINE configs[1]) -------
    # Each rank holds ONLY its row block plus halo rows: 1 row above (heat2d) and 2 rows below (blur
    # reads y..y+2, heat2d y+1); every step starts with the halo exchange over NCCL (one grouped
    # send/recv through the library's own communicator), as heat2d_dist / blurautodist do over MPI.
    from tiramisu_b200 import dist as tdist
    N = GRID
    r0, r1 = tdist.row_shard(N, world, rank)
    top = 1 if rank > 0 else 0
    bottom = 2 if rank < world - 1 else 0
    rows = r1 - r0
    b0, b1 = min(r0, N - 2), min(r1, N - 2)
    with torch.cuda.stream(stream):
        g = torch.Generator(device="cuda").manual_seed(1 + rank)
        d_loc = torch.zeros((top + rows + bottom, N), dtype=torch.float32, device="cuda")
        d_loc[top:top + rows] = torch.rand((rows, N), dtype=torch.float32, device="cuda", generator=g)
        d_heat = torch.empty((rows, N), dtype=torch.float32, device="cuda")
        d_blur = torch.empty((max(b1 - b0, 1), N - 2), dtype=torch.float32, device="cuda")
    stream.synchronize()
    row_b = N * 4
    # virtual row-0 pointers: the launchers index rows globally (row_begin/row_end are global rows)
    ip = d_loc.data_ptr() - (r0 - top) * row_b
    hp = d_heat.data_ptr() - r0 * row_b
    bp = d_blur.data_ptr() - b0 * (N - 2) * 4
    own = d_loc.data_ptr() + top * row_b                       # first owned row

    exchange = None
    if world > 1:
        import ctypes as C
        tdist.comm_init_from_torch()
        up, dn = rank - 1, rank + 1
        sp = (C.c_int32 * 2)(up if rank > 0 else -1, dn if rank < world - 1 else -1)
        sb = (C.c_void_p * 2)(own, own + (rows - 1) * row_b)                    # my first 2 rows up, my last row down
        sn = (C.c_uint64 * 2)(2 * row_b, row_b)
        rp = (C.c_int32 * 2)(up if rank > 0 else -1, dn if rank < world - 1 else -1)
        rb = (C.c_void_p * 2)(d_loc.data_ptr(), own + rows * row_b)             # top halo (1 row), bottom halo (2 rows)
        rn = (C.c_uint64 * 2)(row_b, 2 * row_b)

        def exchange():
            _lib.check(L.tiramisu_b200_comm_exchange(2, sp, sb, sn, rp, rb, rn))

    def heat(ya=r0, yb=r1):
        if yb > ya:
            _lib.check(L.tiramisu_b200_heat2d(ip, hp, N, N, 0, ya, yb))

    def blur(ya=b0, yb=b1):
        if yb > ya:
            _lib.check(L.tiramisu_b200_blur_f32(ip, bp, N, N, ya, yb))

    if world == 1:
        def step():
            heat(); blur()
    else:
        # Two streams per step: the comm stream does the halo exchange and then the few rows that read
        # halo rows; the launch stream does the interior rows meanwhile and joins at the end.
        comm_stream = torch.cuda.Stream()
        ev_free, ev_done = torch.cuda.Event(), torch.cuda.Event()
        hi0, hi1 = r0 + top, r1 - (1 if bottom else 0)          # heat2d rows that need no halo
        bi1 = max(b0, b1 - (2 if bottom else 0))                # blur rows [b0, bi1) need no halo

        def step():
            ev_free.record(stream)                              # previous step's readers/writers are done
            comm_stream.wait_event(ev_free)
            api.set_stream(comm_stream.cuda_stream)
            exchange()
            heat(r0, hi0); heat(hi1, r1); blur(bi1, b1)
            api.set_stream(stream.cuda_stream)
            ev_done.record(comm_stream)
            heat(hi0, hi1); blur(b0, bi1)
            stream.wait_event(ev_done)

    if world > 1:
        # halo sanity: what I received must be what my neighbours own (checksums over the torch group)
        step(); stream.synchronize()
        mine = torch.stack([d_loc[top:top + 2].double().sum(), d_loc[top + rows - 1].double().sum(),
                            d_loc[0:top].double().sum(), d_loc[top + rows:].double().sum()])
        allv = [torch.zeros_like(mine) for _ in range(world)]
        torch.distributed.all_gather(allv, mine)
        if rank > 0:
            assert abs(float(allv[rank - 1][1] - mine[2])) < 1e-6, "top halo != neighbour's last row"
        if rank < world - 1:
            assert abs(float(allv[rank + 1][0] - mine[3])) < 1e-6, "bottom halo != neighbour's first two rows"

    # ---- default at N > 1: no exchange at all — the block lives in symmetric memory mapped into every
    # process (CUDA IPC over NVLink) and each sweep reads the neighbours' boundary rows in place
    # (stencil.cu: Halo).  One launch per stencil per step, like N = 1.  TB200_HALO=nccl keeps the
    # NCCL exchange above as the timed step; its time is reported next to the peer one either way.
    halo_mode = os.environ.get("TB200_HALO", "peer") if world > 1 else "none"
    sym = ref_heat = ref_blur = None
    nccl_step_ms = None
    halo_fallback = None
    if halo_mode == "peer":
        import ctypes as C

        class _Mem:                                           # torch view of library-owned device memory
            def __init__(self, ptr, shape):
                self.__cuda_array_interface__ = {"shape": shape, "typestr": "<f4", "data": (ptr, False), "version": 2}

        rows_of = lambda r: tdist.row_shard(N, world, r)      # noqa: E731
        # symmetric memory needs CUDA IPC between the ranks' processes; where that is not available every
        # rank falls back (together) to the NCCL exchange step above
        try:
            sym = tdist.SymmetricRows(max(rows_of(r)[1] - rows_of(r)[0] for r in range(world)), N, 1)
        except Exception as e:                                # noqa: BLE001
            sym, halo_fallback = None, repr(e)
        agree = torch.tensor([0 if sym is None else 1], device="cuda")
        torch.distributed.all_reduce(agree, op=torch.distributed.ReduceOp.MIN)
        if int(agree.item()) == 0:
            halo_mode = "nccl"
            halo_fallback = halo_fallback or "another rank could not map symmetric memory"
    if halo_mode == "peer":
        with torch.cuda.stream(stream):
            torch.as_tensor(_Mem(sym.buffer(0), (rows, N)), device="cuda").copy_(d_loc[top:top + rows])
        stream.synchronize()
        torch.distributed.barrier()
        sp_ip = sym.buffer(0) - r0 * row_b
        halo = sym.halo(0, r0, r1, rows_of)
        ref_heat, ref_blur = d_heat.clone(), d_blur.clone()   # what the NCCL-exchange step just produced
        nccl_step = step
        nccl_step_ms = time_steps(torch, stream, nccl_step, max(5, args.steps // 2), 3, world)

        def heat(ya=r0, yb=r1):                               # noqa: F811
            if yb > ya:
                _lib.check(L.tiramisu_b200_heat2d_peer(sp_ip, hp, N, N, 0, ya, yb, C.byref(halo)))

        def blur(ya=b0, yb=b1):                               # noqa: F811
            if yb > ya:
                _lib.check(L.tiramisu_b200_blur_f32_peer(sp_ip, bp, N, N, ya, yb, C.byref(halo)))

        # the two stencils are independent (same input, different outputs): issue them on two streams so
        # that the tail of one fills with CTAs of the other (at 1/8 of the grid a kernel lasts ~40 us)
        side = torch.cuda.Stream()
        ev_fork, ev_join = torch.cuda.Event(), torch.cuda.Event()

        def step():                                           # noqa: F811
            ev_fork.record(stream)
            side.wait_event(ev_fork)
            heat()
            api.set_stream(side.cuda_stream)
            blur()
            api.set_stream(stream.cuda_stream)
            ev_join.record(side)
            stream.wait_event(ev_join)

        with torch.cuda.stream(stream):
            d_heat.fill_(float("nan")); d_blur.fill_(float("nan"))
        step(); stream.synchronize()
        assert torch.equal(d_heat, ref_heat), "peer-halo heat2d differs from the NCCL-exchange result"
        assert torch.equal(d_blur[: b1 - b0], ref_blur[: b1 - b0]), "peer-halo blur differs from the NCCL-exchange result"

    launches0 = api.kernel_launches()
    with ClockSampler(local) as clk:
        clk.wait_started()             # nvidia-smi takes a second or more to come up on a multi-GPU box
        t_load0 = time.monotonic()
        ms = time_steps(torch, stream, step, args.steps, args.warmup, world)
        launches = (api.kernel_launches() - launches0) * world // (args.steps + args.warmup) * args.steps
        heat_ms, heat_best = per_kernel_ms(torch, stream, heat)
        blur_ms, blur_best = per_kernel_ms(torch, stream, blur)
        # keep the GPU under the same load for ~0.8 s so that nvidia-smi (50 ms period) samples it; the step count
        # comes from the all-reduced `ms`, so every rank runs the same number (the NCCL-exchange step is collective)
        n_keep = min(20000, int(800.0 / max(ms, 1e-3)) + 1)
        for i in range(n_keep):
            step()
            if i % 64 == 63:
                stream.synchronize()
                clk.mark(t_load0, time.monotonic())
        stream.synchronize()
        clk.mark(t_load0, time.monotonic())
        if world > 1:
            torch.distributed.barrier()
    bytes_step = alg_bytes(N, N)["step"]
    value = bytes_step / (ms * 1e-3) / 1e9

    heat_bytes = 2 * (r1 - r0) * N * 4
    blur_bytes = 4 * ((b1 - b0) * (N - 2) + (b1 - b0 + 2) * N)
    peak = float(peaks["hbm_gbs"])
    heat_gbs, blur_gbs = heat_bytes / (heat_ms * 1e-3) / 1e9, blur_bytes / (blur_ms * 1e-3) / 1e9
    kernels = {"heat2d": {"kernel": "tiramisu_sm100_heat2d_v4", "ms": heat_ms, "best_ms": heat_best, "GBps": heat_gbs,
                          "frac": heat_gbs / peak, "algorithmic_bytes": heat_bytes,
                          "traffic": ncu_traffic(HEAT_NCU) if world == 1 else None, "traffic_source": f"profiles/{HEAT_NCU}"},
               "blur_f32": {"kernel": "tiramisu_sm100_blur_f32_v4", "ms": blur_ms, "best_ms": blur_best, "GBps": blur_gbs,
                            "frac": blur_gbs / peak, "algorithmic_bytes": blur_bytes,
                            "traffic": ncu_traffic(BLUR_NCU) if world == 1 else None, "traffic_source": f"profiles/{BLUR_NCU}"}}
    # the roofline object describes the DOMINANT kernel of the step = the one with the longer launch
    dom = "blur_f32" if blur_ms >= heat_ms else "heat2d"
    roof = {"kernel": kernels[dom]["kernel"], "bound": "hbm", "achieved": kernels[dom]["GBps"], "peak": peak, "unit": "GB/s",
            "frac": kernels[dom]["frac"], "peak_source": f"{peak_src} (MEASURED_PEAKS.json hbm_gbs, burst copy)",
            "traffic": kernels[dom]["traffic"],
            "traffic_source": kernels[dom]["traffic_source"] + " (ncu --set full, 16384^2, one launch)",
            "algorithmic_bytes": kernels[dom]["algorithmic_bytes"], "ms": kernels[dom]["ms"],
            "share_of_step": kernels[dom]["ms"] / (heat_ms + blur_ms),
            "step": {"achieved": (heat_bytes + blur_bytes) / ((heat_ms + blur_ms) * 1e-3) / 1e9,
                     "frac": (heat_bytes + blur_bytes) / ((heat_ms + blur_ms) * 1e-3) / 1e9 / peak}}
    if exchange:
        ex_ms, _ = per_kernel_ms(torch, stream, exchange)
        kernels["halo_exchange_nccl"] = {"ms": ex_ms, "bytes_per_rank": 3 * row_b * (2 if 0 < rank < world - 1 else 1),
                                         "in_timed_step": halo_mode == "nccl"}
    if nccl_step_ms is not None:
        kernels["step_with_nccl_exchange_ms"] = nccl_step_ms
    d_in = d_loc

    # ---- e2e: the same step through the host-buffer generated-function ABI (pinned host memory; the
    # H2D copy of the inputs and the D2H copy of the results are inside the timed region).  At N > 1
    # every rank runs its own row block (+ halo rows, already on the host) as one grid.
    LR = d_in.shape[0]
    h_in = torch.empty((LR, N), dtype=torch.float32).pin_memory()
    h_in.copy_(d_in)
    h_heat = torch.empty((LR, N), dtype=torch.float32).pin_memory()
    h_blur = torch.empty((LR - 2, N - 2), dtype=torch.float32).pin_memory()

    def e2e_step():
        api.heat2d_tiramisu(h_in, h_heat)
        api.blurxy_f32_tiramisu(h_in, h_blur)

    e2e_step()
    k = max(1, min(args.steps, 5))
    if world > 1:
        torch.distributed.barrier()
    t0 = time.perf_counter()
    for _ in range(k):
        e2e_step()
    dt = (time.perf_counter() - t0) / k
    e2e_bytes = float(alg_bytes(LR, N)["step"])
    h2d, d2h = 2.0 * LR * N * 4, float(LR * N * 4 + (LR - 2) * (N - 2) * 4)
    if world > 1:
        t = torch.tensor([dt], device="cuda", dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        dt = float(t.item())
        tot = torch.tensor([e2e_bytes, h2d, d2h], device="cuda", dtype=torch.float64)
        torch.distributed.all_reduce(tot)
        e2e_bytes, h2d, d2h = (float(v) for v in tot.tolist())
    e2e = {"value": e2e_bytes / dt / 1e9, "unit": UNIT, "ms_per_step": dt * 1e3, "steps": k,
           "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "how": "heat2d_tiramisu + blurxy_f32_tiramisu on pinned host buffers (row-chunk pipelined H2D | kernel | D2H)"}

    # ---- CPU baseline (rank 0, N=1): the oracle port on a bounded sample ---------------------------
    cpu = None
    if world == 1 and not args.no_cpu:
        from oracle import oracle
        cores = oracle.set_num_threads(None)
        rows = max(64, N // 4)
        a = np.random.default_rng(1).random((rows, N), dtype=np.float32)
        oh, ob = np.zeros_like(a), np.zeros((rows - 2, N - 2), np.float32)    # preallocated and touched
        oracle.heat2d(a, "cpu", oh); oracle.blur_f32(a, ob)
        reps = 3
        t0 = time.perf_counter()
        for _ in range(reps):
            oracle.heat2d(a, "cpu", oh); oracle.blur_f32(a, ob)
        dt = (time.perf_counter() - t0) / reps
        oracle.set_num_threads(1)
        r1c = max(64, N // 64)
        a1 = a[:r1c].copy()
        oracle.heat2d(a1, "cpu", oh[:r1c]); oracle.blur_f32(a1, ob[:r1c - 2])
        t0 = time.perf_counter()
        oracle.heat2d(a1, "cpu", oh[:r1c]); oracle.blur_f32(a1, ob[:r1c - 2])
        one = alg_bytes(r1c, N)["step"] / (time.perf_counter() - t0) / 1e9
        oracle.set_num_threads(None)
        cpu = {"value": alg_bytes(rows, N)["step"] / dt / 1e9, "unit": UNIT, "cores": cores,
               "kind": "port", "sample": f"{rows} of {N} rows x {N} cols, {reps} reps",
               "one_core": {"value": one, "unit": UNIT, "cores": 1, "sample": f"{r1c} of {N} rows"}}
        del a, a1, oh, ob

    # ---- the other BASELINE configs (sgemm 1024^3, conv block, LSTM, sgemm sweep): every rank takes part
    configs = None
    if not args.no_others:
        del d_loc, d_in, d_heat, d_blur, h_in, h_heat, h_blur, ref_heat, ref_blur
        if halo_mode == "peer":
            stream.synchronize()
            sym.free()
            sym = None
        torch.cuda.empty_cache()
        from tools import bench_configs
        try:
            configs = bench_configs.run(torch, stream, peaks, world, rank, cpu=not args.no_cpu,
                                        only=args.only.split(",") if args.only else None, quick=args.quick)
        except Exception as e:  # context only; never fail the headline (a collective failure re-raises inside)
            configs = {"error": repr(e)}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": bench_config(world),
                "halo": {"peer": "halo rows read in place from the neighbours' HBM over NVLink peer mappings inside the sweep kernels (no exchange step)",
                         "nccl": "halo exchange over NCCL each step, overlapped with the interior rows",
                         "none": "single GPU"}.get(halo_mode, halo_mode),
                **({"halo_fallback": halo_fallback} if halo_fallback else {}),
                "roofline": roof, "kernels": kernels, "cpu_baseline": cpu, "e2e": e2e,
                "gpu_launches": int(launches), "clocks": clk.summary(), "configs": configs}
        emit(line)
    if world > 1:
        if halo_mode == "peer" and sym is not None:
            stream.synchronize()
            sym.free()
        tdist.comm_destroy()
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-others", action="store_true")
    ap.add_argument("--only", default="", help="comma-separated prefixes of the entries of `configs` to run")
    ap.add_argument("--quick", action="store_true", help="shorter sgemm sweep (development runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
