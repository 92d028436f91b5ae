"""BASELINE.json configs other than the headline stencil step — sgemm 1024^3 (configs[0]), the
conv+ReLU+maxpool block (configs[2]), the LSTM stack (configs[3]) and the square sgemm sweep
(configs[4]) — each measured like the headline line of bench.py:

  value        device-resident, CUDA events on the launch stream, max over ranks; at N > 1 the work is
               sharded the way SURVEY.md §8e says (sgemm: A and C by rows, B broadcast from rank 0 every
               step; conv / LSTM: batch shards, no per-step collective) and `value` is the whole job
  roofline     the dominant kernel against the measured peak (tensor: TF32 = bf16 / 2 of MEASURED_PEAKS.json)
  e2e          the same work through the reference-facing generated-function ABI (sgemm_tiramisu /
               conv_relu_maxpool_block / lstm) on pinned HOST buffers, copies inside the timed region
  cpu_baseline the oracle port on this box's host cores (all cores and one core), bounded sample — rank 0, N = 1

Imported by bench.py on every rank (the sharded legs are collective).  Nothing here is the headline.
"""
from __future__ import annotations

import time

import numpy as np


# --------------------------------------------------------------------------------------------- helpers
def _tf32_peak(peaks) -> float:
    return float(peaks.get("tf32_tflops", float(peaks["bf16_tflops"]) / 2.0))


class Ctx:
    def __init__(self, torch, stream, peaks, world, rank, cpu: bool, quick: bool = False):
        self.torch, self.stream, self.peaks = torch, stream, peaks
        self.world, self.rank, self.cpu, self.quick = world, rank, cpu and world == 1 and rank == 0, quick
        self.dist = torch.distributed if world > 1 else None

    def max_over_ranks(self, v: float) -> float:
        if self.world == 1:
            return v
        t = self.torch.tensor([v], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, v: float) -> float:
        if self.world == 1:
            return v
        t = self.torch.tensor([v], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t)
        return float(t.item())

    def time_device(self, fn, iters: int, warm: int = 3, extra_streams=()) -> float:
        """ms per call: CUDA events on the launch stream, barrier + synchronize on both sides, max over ranks."""
        torch = self.torch
        for _ in range(warm):
            fn()
        self.stream.synchronize()
        for s in extra_streams:
            s.synchronize()
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        for _ in range(iters):
            fn()
        e1.record(self.stream)
        torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1) / iters)

    def time_host(self, fn, iters: int) -> float:
        """seconds per call of a host-buffer ABI call (it synchronises before returning), max over ranks."""
        fn()
        if self.world > 1:
            self.dist.barrier()
        t0 = time.perf_counter()
        for _ in range(iters):
            fn()
        return self.max_over_ranks((time.perf_counter() - t0) / iters)


def _cpu_time(fn, min_reps=2, budget_s=1.5):
    fn()                                            # first touch / thread start-up
    ts = []
    t_all = time.perf_counter()
    while len(ts) < min_reps or (time.perf_counter() - t_all < budget_s and len(ts) < 5):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return float(np.median(ts))                     # benchmarks/benchmarks.h:12 — median of NB_TESTS runs


def _cpu_entry(flop_all, fn_all, what_all, flop_one, fn_one, what_one):
    """All-cores and one-core GFLOP/s of the oracle port on bounded samples."""
    from oracle import oracle
    cores = oracle.set_num_threads(None)
    t_all = _cpu_time(fn_all)
    oracle.set_num_threads(1)
    t_one = _cpu_time(fn_one, min_reps=1, budget_s=0.5)
    oracle.set_num_threads(None)
    return {"value": flop_all / t_all / 1e9, "unit": "GFLOP/s", "cores": cores, "kind": "port", "sample": what_all,
            "one_core": {"value": flop_one / t_one / 1e9, "unit": "GFLOP/s", "cores": 1, "sample": what_one}}


def _roof_tensor(ctx, kernel, flop, ms, traffic_file=None):
    peak = _tf32_peak(ctx.peaks)
    ach = flop / (ms * 1e-3) / 1e12
    return {"kernel": kernel, "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
            "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst) / 2 = TF32 dense", "traffic": _traffic(traffic_file),
            "traffic_source": traffic_file, "algorithmic_flop": flop, "ms": ms}


def _traffic(fname):
    if not fname:
        return None
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", fname)
    if not os.path.exists(path):
        return None
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = 0.0
    for line in open(path):
        parts = line.split()
        if len(parts) == 3 and parts[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum") and parts[2] in unit:
            tot += float(parts[1]) * unit[parts[2]]
    return tot or None


def _pinned(torch, shape, dtype=None):
    return torch.empty(shape, dtype=dtype or torch.float32).pin_memory()


# --------------------------------------------------------------------------------------------- sgemm
def sgemm(ctx: Ctx, n: int, e2e: bool = True) -> dict:
    """C[n x n] = 3*A*B + 2*C (sgemm_wrapper.cpp:24 uses a = b = 3; gpu/wrapper.cpp 3 and 2), fp32 operands run as
    TF32 on tcgen05.  N > 1: rows of A and C sharded, B (n x n) broadcast from rank 0 inside every timed step, in
    K panels so that panel p+1 travels while the GEMM on panel p runs."""
    torch = ctx.torch
    from tiramisu_b200 import _lib, api, dist as tdist
    L = _lib.lib()
    world, rank = ctx.world, ctx.rank
    r0, r1 = tdist.row_shard(n, world, rank)
    rows = r1 - r0
    main_s = ctx.stream
    with torch.cuda.stream(main_s):
        g = torch.Generator(device="cuda").manual_seed(5 + rank)
        A = torch.rand((rows, n), device="cuda", generator=g) * 2 - 1
        C = torch.zeros((rows, n), device="cuda")
        B = (torch.rand((n, n), device="cuda", generator=g) * 2 - 1) if rank == 0 else torch.empty((n, n), device="cuda")
    main_s.synchronize()
    iters = 20 if n <= 4096 else (5 if n <= 8192 else 3)
    flop = 2.0 * n ** 3
    out = {"metric": "sgemm_GFLOPs", "unit": "GFLOP/s", "n": n, "dtype": "tf32 operands, fp32 accumulate",
           "scaling": "strong", "n_gpus": world}
    if world == 1:
        # L2 hygiene: rotate through enough operand sets to exceed the 126 MB L2 for the small sizes
        nsets = max(1, int(np.ceil(160e6 / (3 * n * n * 4)))) if n <= 2048 else 1
        sets = [(A, B, C)] + [(A.clone(), B.clone(), C.clone()) for _ in range(nsets - 1)]
        idx = [0]

        def step():
            a, b, c = sets[idx[0] % len(sets)]
            idx[0] += 1
            api.cublas_sgemm(a.data_ptr(), b.data_ptr(), c.data_ptr(), n, n, n, 3.0, 2.0)

        ms = ctx.time_device(step, iters)
        out["l2"] = f"{nsets} operand sets rotated (> L2)" if nsets > 1 else "operands > L2, no flush"
        if n <= 2048:
            # A 1024^3 product lasts a few microseconds — less than one host-side launch through ctypes + the CUDA
            # runtime — so back-to-back stream launches measure the host.  The same launches captured once in a CUDA
            # graph (one pass over all rotating operand sets) and replayed give the device-side period; both are reported,
            # `value` and the roofline use the graph number when it is the smaller one.
            try:
                out["ms_stream_launched"] = ms
                g = torch.cuda.CUDAGraph()
                main_s.synchronize()
                with torch.cuda.graph(g, stream=main_s):
                    for _ in range(len(sets)):
                        step()
                reps = max(3, 40 // len(sets))
                with torch.cuda.stream(main_s):
                    for _ in range(2):
                        g.replay()
                    main_s.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(main_s)
                    for _ in range(reps):
                        g.replay()
                    e1.record(main_s)
                torch.cuda.synchronize()
                ms_graph = e0.elapsed_time(e1) / (reps * len(sets))
                out["ms_graph_replay"] = ms_graph
                out["launch"] = "CUDA graph replay of one pass over the operand sets" if ms_graph < ms else "stream launches"
                ms = min(ms, ms_graph)
                del g
                if n == 1024:
                    # the launch floor: an EMPTY kernel of the GEMM's launch shape (128 CTAs x 192 threads, 2-CTA clusters,
                    # 198 KB of shared memory each), graph-replayed the same way
                    g0 = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g0, stream=main_s):
                        for _ in range(16):
                            _lib.check(L.tiramisu_b200_null_launch(128, 192, 197888, 2))
                    with torch.cuda.stream(main_s):
                        g0.replay()
                        main_s.synchronize()
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record(main_s)
                        for _ in range(10):
                            g0.replay()
                        e1.record(main_s)
                    torch.cuda.synchronize()
                    out["launch_floor_ms"] = e0.elapsed_time(e1) / 160
                    del g0
            except Exception as e:                        # context only: keep the stream-launched number
                out["graph_error"] = repr(e)
        del sets
    else:
        # one call per step moves rank 0's B through symmetric staging memory (DMA-engine scatter + all-gather of column
        # panels over NVLink, no SM kernels) and runs the GEMM, gated inside the kernel on the panels' arrival words
        symB = tdist.SymmetricB(n, n)
        panels = symB.sgemm_panels()
        Bp = B.data_ptr() if rank == 0 else None

        def step():
            symB.sgemm(Bp, A.data_ptr(), C.data_ptr(), rows, 3.0, 2.0, 0, panels)

        ms = ctx.time_device(step, iters)
        ms_gemm = ctx.time_device(lambda: api.cublas_sgemm(A.data_ptr(), symB.data_ptr(), C.data_ptr(), rows, n, n, 3.0, 2.0), iters)
        ms_bcast = ctx.time_device(lambda: symB.sgemm(Bp, A.data_ptr(), C.data_ptr(), 0, 3.0, 2.0, 0, panels), iters)
        out.update({"sharding": f"A, C rows/{world}; B ({n * n * 4} B) travels from rank 0 every step in {panels} column panel(s), staged "
                                "panel-major, as peer copies by the DMA engines (scatter + all-gather) while ONE persistent GEMM launch "
                                "per rank consumes the panels in order, gated per tile in the TMA producer warp",
                    "ms_gemm_only": ms_gemm, "GFLOPs_gemm_only": flop / (ms_gemm * 1e-3) / 1e9, "ms_bcast_only": ms_bcast,
                    "bcast_GBps_per_rank": n * n * 4 / (ms_bcast * 1e-3) / 1e9, "bcast_bytes_per_rank_per_step": n * n * 4})
        # the round-1 form for comparison: ncclBroadcast of K panels on a second stream + events
        comm_s = torch.cuda.Stream()
        np_old = 4 if n >= 4096 else 1
        kw = n // np_old
        evs = [torch.cuda.Event() for _ in range(np_old)]
        ev_free = torch.cuda.Event()

        def step_nccl():
            ev_free.record(main_s)                       # the previous step's GEMMs no longer read B
            comm_s.wait_event(ev_free)
            api.set_stream(comm_s.cuda_stream)
            for p in range(np_old):
                _lib.check(L.tiramisu_b200_comm_broadcast(B.data_ptr() + p * kw * n * 4, kw * n * 4, 0))
                evs[p].record(comm_s)
            api.set_stream(main_s.cuda_stream)
            for p in range(np_old):
                main_s.wait_event(evs[p])
                api.cublas_sgemm(A.data_ptr(), B.data_ptr(), C.data_ptr(), rows, n, kw, 3.0, 2.0 if p == 0 else 1.0,
                                 ldA=n, ldB=n, ldC=n, offA=p * kw, offB=p * kw * n)

        ms_nccl = ctx.time_device(step_nccl, iters, extra_streams=(comm_s,))
        out["ms_ncclBroadcast_form"] = ms_nccl
        out["ms_dma_fused_form"] = ms
        # both forms are this library's code; the line's value is the faster one on this box and says which
        # (r02 runs 12-17: the DMA form wins at 2 GPUs, at 8 the fine-grained peer copies between IPC-mapped blocks
        # only reach ~330 GB/s per rank and ncclBroadcast + events is ahead)
        out["form"] = "dma_fused" if ms <= ms_nccl else "ncclBroadcast + stream events"
        ms = min(ms, ms_nccl)
        symB.free()
    out.update({"value": flop / (ms * 1e-3) / 1e9, "ms": ms,
                "roofline": _roof_tensor(ctx, "tiramisu_sm100_gemm_tf32*", flop, ms,
                                         "r02_sgemm1024_ncu_full.txt" if n == 1024 else None)})
    out["roofline"]["hbm_bound_ms"] = 16.0 * n * n / (float(ctx.peaks["hbm_gbs"]) * 1e9) * 1e3
    out["roofline"]["tensor_bound_ms"] = flop / (_tf32_peak(ctx.peaks) * 1e12) * 1e3
    if "launch_floor_ms" in out:
        # at this size flop, byte and launch bounds are co-limiting — report the fraction of the largest
        bound = max(out["roofline"]["hbm_bound_ms"], out["roofline"]["tensor_bound_ms"], out["launch_floor_ms"])
        out["roofline"].update({"launch_floor_ms": out["launch_floor_ms"], "max_bound_ms": bound, "frac_of_max_bound": bound / ms})
    if world == 1 and n <= 4096 and not ctx.quick:
        try:
            out["cublas_tf32_GFLOPs_same_run"] = _cublas_tf32(torch, n, iters) * 1e3
        except Exception as e:                       # context only
            out["cublas_tf32_GFLOPs_same_run"] = repr(e)

    if e2e:
        # the reference-facing call: sgemm_tiramisu(SIZES, alpha, beta, A, B, C) on pinned host buffers
        hA, hB, hC = _pinned(torch, (rows, n)), _pinned(torch, (n, n)), _pinned(torch, (rows, n))
        hA.copy_(A); hC.zero_()
        if rank == 0:
            hB.copy_(B)
        else:
            hB.zero_()
        sizes = np.array([rows, n, n], np.int32)
        al, be = np.array([3.0], np.float32), np.array([2.0], np.float32)
        dt = ctx.time_host(lambda: api.sgemm_tiramisu(sizes, al, be, hA, hB, hC), 3 if n <= 4096 else 2)
        h2d = ctx.sum_over_ranks(float(2 * rows * n * 4 + n * n * 4))
        d2h = ctx.sum_over_ranks(float(rows * n * 4))
        out["e2e"] = {"value": flop / dt / 1e9, "unit": "GFLOP/s", "ms": dt * 1e3, "h2d_bytes_per_step": int(h2d),
                      "d2h_bytes_per_step": int(d2h), "how": "sgemm_tiramisu(SIZES, alpha, beta, A, B, C) on pinned host buffers"
                      + ("" if world == 1 else "; every rank uploads its row shard and B")}
        del hA, hB, hC

    if ctx.cpu:
        from oracle import oracle
        rs = min(n, max(64, (int(6e9 / (2.0 * n * n)) // 4) * 4))          # row slab: ~6 GFLOP of CPU work per rep
        r1c = min(rs, 64)
        hA = np.random.default_rng(5).uniform(-1, 1, (rs, n)).astype(np.float32)
        hB = np.random.default_rng(6).uniform(-1, 1, (n, n)).astype(np.float32)
        hC = np.zeros((rs, n), np.float32)
        out["cpu_baseline"] = _cpu_entry(
            2.0 * rs * n * n, lambda: oracle.sgemm_sched(3.0, 2.0, hA, hB, hC), f"{rs} of {n} rows of C (oracle.sgemm_sched: tiled, 8-wide lanes)",
            2.0 * r1c * n * n, lambda: oracle.sgemm_sched(3.0, 2.0, hA[:r1c], hB, hC[:r1c]), f"{r1c} of {n} rows of C")
    del A, B, C
    torch.cuda.empty_cache()
    return out


def _cublas_tf32(torch, n, iters):
    """Library reference point (context only, never on the product path): cuBLAS TF32 via torch.matmul."""
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        nsets = max(1, int(np.ceil(160e6 / (3 * n * n * 4)))) if n <= 2048 else 1
        sets = [(torch.rand((n, n), device="cuda") * 2 - 1, torch.rand((n, n), device="cuda") * 2 - 1) for _ in range(nsets)]
        for a, b in sets[:3]:
            a @ b
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            a, b = sets[i % nsets]
            a @ b
        e1.record()
        torch.cuda.synchronize()
        return 2.0 * n ** 3 / (e0.elapsed_time(e1) / iters * 1e-3) / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


# --------------------------------------------------------------------------------------------- conv block
def conv(ctx: Ctx) -> dict:
    """BASELINE configs[2]: conv+ReLU+maxpool fp32 N=32, 64->64, 3x3, 224x224; batch-sharded at N > 1."""
    torch = ctx.torch
    from tiramisu_b200 import _lib, api, dist as tdist
    L = _lib.lib()
    B, N, FIn, FOut = 32, 224, 64, 64
    b0, b1 = tdist.batch_shard(B, ctx.world, ctx.rank)
    nb = b1 - b0
    with torch.cuda.stream(ctx.stream):
        g = torch.Generator(device="cuda").manual_seed(3 + ctx.rank)
        x = torch.rand((max(nb, 1), N + 2, N + 2, FIn), device="cuda", generator=g) * 2 - 1
        gw = torch.Generator(device="cuda").manual_seed(100 + ctx.rank)
        w = torch.rand((FOut // 8, 3, 3, FIn, 8), device="cuda", generator=gw) * 2 - 1
        b = torch.rand((FOut,), device="cuda", generator=gw) * 2 - 1
        o = torch.empty((max(nb, 1), FOut // 8, N // 2, N // 2, 8), device="cuda")
    ctx.stream.synchronize()
    if ctx.world > 1:                                   # weights: rank 0's values everywhere, once (147 kB)
        for t in (w, b):
            _lib.check(L.tiramisu_b200_comm_broadcast(t.data_ptr(), t.numel() * 4, 0))

    def step():
        if nb:
            _lib.check(L.tiramisu_b200_conv_relu_maxpool(x.data_ptr(), w.data_ptr(), b.data_ptr(), o.data_ptr(), nb, N, FIn, FOut))

    ms = ctx.time_device(step, 10)
    flop = 2.0 * B * N * N * FOut * FIn * 9
    out = {"metric": "conv_relu_maxpool_GFLOPs", "unit": "GFLOP/s", "value": flop / (ms * 1e-3) / 1e9, "ms": ms,
           "dtype": "tf32 operands, fp32 accumulate", "scaling": "strong", "n_gpus": ctx.world,
           "sharding": f"batch {B}/{ctx.world} images per rank, weights broadcast once, no per-step collective",
           "l2": "input shard (418 MB / n_gpus) > L2 up to 2 GPUs; no flush",
           "roofline": _roof_tensor(ctx, "tiramisu_sm100_conv_block_tf32", flop, ms, "r01_conv_pair_ncu_full.txt" if ctx.world == 1 else None),
           "bytes_GBps": B * ((N + 2) * (N + 2) * FIn + FOut * (N // 2) * (N // 2)) * 4 / (ms * 1e-3) / 1e9}

    hx, ho = _pinned(torch, tuple(x.shape)), _pinned(torch, tuple(o.shape))
    hx.copy_(x)
    hw, hb = w.cpu().numpy(), b.cpu().numpy()
    if nb:
        dt = ctx.time_host(lambda: api.conv_relu_maxpool_block(hx, hw, hb, ho), 3)
    else:
        dt = ctx.time_host(lambda: None, 3)
    out["e2e"] = {"value": flop / dt / 1e9, "unit": "GFLOP/s", "ms": dt * 1e3,
                  "h2d_bytes_per_step": int(ctx.sum_over_ranks(float(nb * (N + 2) * (N + 2) * FIn * 4 + hw.nbytes + hb.nbytes))),
                  "d2h_bytes_per_step": int(ctx.sum_over_ranks(float(nb * FOut * (N // 2) * (N // 2) * 4))),
                  "how": "conv_relu_maxpool_block(input, filter, bias, output) on pinned host buffers"}
    if ctx.cpu:
        from oracle import oracle
        rng = np.random.default_rng(3)
        cx = rng.uniform(-1, 1, (2, N + 2, N + 2, FIn)).astype(np.float32)
        per_img = 2.0 * N * N * FOut * FIn * 9
        crop = cx[:1, :58, :58].copy()
        out["cpu_baseline"] = _cpu_entry(2 * per_img, lambda: oracle.conv_relu_maxpool(cx, hw, hb), f"2 of {B} images",
                                         2.0 * 56 * 56 * FOut * FIn * 9, lambda: oracle.conv_relu_maxpool(crop, hw, hb),
                                         "one 56x56 crop (1/16 image)")
    del x, o, hx, ho
    torch.cuda.empty_cache()
    return out


# --------------------------------------------------------------------------------------------- LSTM
def lstm(ctx: Ctx) -> dict:
    """BASELINE configs[3]: LSTM hidden=1024, 4 layers, seq=100, batch=64; batch-sharded at N > 1."""
    torch = ctx.torch
    from tiramisu_b200 import _lib, api, dist as tdist
    L = _lib.lib()
    F, Bn, Ly, T = 1024, 64, 4, 100
    s0, s1 = tdist.batch_shard(Bn, ctx.world, ctx.rank)
    nb = s1 - s0
    with torch.cuda.stream(ctx.stream):
        gw = torch.Generator(device="cuda").manual_seed(200 + ctx.rank)
        W = (torch.rand((Ly, 2, 4 * F, F), device="cuda", generator=gw) * 2 - 1) / 32
        bias = torch.rand((Ly, 4 * F), device="cuda", generator=gw) * 2 - 1
        g = torch.Generator(device="cuda").manual_seed(4 + ctx.rank)
        x = torch.rand((T, max(nb, 1), F), device="cuda", generator=g) * 2 - 1
        y = torch.empty_like(x)
    ctx.stream.synchronize()
    if ctx.world > 1:
        for t in (W, bias):
            _lib.check(L.tiramisu_b200_comm_broadcast(t.data_ptr(), t.numel() * 4, 0))

    def step():
        if nb:
            _lib.check(L.tiramisu_b200_lstm(W.data_ptr(), bias.data_ptr(), x.data_ptr(), y.data_ptr(), F, nb, Ly, T))

    ms = ctx.time_device(step, 5, warm=2)
    flop = T * Ly * 2.0 * (2 * Bn * 4 * F * F)
    out = {"metric": "lstm_GFLOPs", "unit": "GFLOP/s", "value": flop / (ms * 1e-3) / 1e9, "ms": ms,
           "us_per_cell": ms * 1e3 / (T * Ly), "scaling": "strong", "n_gpus": ctx.world,
           "dtype": "tensor-core operands (see engine), fp32 accumulate and fp32 cell state",
           "sharding": f"batch {Bn}/{ctx.world} sequences per rank, weights broadcast once, no per-step collective",
           "l2": "weights 134 MB > L2; no flush",
           "roofline": _roof_tensor(ctx, "tiramisu_sm100_lstm_layer_*", flop, ms, None)}
    try:
        out["engine"] = L.tiramisu_b200_lstm_last_engine().decode()
    except Exception:                                   # older library without the query
        pass

    hW, hb = _pinned(torch, tuple(W.shape)), _pinned(torch, tuple(bias.shape))
    hx, hy = _pinned(torch, tuple(x.shape)), _pinned(torch, tuple(y.shape))
    hW.copy_(W); hb.copy_(bias); hx.copy_(x)
    params = np.array([F, max(nb, 1), Ly, T], np.int32)
    dt = ctx.time_host((lambda: api.lstm(params, hW, hb, hx, hy)) if nb else (lambda: None), 3)
    out["e2e"] = {"value": flop / dt / 1e9, "unit": "GFLOP/s", "ms": dt * 1e3,
                  "h2d_bytes_per_step": int(ctx.sum_over_ranks(float(hW.numel() * 4 + hb.numel() * 4 + T * nb * F * 4))),
                  "d2h_bytes_per_step": int(ctx.sum_over_ranks(float(T * nb * F * 4))),
                  "how": "lstm(params, Weights, biases, x, y) on pinned host buffers"}
    if ctx.cpu:
        from oracle import oracle
        cW, cb = hW.numpy(), hb.numpy()
        Tc = 2
        cx = hx.numpy()[:Tc].copy()
        cell = 2.0 * (2 * 4 * F * F)                                    # FLOP per (layer, step, sequence)
        out["cpu_baseline"] = _cpu_entry(Tc * Ly * Bn * cell, lambda: oracle.lstm(cW, cb, cx), f"{Tc} of {T} time steps, all {Bn} sequences",
                                         1 * Ly * 4 * cell, lambda: oracle.lstm(cW, cb, cx[:1, :4].copy()), "1 time step, 4 sequences")
    del W, x, y, hW, hx, hy
    torch.cuda.empty_cache()
    return out


# --------------------------------------------------------------------------------------------- driver
def run(torch, stream, peaks, world, rank, cpu=True, only=None, quick=False) -> dict:
    """Latency-bound items first, the power-hungry large GEMMs last, with a pause before each: a long TF32
    GEMM leaves the part power-capped (clocks.sm ~900 MHz for a second or so, tools/clock_probe.py) and
    whatever runs next would be timed at half clock."""
    ctx = Ctx(torch, stream, peaks, world, rank, cpu, quick)
    items = [("sgemm_1024", lambda: sgemm(ctx, 1024)), ("lstm", lambda: lstm(ctx)), ("conv_relu_maxpool", lambda: conv(ctx))]
    sweep = (2048, 4096) if quick else (2048, 4096, 8192, 16384)
    items += [(f"sgemm_{n}", (lambda n=n: sgemm(ctx, n, e2e=not quick))) for n in sweep]
    res = {}
    for name, fn in items:
        if only and not any(name.startswith(o) for o in only):
            continue
        time.sleep(0.5)
        try:
            r = fn()
            r.update(_smi_clock(torch))
            res[name] = r
        except Exception as e:                          # never fail the headline line
            res[name] = {"error": repr(e)}
            if world > 1:
                raise                                   # a rank that drops out of a collective would hang the others
    return res


def _smi_clock(torch):
    import subprocess
    try:
        idx = torch.cuda.current_device()
        out = subprocess.run(["nvidia-smi", f"--id={idx}", "--query-gpu=clocks.sm,clocks_event_reasons.sw_power_cap",
                              "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout.strip().splitlines()[0]
        mhz, cap = [x.strip() for x in out.split(",")]
        return {"sm_mhz_after": float(mhz), "sw_power_cap_after": cap}
    except Exception:
        return {}


# --------------------------------------------------------------------------------------------- CPU-only arm
def run_cpu_reference(steps: int) -> dict:
    """--impl reference: the oracle port of every config on the host cores (bounded samples), no GPU."""
    from oracle import oracle
    res = {}
    rng = np.random.default_rng(0)
    for n in (1024, 4096, 16384):
        rs = min(n, max(64, (int(6e9 / (2.0 * n * n)) // 4) * 4))
        A = rng.uniform(-1, 1, (rs, n)).astype(np.float32)
        B = rng.uniform(-1, 1, (n, n)).astype(np.float32)
        Cm = np.zeros((rs, n), np.float32)
        r1c = min(rs, 64)
        e = _cpu_entry(2.0 * rs * n * n, lambda: oracle.sgemm_sched(3.0, 2.0, A, B, Cm), f"{rs} of {n} rows of C",
                       2.0 * r1c * n * n, lambda: oracle.sgemm_sched(3.0, 2.0, A[:r1c], B, Cm[:r1c]), f"{r1c} of {n} rows of C")
        res[f"sgemm_{n}"] = {"metric": "sgemm_GFLOPs", "unit": "GFLOP/s", "value": e["value"], "cpu_baseline": e}
        del A, B, Cm
    Bc, N, FIn, FOut = 32, 224, 64, 64
    cx = rng.uniform(-1, 1, (2, N + 2, N + 2, FIn)).astype(np.float32)
    cw = rng.uniform(-1, 1, (FOut // 8, 3, 3, FIn, 8)).astype(np.float32)
    cb = rng.uniform(-1, 1, (FOut,)).astype(np.float32)
    per_img = 2.0 * N * N * FOut * FIn * 9
    e = _cpu_entry(2 * per_img, lambda: oracle.conv_relu_maxpool(cx, cw, cb), f"2 of {Bc} images",
                   2.0 * 56 * 56 * FOut * FIn * 9, lambda: oracle.conv_relu_maxpool(cx[:1, :58, :58].copy(), cw, cb), "one 56x56 crop")
    res["conv_relu_maxpool"] = {"metric": "conv_relu_maxpool_GFLOPs", "unit": "GFLOP/s", "value": e["value"], "cpu_baseline": e}
    F, Bn, Ly, T = 1024, 64, 4, 100
    W = (rng.uniform(-1, 1, (Ly, 2, 4 * F, F)) / 32).astype(np.float32)
    bias = rng.uniform(-1, 1, (Ly, 4 * F)).astype(np.float32)
    x = rng.uniform(-1, 1, (2, Bn, F)).astype(np.float32)
    cell = 2.0 * (2 * 4 * F * F)
    e = _cpu_entry(2 * Ly * Bn * cell, lambda: oracle.lstm(W, bias, x), f"2 of {T} time steps, all {Bn} sequences",
                   Ly * 4 * cell, lambda: oracle.lstm(W, bias, x[:1, :4].copy()), "1 time step, 4 sequences")
    res["lstm"] = {"metric": "lstm_GFLOPs", "unit": "GFLOP/s", "value": e["value"], "cpu_baseline": e}
    return res
