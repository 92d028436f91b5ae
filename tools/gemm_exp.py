"""GEMM tuning sweep (one process, env knobs flipped between timings): kernel choice, L2 prefetch distance,
programmatic dependent launch.  usage: python tools/gemm_exp.py [sizes...]   -> one JSON line per variant."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tiramisu_b200 import api  # noqa: E402


def timed(stream, fn, iters):
    for _ in range(3):
        fn()
    stream.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(iters):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def timed_graph(stream, fn, n_launch, reps):
    """device-side period: n_launch calls captured once in a CUDA graph, replayed `reps` times"""
    g = torch.cuda.CUDAGraph()
    stream.synchronize()
    with torch.cuda.graph(g, stream=stream):
        for _ in range(n_launch):
            fn()
    with torch.cuda.stream(stream):
        g.replay(); g.replay()
        stream.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            g.replay()
        e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * n_launch)


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [1024, 2048, 4096]
    stream = torch.cuda.Stream()
    api.set_stream(stream.cuda_stream)
    for n in sizes:
        nsets = max(1, int(np.ceil(160e6 / (3 * n * n * 4)))) if n <= 2048 else 1
        with torch.cuda.stream(stream):
            sets = [(torch.rand((n, n), device="cuda") * 2 - 1, torch.rand((n, n), device="cuda") * 2 - 1, torch.zeros((n, n), device="cuda"))
                    for _ in range(nsets)]
        idx = [0]

        def fn():
            a, b, c = sets[idx[0] % nsets]
            idx[0] += 1
            api.cublas_sgemm(a.data_ptr(), b.data_ptr(), c.data_ptr(), n, n, n, 3.0, 2.0)

        if n <= 1024:
            variants = [{"TB200_GEMM_KERNEL": "1"}, {"TB200_GEMM_KERNEL": "4", "TB200_GEMM_NO_PDL": "1"}, {"TB200_GEMM_KERNEL": "4"},
                        {"TB200_GEMM_KERNEL": "4", "TB200_GEMM_PREFETCH": "16"}, {}]
        else:
            variants = [{}, {"TB200_GEMM_STREAMK": "0"}, {"TB200_GEMM_STREAMK": "1"}, {"TB200_GEMM_KERNEL": "2"}]
        for env in variants:
            for k in ("TB200_GEMM_KERNEL", "TB200_GEMM_PREFETCH", "TB200_GEMM_NO_PDL", "TB200_GEMM_STREAMK"):
                os.environ.pop(k, None)
            os.environ.update(env)
            ms = timed(stream, fn, 30 if n <= 2048 else 10)
            rec = {"n": n, "env": env, "us": ms * 1e3, "TFLOPs": 2.0 * n ** 3 / (ms * 1e-3) / 1e12, "operand_sets": nsets}
            if n <= 2048:
                try:
                    mg = timed_graph(stream, fn, max(nsets, 8), 5)
                    rec.update({"us_graph": mg * 1e3, "TFLOPs_graph": 2.0 * n ** 3 / (mg * 1e-3) / 1e12})
                except Exception as e:
                    rec["graph_error"] = repr(e)
            print(json.dumps(rec), flush=True)
        # warm-L2 number for the small sizes (one operand set): what a caller that re-uses operands sees
        if nsets > 1:
            for k in ("TB200_GEMM_KERNEL", "TB200_GEMM_PREFETCH", "TB200_GEMM_NO_PDL", "TB200_GEMM_STREAMK"):
                os.environ.pop(k, None)
            a, b, c = sets[0]
            ms = timed(stream, lambda: api.cublas_sgemm(a.data_ptr(), b.data_ptr(), c.data_ptr(), n, n, n, 3.0, 2.0), 30)
            print(json.dumps({"n": n, "env": "default, ONE operand set (L2-warm)", "us": ms * 1e3, "TFLOPs": 2.0 * n ** 3 / (ms * 1e-3) / 1e12}), flush=True)
        del sets
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
