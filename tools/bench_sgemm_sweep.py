"""BASELINE config 5: square sgemm sweep, A and C row-sharded over the ranks, B broadcast from rank 0
over NCCL (the library's own communicator), pipelined by K panels so the broadcast of panel p+1
overlaps the tcgen05 GEMM on panel p.  One process per GPU:

    python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 tools/bench_sgemm_sweep.py [sizes...]

Prints one JSON line per size on rank 0 (aggregate TFLOP/s = 2 n^3 / max-over-ranks time, CUDA events).
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tiramisu_b200 import _lib, api, dist as tdist  # noqa: E402


def main():
    os.environ["NCCL_DEBUG"] = os.environ.get("TB200_NCCL_DEBUG", "WARN")
    sizes = [int(a) for a in sys.argv[1:]] or [1024, 2048, 4096, 8192, 16384]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
        tdist.comm_init_from_torch()
    L = _lib.lib()
    main_s, comm_s = torch.cuda.Stream(), torch.cuda.Stream()
    api.set_stream(main_s.cuda_stream)
    panels = int(os.environ.get("TB200_SWEEP_PANELS", "4"))
    # ncclBroadcast by default; TB200_SWEEP_BCAST=sag selects the scatter + all-gather form (measured
    # slower on 8 B200s: 3.85 vs 3.07 ms for the 1 GiB B of 16384^3, profiles/r01_sgemm_sweep_8gpu.jsonl)
    bcast = L.tiramisu_b200_comm_broadcast_large if os.environ.get("TB200_SWEEP_BCAST") == "sag" else L.tiramisu_b200_comm_broadcast

    for n in sizes:
        r0, r1 = tdist.row_shard(n, world, rank)
        rows = r1 - r0
        with torch.cuda.stream(main_s):
            g = torch.Generator(device="cuda").manual_seed(5 + rank)
            A = torch.rand((rows, n), device="cuda", generator=g) * 2 - 1
            C = torch.zeros((rows, n), device="cuda")
            B = (torch.rand((n, n), device="cuda", generator=g) * 2 - 1) if rank == 0 else torch.empty((n, n), device="cuda")
        main_s.synchronize()
        kw = n // panels
        evs = [torch.cuda.Event() for _ in range(panels)]
        ev_free = torch.cuda.Event()

        def step(pipelined=True):
            if world > 1:
                ev_free.record(main_s)                 # last step's GEMMs no longer read B
                comm_s.wait_event(ev_free)
                api.set_stream(comm_s.cuda_stream)
                for p in range(panels):
                    _lib.check(bcast(B.data_ptr() + p * kw * n * 4, kw * n * 4, 0))
                    evs[p].record(comm_s)
                api.set_stream(main_s.cuda_stream)
            for p in range(panels):
                if world > 1:
                    main_s.wait_event(evs[p] if pipelined else evs[-1])
                api.cublas_sgemm(A.data_ptr(), B.data_ptr(), C.data_ptr(), rows, n, kw, 1.0, 0.0 if p == 0 else 1.0,
                                 ldA=n, ldB=n, ldC=n, offA=p * kw, offB=p * kw * n)

        def gemm_only():
            api.cublas_sgemm(A.data_ptr(), B.data_ptr(), C.data_ptr(), rows, n, n, 1.0, 0.0)

        def timeit(fn, iters):
            for _ in range(2):
                fn()
            main_s.synchronize(); comm_s.synchronize()
            if world > 1:
                torch.distributed.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(main_s)
            for _ in range(iters):
                fn()
            e1.record(main_s)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            if world > 1:
                t = torch.tensor([ms], device="cuda")
                torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
                ms = float(t.item())
            return ms

        iters = 20 if n <= 4096 else (5 if n <= 8192 else 3)
        ms_pipe = timeit(step, iters)
        ms_gemm = timeit(gemm_only, iters)
        if rank == 0:
            fl = 2.0 * n ** 3
            print(json.dumps({"metric": "sgemm_TFLOPs", "n": n, "n_gpus": world, "panels": panels,
                              "bcast": os.environ.get("TB200_SWEEP_BCAST", "nccl"),
                              "ms_step_bcast_pipelined": ms_pipe, "TFLOPs": fl / (ms_pipe * 1e-3) / 1e12,
                              "ms_gemm_only": ms_gemm, "TFLOPs_gemm_only": fl / (ms_gemm * 1e-3) / 1e12,
                              "bcast_bytes_per_rank": n * n * 4 if world > 1 else 0}), flush=True)
        del A, B, C
        torch.cuda.empty_cache()
    if world > 1:
        tdist.comm_destroy()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
