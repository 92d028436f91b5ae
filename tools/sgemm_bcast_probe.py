"""2-rank (or more) diagnosis of the fused B broadcast: column form, its broadcast alone (rows = 0), stream-side
gating, GEMM alone.  torchrun --nproc-per-node 2 tools/sgemm_bcast_probe.py [n ...]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tiramisu_b200 import _lib, api, dist as tdist  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    tdist.comm_init_from_torch()
    s = torch.cuda.Stream()
    api.set_stream(s.cuda_stream)
    for n in [int(a) for a in sys.argv[1:]] or [8192, 16384]:
        r0, r1 = tdist.row_shard(n, world, rank)
        rows = r1 - r0
        with torch.cuda.stream(s):
            A = torch.rand((rows, n), device="cuda")
            Cm = torch.zeros((rows, n), device="cuda")
        symB = tdist.SymmetricB(n, n)
        with torch.cuda.stream(s):
            B = torch.rand((n, n), device="cuda") if rank == 0 else None
        Bp = B.data_ptr() if rank == 0 else None
        s.synchronize()

        def timed(fn, iters=4):
            for _ in range(2):
                fn()
            s.synchronize(); torch.distributed.barrier(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            for _ in range(iters):
                fn()
            e1.record(s)
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda")
            mine = float(t.item())
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            return float(t.item()), mine

        res = {"n": n, "world": world}
        for panels in sorted({symB.sgemm_panels(), 16, 8, 4, 2}, reverse=True):
            res[f"p{panels}_fused"] = timed(lambda: symB.sgemm(Bp, A.data_ptr(), Cm.data_ptr(), rows, 1.0, 0.0, 0, panels))
            res[f"p{panels}_bcast_only"] = timed(lambda: symB.sgemm(Bp, A.data_ptr(), Cm.data_ptr(), 0, 1.0, 0.0, 0, panels))
        if n >= 12288:
            os.environ["TB200_GEMM_NO_KPANELS"] = "1"
            pd = symB.sgemm_panels()
            res[f"p{pd}_fused_one_launch_over_K"] = timed(lambda: symB.sgemm(Bp, A.data_ptr(), Cm.data_ptr(), rows, 1.0, 0.0, 0, pd))
            res["gemm_only_one_launch_over_K"] = timed(lambda: api.cublas_sgemm(A.data_ptr(), symB.data_ptr(), Cm.data_ptr(), rows, n, n, 1.0, 0.0))
            del os.environ["TB200_GEMM_NO_KPANELS"]
        res["plain_bcast_8"] = timed(lambda: symB.broadcast(0, 8))
        res["gemm_only"] = timed(lambda: api.cublas_sgemm(A.data_ptr(), symB.data_ptr(), Cm.data_ptr(), rows, n, n, 1.0, 0.0))
        gathered = [None] * world
        torch.distributed.all_gather_object(gathered, {k: (round(v[1], 4) if isinstance(v, tuple) else v) for k, v in res.items()})
        if rank == 0:
            print(json.dumps({"max_over_ranks_ms": {k: (round(v[0], 4) if isinstance(v, tuple) else v) for k, v in res.items()},
                              "per_rank_ms": gathered}), flush=True)
        symB.free()
        del A, Cm
        torch.cuda.empty_cache()
    tdist.comm_destroy()
    torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
