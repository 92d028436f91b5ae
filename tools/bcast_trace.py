"""Timeline of one traced broadcast per configuration (TB200_BCAST_TRACE=1).  torchrun --nproc-per-node N tools/bcast_trace.py n panels..."""
import os
import sys

import torch

os.environ["TB200_BCAST_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tiramisu_b200 import _lib, api, dist as tdist  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    tdist.comm_init_from_torch()
    L = _lib.lib()
    s = torch.cuda.Stream()
    api.set_stream(s.cuda_stream)
    n = int(sys.argv[1])
    symB = tdist.SymmetricB(n, n)
    with torch.cuda.stream(s):
        B = torch.rand((n, n), device="cuda") if rank == 0 else None
        A = torch.rand((n // world, n), device="cuda")
        Cm = torch.zeros((n // world, n), device="cuda")
    s.synchronize()
    for spec in sys.argv[2:]:
        kind, panels = spec[0], int(spec[1:])
        for rep in range(3):                       # the third one is printed
            torch.cuda.synchronize(); torch.distributed.barrier()
            if kind == "s":
                symB.sgemm(B.data_ptr() if rank == 0 else None, A.data_ptr(), Cm.data_ptr(), 0, 1.0, 0.0, 0, panels)
            elif kind == "f":
                symB.sgemm(B.data_ptr() if rank == 0 else None, A.data_ptr(), Cm.data_ptr(), n // world, 1.0, 0.0, 0, panels)
            else:
                symB.broadcast(0, panels)
            torch.cuda.synchronize()
        if rank in (0, 1, world - 1):
            sys.stderr.write(f"[rank {rank}] ==== n={n} {'staged sgemm broadcast (no GEMM)' if kind == 's' else 'fused sgemm' if kind == 'f' else 'plain broadcast'} panels={panels}\n")
            L.tiramisu_b200_comm_broadcast_trace_dump()
        torch.distributed.barrier()
    symB.free()
    tdist.comm_destroy()
    torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
