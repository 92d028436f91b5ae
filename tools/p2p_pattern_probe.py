"""NVLink copy-engine patterns between the ranks' symmetric (CUDA-IPC mapped) blocks: pull vs push, one source vs a
ring, one stream vs several — decides how the B panels of the row-sharded sgemm travel (r02 run 13).
torchrun --nproc-per-node 8 tools/p2p_pattern_probe.py"""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tiramisu_b200 import _lib, dist as tdist  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    tdist.comm_init_from_torch()
    L = _lib.lib()
    rt = C.CDLL("libcudart.so.12")
    rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
    chunk = 128 << 20
    nbytes = chunk * world
    loc = C.c_void_p()
    peers = (C.c_void_p * world)()
    _lib.check(L.tiramisu_b200_comm_alloc_symmetric(nbytes, C.byref(loc), peers))
    streams = [torch.cuda.Stream() for _ in range(8)]

    def copy(dst_rank, dst_off, src_rank, src_off, n, s):
        rc = rt.cudaMemcpyAsync(peers[dst_rank] + dst_off, peers[src_rank] + src_off, n, 3, C.c_void_p(s.cuda_stream))
        assert rc == 0, rc

    def timed(name, fn, moved_bytes, reps=3):
        fn(); torch.cuda.synchronize(); torch.distributed.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for s in streams:
            s.wait_event(e0)
        for _ in range(reps):
            fn()
        for s in streams:
            ev = torch.cuda.Event(); ev.record(s); torch.cuda.current_stream().wait_event(ev)
        e1.record()
        torch.cuda.synchronize()
        torch.distributed.barrier()
        ms = e0.elapsed_time(e1) / reps
        t = torch.tensor([ms], device="cuda")
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        if rank == 0:
            print(json.dumps({"pattern": name, "ms": round(float(t.item()), 4), "GBps_per_busiest_port": round(moved_bytes / float(t.item()) / 1e6, 1)}), flush=True)

    # (a) everybody pulls its chunk from rank 0 (root egress carries world-1 chunks)
    timed("scatter: every rank PULLS one 128 MiB chunk from rank 0", lambda: rank and copy(rank, 0, 0, rank * chunk, chunk, streams[0]), chunk * (world - 1))
    # (b) rank 0 pushes a chunk to every rank: one stream, two, one per destination
    for ns in (1, 2, 4, 7):
        def push(ns=ns):
            if rank == 0:
                for q in range(1, world):
                    copy(q, 0, 0, q * chunk, chunk, streams[(q - 1) % ns])
        timed(f"scatter: rank 0 PUSHES a chunk to every rank on {ns} stream(s)", push, chunk * (world - 1))
    # (c) ring: one chunk from / to the next rank, all ranks at once
    timed("ring: every rank PULLS a chunk from rank+1", lambda: copy(rank, 0, (rank + 1) % world, chunk, chunk, streams[0]), chunk)
    timed("ring: every rank PUSHES a chunk to rank+1", lambda: copy((rank + 1) % world, chunk, rank, 0, chunk, streams[0]), chunk)
    # (d) all-gather step: every rank moves its chunk to / from all others
    def ag_pull(ns):
        for k in range(1, world):
            q = (rank + k) % world
            copy(rank, q * chunk, q, q * chunk, chunk, streams[(k - 1) % ns])
    def ag_push(ns):
        for k in range(1, world):
            q = (rank + k) % world
            copy(q, rank * chunk, rank, rank * chunk, chunk, streams[(k - 1) % ns])
    for ns in (1, 2, 4):
        timed(f"all-gather: every rank PULLS the other chunks on {ns} stream(s)", lambda ns=ns: ag_pull(ns), chunk * (world - 1))
        timed(f"all-gather: every rank PUSHES its chunk to the others on {ns} stream(s)", lambda ns=ns: ag_push(ns), chunk * (world - 1))
    # (e) small copies: 8 MiB slices (what a 64 MiB panel cut in world-1 slices looks like)
    small = 8 << 20
    def ag_push_small():
        for rep in range(16):
            for k in range(1, world):
                q = (rank + k) % world
                copy(q, rank * chunk + rep * small, rank, rank * chunk + rep * small, small, streams[(k - 1) % 2])
    timed("all-gather: PUSH, 16 x 8 MiB slices per destination, 2 streams", ag_push_small, 16 * small * (world - 1))
    torch.cuda.synchronize()
    torch.distributed.barrier()
    _lib.check(L.tiramisu_b200_comm_free_symmetric(loc))
    tdist.comm_destroy()
    torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
