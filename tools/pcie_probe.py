"""Host<->device copy bandwidth of this box (pinned memory): H2D alone, D2H alone, both at once — on one GPU, or
on all ranks AT THE SAME TIME under torchrun (what the e2e legs of bench.py at N > 1 are bounded by: every rank
moves its shard over its own PCIe link, but all links end in the same host memory).

    python tools/pcie_probe.py
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/pcie_probe.py
"""
import json
import os
import time

import torch

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))

n = (1 << 30) // (1 if world <= 2 else 2)
h_a = torch.empty(n, dtype=torch.uint8).pin_memory()
h_b = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    if world > 1:
        torch.distributed.barrier()


def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        barrier()
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best


def h2d():
    with torch.cuda.stream(s1):
        d_a.copy_(h_a, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_b.copy_(d_b, non_blocking=True)


def both():
    h2d(); d2h()


res = {"h2d_GBps": n / t(h2d) / 1e9, "d2h_GBps": n / t(d2h) / 1e9, "duplex_each_GBps": n / t(both) / 1e9}
if world > 1:
    v = torch.tensor([res["h2d_GBps"], res["d2h_GBps"], res["duplex_each_GBps"]], device="cuda", dtype=torch.float64)
    lo = v.clone(); torch.distributed.all_reduce(lo, op=torch.distributed.ReduceOp.MIN)
    sm = v.clone(); torch.distributed.all_reduce(sm)
    res = {"ranks_concurrent": world, "bytes_per_copy": n,
           "per_rank_min_GBps": dict(zip(("h2d", "d2h", "duplex_each"), lo.tolist())),
           "aggregate_GBps": dict(zip(("h2d", "d2h", "duplex_each_direction"), sm.tolist()))}
res["cpus"] = os.cpu_count()
try:
    res["numa_nodes"] = len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")])
except OSError:
    pass
if rank == 0:
    print(json.dumps(res))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open(f"gpurun_out/pcie_probe_n{world}.json", "w"))
if world > 1:
    torch.distributed.destroy_process_group()
