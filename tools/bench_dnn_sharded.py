"""Batch-sharded conv+ReLU+maxpool (BASELINE configs[2]) and LSTM (configs[3]) on N GPUs, one process per GPU:

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_dnn_sharded.py

SURVEY §8e: both blocks shard over the batch (independent units), weights are broadcast once over the
library's NCCL communicator, there is no per-step collective.  Reported per block:
  strong — the BASELINE batch (32 images / 64 sequences) split over the ranks,
  weak   — the BASELINE batch on every rank (replicas).
Timing: CUDA events on the launch stream, barrier on both sides, max over ranks; aggregate TFLOP/s.
Parity: every rank's shard must equal the same rows of the full-batch result computed on that rank
(conv bit-exact — images are independent; LSTM to 1e-3, a different batch picks a different tile shape).
One JSON line per (block, mode) on rank 0."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tiramisu_b200 import _lib, api, dist as tdist  # noqa: E402


def timed(stream, fn, iters, warm, world):
    for _ in range(warm):
        fn()
    stream.synchronize()
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(iters):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def main():
    os.environ.pop("NCCL_DEBUG", None)
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
        tdist.comm_init_from_torch()
    L = _lib.lib()
    stream = torch.cuda.Stream()
    api.set_stream(stream.cuda_stream)

    def bcast(t):           # weights: rank 0's values everywhere, once, through the library's communicator
        if world > 1:
            _lib.check(L.tiramisu_b200_comm_broadcast(t.data_ptr(), t.numel() * t.element_size(), 0))

    out = []
    # ---- conv+ReLU+maxpool, 32 x 224 x 224, 64 -> 64 --------------------------------------------------------
    B, N, FIn, FOut = 32, 224, 64, 64
    with torch.cuda.stream(stream):
        g = torch.Generator(device="cuda").manual_seed(3)                     # same seed on every rank: every rank can form the full batch
        x = torch.rand((B, N + 2, N + 2, FIn), device="cuda", generator=g) * 2 - 1
        gw = torch.Generator(device="cuda").manual_seed(100 + rank)           # weights differ per rank until the broadcast
        w = torch.rand((FOut // 8, 3, 3, FIn, 8), device="cuda", generator=gw) * 2 - 1
        b = torch.rand((FOut,), device="cuda", generator=gw) * 2 - 1
        bcast(w); bcast(b)
        o_full = torch.empty((B, FOut // 8, N // 2, N // 2, 8), device="cuda")
    stream.synchronize()

    def conv(xs, os_):
        _lib.check(L.tiramisu_b200_conv_relu_maxpool(xs.data_ptr(), w.data_ptr(), b.data_ptr(), os_.data_ptr(), xs.shape[0], N, FIn, FOut))

    b0, b1 = tdist.batch_shard(B, world, rank)
    o_shard = torch.empty((b1 - b0,) + tuple(o_full.shape[1:]), device="cuda")
    conv(x, o_full); conv(x[b0:b1], o_shard); stream.synchronize()
    conv_ok = bool(torch.equal(o_shard, o_full[b0:b1]))
    flop = 2.0 * B * N * N * FOut * FIn * 9
    for mode, fn, total in (("strong", lambda: conv(x[b0:b1], o_shard), flop), ("weak", lambda: conv(x, o_full), flop * world)):
        ms = timed(stream, fn, 10, 3, world)
        out.append({"block": "conv_relu_maxpool_32x224x224_64to64", "scaling": mode, "n_gpus": world, "ms": ms,
                    "TFLOPs": total / (ms * 1e-3) / 1e12, "batch_per_gpu": (b1 - b0) if mode == "strong" else B, "shard_parity": conv_ok})
    del x, o_full, o_shard
    torch.cuda.empty_cache()

    # ---- LSTM F=1024, L=4, T=100, B=64 -----------------------------------------------------------------------
    F, Bn, Ly, T = 1024, 64, 4, 100
    with torch.cuda.stream(stream):
        gw = torch.Generator(device="cuda").manual_seed(200 + rank)
        W = (torch.rand((Ly, 2, 4 * F, F), device="cuda", generator=gw) * 2 - 1) / 32
        bias = torch.rand((Ly, 4 * F), device="cuda", generator=gw) * 2 - 1
        bcast(W); bcast(bias)
        g = torch.Generator(device="cuda").manual_seed(4)
        xs = torch.rand((T, Bn, F), device="cuda", generator=g) * 2 - 1
        y_full = torch.empty_like(xs)
    stream.synchronize()
    s0, s1 = tdist.batch_shard(Bn, world, rank)
    x_sh = xs[:, s0:s1].contiguous()
    y_sh = torch.empty_like(x_sh)

    def lstm(xi, yo):
        _lib.check(L.tiramisu_b200_lstm(W.data_ptr(), bias.data_ptr(), xi.data_ptr(), yo.data_ptr(), F, xi.shape[1], Ly, T))

    lstm(xs, y_full); lstm(x_sh, y_sh); stream.synchronize()
    err = float((y_sh - y_full[:, s0:s1]).abs().max())
    flop = T * Ly * 2.0 * (2 * Bn * 4 * F * F)
    for mode, fn, total in (("strong", lambda: lstm(x_sh, y_sh), flop), ("weak", lambda: lstm(xs, y_full), flop * world)):
        ms = timed(stream, fn, 5, 2, world)
        out.append({"block": "lstm_F1024_L4_T100_B64", "scaling": mode, "n_gpus": world, "ms": ms, "TFLOPs": total / (ms * 1e-3) / 1e12,
                    "batch_per_gpu": (s1 - s0) if mode == "strong" else Bn, "shard_max_abs_diff": err, "shard_parity": err < 1e-3})

    ok = all(o["shard_parity"] for o in out)
    if world > 1:
        flag = torch.tensor([1 if ok else 0], device="cuda")
        torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN)
        ok = int(flag.item()) == 1
    if rank == 0:
        for o in out:
            print(json.dumps(o), flush=True)
        print("DNN_SHARDED_OK" if ok else "DNN_SHARDED_PARITY_FAILED", flush=True)
    if world > 1:
        tdist.comm_destroy()
        torch.distributed.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
