#!/bin/bash
# r02 run 14 (8 GPUs): event timeline of the DMA broadcast (where do 3.4 ms go when the link-level probe says 1.4 + 1.4?)
set -u
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29741 tools/bcast_trace.py 16384 p1 p4 s4 s16 > gpurun_out/r02_run14_bcast_trace.log 2>&1
echo rc=$?; grep -c "rank" gpurun_out/r02_run14_bcast_trace.log; grep "rank 1\]" gpurun_out/r02_run14_bcast_trace.log | head -150
