#!/bin/bash
# r02 run 9 (2 GPUs): where does the fused B broadcast lose its overlap?  (column / K panel forms, broadcast alone, stream-side gate)
set -u
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29721 tools/sgemm_bcast_probe.py 8192 16384 2>&1 | grep -v "^\*\|OMP_NUM" | tee gpurun_out/r02_run9_sgemm_bcast_probe.jsonl
