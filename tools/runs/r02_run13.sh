#!/bin/bash
# r02 run 13 (8 GPUs): copy-engine traffic patterns over NVSwitch between IPC-mapped blocks (pull / push, 1..7 streams)
set -u
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29731 tools/p2p_pattern_probe.py 2>&1 | grep -v "^\*\|OMP_NUM" | tee gpurun_out/r02_run13_p2p_patterns.jsonl
