#!/bin/bash
# r02 run 8 (2 GPUs): peer 1D / 2D copy bandwidth probe + K-panel GEMM cost
set -u
mkdir -p gpurun_out
timeout 300 python tools/copy2d_probe.py 2>&1 | tee gpurun_out/r02_run8_copy2d_probe.jsonl
