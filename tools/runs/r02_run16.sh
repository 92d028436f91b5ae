#!/bin/bash
# r02 run 16 (2 GPUs): root-push scatter, flag fan-out kernels, pipelined re-layout, helper kernels preloaded: parity + overlap probe
set -u
mkdir -p gpurun_out
echo "== dist_check (2 ranks)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 tools/dist_check.py > gpurun_out/r02_run16_dist_check_2gpu.log 2>&1
echo "rc=$?"; grep -c " ok" gpurun_out/r02_run16_dist_check_2gpu.log; grep -v " ok$" gpurun_out/r02_run16_dist_check_2gpu.log | grep -v "^\*\|OMP_NUM\|^$" | tail -12
echo "== probe"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29721 tools/sgemm_bcast_probe.py 8192 16384 2>&1 | grep -v "^\*\|OMP_NUM" | tee gpurun_out/r02_run16_sgemm_bcast_probe.jsonl | cut -c1-1200
