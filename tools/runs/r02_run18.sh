#!/bin/bash
# r02 run 18 (4 GPUs): generated distributed functions over NCCL (tools/dist_gen_check.py at 4 and 2 ranks)
set -u
mkdir -p gpurun_out
for k in 4 2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $k --master-addr 127.0.0.1 --master-port 2975$k tools/dist_gen_check.py 2>&1 | grep -v "^\*\|OMP_NUM" | tee gpurun_out/r02_run18_dist_gen_check_${k}gpu.log | tail -12
done
