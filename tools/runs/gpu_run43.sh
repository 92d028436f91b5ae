#!/bin/bash
# r01 run 43: ncu --set full of two generated kernels: the warp-shuffle reduction (LSTM gate sums, lowered from the reference's
# LSTM/cpu/generator.cpp) and the faithful lowering of the reference's hand-scheduled GEMM (sgemm/gpu/generator1.cpp)
set -u
mkdir -p gpurun_out
python -c "from tiramisu_b200 import build; build.build()" > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_k4 -s 45 -c 1 -o gpurun_out/prof_gen_lstm_warpreduce python tools/lstm_gen_timing.py > /dev/null 2>&1
cd /tmp && timeout 600 ncu --set full --clock-control none --import-source on -k regex:tiramisu_sm100_matmul -s 1 -c 1 -o $GRAFT_REPO_ROOT/gpurun_out/prof_gen_generator1 $GRAFT_REPO_ROOT/tests/compat/_bin/gen_sgemm_gpu_generator1 2 0 1 > /dev/null 2>&1
ls -la $GRAFT_REPO_ROOT/gpurun_out/*.ncu-rep
