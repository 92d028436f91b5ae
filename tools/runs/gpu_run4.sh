#!/bin/bash
# ncu captures of the current kernels (one kernel each)
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
NCU="ncu --set full --clock-control none --import-source on -f"
echo "== ncu gemm cg2 4096" ; timeout 600 $NCU -k regex:gemm_tf32_persistent -s 2 -c 1 -o gpurun_out/prof_gemm_cg2_4096 python tools/run_one.py sgemm4096 > gpurun_out/ncu_gemm.log 2>&1 ; tail -1 gpurun_out/ncu_gemm.log
echo "== ncu gemm cg2 8192" ; timeout 600 $NCU -k regex:gemm_tf32_persistent -s 2 -c 1 -o gpurun_out/prof_gemm_cg2_8192 python tools/run_one.py sgemm8192 > gpurun_out/ncu_gemm8.log 2>&1 ; tail -1 gpurun_out/ncu_gemm8.log
echo "== ncu conv" ; timeout 600 $NCU -k regex:conv_block_tf32 -s 1 -c 1 -o gpurun_out/prof_conv2 python tools/run_one.py conv > gpurun_out/ncu_conv.log 2>&1 ; tail -1 gpurun_out/ncu_conv.log
echo "== ncu lstm layer" ; timeout 900 $NCU -k regex:lstm_layer -s 1 -c 1 -o gpurun_out/prof_lstm_layer python tools/run_one.py lstm > gpurun_out/ncu_lstm.log 2>&1 ; tail -1 gpurun_out/ncu_lstm.log
ls -la gpurun_out/*.ncu-rep
