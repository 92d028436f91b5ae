#!/bin/bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
echo "== conv tests" ; timeout 600 python -m pytest tests/test_gpu_dnn.py -m gpu -q -k conv 2>&1 | tail -8 | tee gpurun_out/test_conv.log
for k in 2 3; do echo "== conv time kernel=$k"; TB200_CONV_KERNEL=$k timeout 300 python tools/run_one.py conv 2>&1 | tail -1; done
echo "== sgemm 16384 (k panels)"; timeout 300 python tools/run_one.py sgemm16384 2>&1 | tail -1
echo "== gemm tests"; timeout 600 python -m pytest tests/test_gpu_gemm.py -m gpu -q 2>&1 | tail -4
NCU="ncu --set full --clock-control none --import-source on -f"
echo "== ncu conv pair" ; TB200_CONV_KERNEL=3 timeout 600 $NCU -k regex:conv_block_tf32 -s 1 -c 1 -o gpurun_out/prof_conv_pair python tools/run_one.py conv > gpurun_out/ncu_conv.log 2>&1 ; tail -1 gpurun_out/ncu_conv.log
