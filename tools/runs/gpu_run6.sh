#!/bin/bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
for d in 0 1; do for c in 1 4; do echo "== lstm dbg=$d cluster=$c"; TB200_LSTM_DBG=$d TB200_LSTM_KERNEL=2 TB200_LSTM_CLUSTER=$c timeout 300 python tools/run_one.py lstm 2>&1 | tail -1; done; done
NCU="ncu --set full --clock-control none --import-source on -f"
echo "== ncu lstm layer" ; TB200_LSTM_CLUSTER=1 timeout 900 $NCU -k regex:lstm_layer -s 1 -c 1 -o gpurun_out/prof_lstm_layer2 python tools/run_one.py lstm > gpurun_out/ncu_lstm.log 2>&1 ; tail -1 gpurun_out/ncu_lstm.log
echo "== ncu conv" ; timeout 600 $NCU -k regex:conv_block_tf32 -s 1 -c 1 -o gpurun_out/prof_conv3 python tools/run_one.py conv > gpurun_out/ncu_conv.log 2>&1 ; tail -1 gpurun_out/ncu_conv.log
echo "== ncu gemm 1024" ; timeout 600 $NCU -k regex:gemm_tf32 -s 5 -c 1 -o gpurun_out/prof_gemm1024 python tools/run_one.py sgemm1024 > gpurun_out/ncu_g1024.log 2>&1 ; tail -1 gpurun_out/ncu_g1024.log
