#!/bin/bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
echo "== lstm trace"; TB200_LSTM_TRACE=gpurun_out/lstm_trace.bin TB200_LSTM_KERNEL=3 timeout 300 python tools/run_one.py lstm 2>&1 | tail -1
python tools/lstm_trace.py gpurun_out/lstm_trace.bin 100
echo "== gemm k1" ; timeout 300 python tools/gemm_diag2.py 1 all > gpurun_out/gemm_k1.log 2>&1; grep -c "bad=0" gpurun_out/gemm_k1.log ; grep -v "bad=0" gpurun_out/gemm_k1.log | tail -9
