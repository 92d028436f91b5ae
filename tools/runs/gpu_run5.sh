#!/bin/bash
# run 5: lean MMA issue (uniform control flow), conv all-taps tile
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
echo "== conv tests kernel1/2 (base_offset=0)" ; timeout 600 python -m pytest tests/test_gpu_dnn.py -m gpu -q -k conv 2>&1 | tail -8 | tee gpurun_out/test_conv.log
echo "== conv tests (base_offset=1)" ; TB200_CONV_BASE_OFFSET=1 timeout 600 python -m pytest tests/test_gpu_dnn.py -m gpu -q -k "conv and ints" 2>&1 | tail -8 | tee gpurun_out/test_conv_bo1.log
for k in 1 2; do echo "== conv time kernel=$k"; TB200_CONV_KERNEL=$k timeout 300 python tools/run_one.py conv 2>&1 | tail -1; done
echo "== conv time kernel=2 bo=1"; TB200_CONV_KERNEL=2 TB200_CONV_BASE_OFFSET=1 timeout 300 python tools/run_one.py conv 2>&1 | tail -1
echo "== gemm k1" ; timeout 300 python tools/gemm_diag2.py 1 time 2>&1 | tail -8
echo "== gemm k3" ; timeout 300 python tools/gemm_diag2.py 3 all > gpurun_out/gemm_k3.log 2>&1 ; echo "rc=$?" ; grep -c "bad=0" gpurun_out/gemm_k3.log ; grep -v "bad=0" gpurun_out/gemm_k3.log | tail -10
echo "== gemm k2" ; timeout 300 python tools/gemm_diag2.py 2 time 2>&1 | tail -8
echo "== lstm tests" ; timeout 900 python -m pytest tests/test_gpu_dnn.py -m gpu -q -k lstm 2>&1 | tail -6 | tee gpurun_out/test_lstm.log
for cfg in "1 1" "2 1" "2 2" "2 4"; do set -- $cfg; echo "== lstm time kernel=$1 cluster=$2"; TB200_LSTM_KERNEL=$1 TB200_LSTM_CLUSTER=$2 timeout 300 python tools/run_one.py lstm 2>&1 | tail -1; done
echo "== all gpu tests" ; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/test_all.log
