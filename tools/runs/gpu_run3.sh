#!/bin/bash
# round-1 run 3: new kernels (conv ky-reuse, persistent GEMM CG=1/2, persistent LSTM layer), pipelined e2e
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
echo "== pcie" ; timeout 120 python tools/pcie_probe.py 2>&1 | tail -1
echo "== gemm k1 BN sweep @1024 (time only)"
for bn in 64 128 256; do TB200_GEMM_BN=$bn timeout 120 python tools/gemm_diag2.py 1 time 2>&1 | grep "n=1024" | sed "s/^/BN=$bn /"; done
echo "== gemm k1 full" ; timeout 300 python tools/gemm_diag2.py 1 time 2>&1 | tail -8
echo "== gemm k2" ; timeout 300 python tools/gemm_diag2.py 2 all > gpurun_out/gemm_k2.log 2>&1 ; echo "rc=$?" ; grep -c "bad=0" gpurun_out/gemm_k2.log ; grep -v "bad=0" gpurun_out/gemm_k2.log | tail -14
echo "== gemm k3" ; timeout 300 python tools/gemm_diag2.py 3 all > gpurun_out/gemm_k3.log 2>&1 ; echo "rc=$?" ; grep -c "bad=0" gpurun_out/gemm_k3.log ; grep -v "bad=0" gpurun_out/gemm_k3.log | tail -14
echo "== conv tests" ; timeout 600 python -m pytest tests/test_gpu_dnn.py -m gpu -q -k conv 2>&1 | tail -8 | tee gpurun_out/test_conv.log
echo "== conv time" ; timeout 300 python tools/run_one.py conv 2>&1 | tail -1
echo "== lstm tests" ; timeout 900 python -m pytest tests/test_gpu_dnn.py -m gpu -q -k lstm 2>&1 | tail -12 | tee gpurun_out/test_lstm.log
for cfg in "1 1" "2 1" "2 2" "2 4"; do set -- $cfg; echo "== lstm time kernel=$1 cluster=$2"; TB200_LSTM_KERNEL=$1 TB200_LSTM_CLUSTER=$2 timeout 300 python tools/run_one.py lstm 2>&1 | tail -1; done
echo "== all gpu tests" ; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/test_all.log
echo "== bench" ; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err ; cat gpurun_out/bench.json ; tail -3 gpurun_out/bench.err
ls gpurun_out
