#!/bin/bash
# r02 run 2: split-K GEMM kernel bring-up (tests + tuning sweep), LSTM step-barrier variants, fp16 range tests
set -u
mkdir -p gpurun_out
echo "== gemm tests"; timeout 900 python -m pytest tests/test_gpu_gemm.py -q -x 2>&1 | tail -15 | tee gpurun_out/r02_run2_pytest_gemm.log
echo "== gemm sweep"; timeout 600 python tools/gemm_exp.py 1024 2048 4096 2>&1 | tee gpurun_out/r02_run2_gemm_exp.jsonl
for sync in 0 2; do
  echo "== lstm timing sync=$sync (kernel 4, pipe 1, then pipe 0)"
  TB200_LSTM_SYNC=$sync TB200_LSTM_KERNEL=4 TB200_LSTM_PIPE=1 timeout 300 python tools/run_one.py lstm 2>&1 | tail -1 | tee -a gpurun_out/r02_run2_lstm_timing.log
  TB200_LSTM_SYNC=$sync TB200_LSTM_KERNEL=4 TB200_LSTM_PIPE=0 timeout 300 python tools/run_one.py lstm 2>&1 | tail -1 | tee -a gpurun_out/r02_run2_lstm_timing.log
done
echo "== dnn tests"; timeout 900 python -m pytest tests/test_gpu_dnn.py -q 2>&1 | tail -15 | tee gpurun_out/r02_run2_pytest_dnn.log
