#!/bin/bash
# r02 run 22 (1 GPU): LSTM fp16 engine with two half-batch pipelines per CTA (h16x2) + rolling layer pipeline: parity tests,
# timing of the four (split, roll) combinations, step timeline; GEMM: k-aligned split of the last wave (tests + sweep)
set -u
mkdir -p gpurun_out
echo "== lstm tests"; timeout 900 python -m pytest tests/test_gpu_dnn.py -q -x -k lstm 2>&1 | tail -8 | tee gpurun_out/r02_run22_pytest_lstm.log
for v in "1 1" "1 0" "0 1" "0 0"; do
  set -- $v
  echo "== lstm timing split=$1 roll=$2 (pipe 1)"
  TB200_LSTM_SPLIT=$1 TB200_LSTM_ROLL=$2 timeout 300 python tools/run_one.py lstm 2>&1 | tail -1 | tee -a gpurun_out/r02_run22_lstm_timing.log
done
echo "== lstm timing split=1 pipe=0"; TB200_LSTM_PIPE=0 timeout 300 python tools/run_one.py lstm 2>&1 | tail -1 | tee -a gpurun_out/r02_run22_lstm_timing.log
echo "== lstm timing split=1 roll=1 chunk 5"; TB200_LSTM_CHUNK=5 timeout 300 python tools/run_one.py lstm 2>&1 | tail -1 | tee -a gpurun_out/r02_run22_lstm_timing.log
echo "== lstm timeline (serial, split)"; TB200_LSTM_PIPE=0 TB200_LSTM_TRACE=gpurun_out/lstm_h16x2_trace_serial.bin timeout 300 python tools/run_one.py lstm > /dev/null 2>&1
python tools/lstm_trace.py gpurun_out/lstm_h16x2_trace_serial.bin 100 h16 2>&1 | tee gpurun_out/r02_run22_lstm_h16x2_timeline_serial.txt
echo "== gemm tests"; timeout 900 python -m pytest tests/test_gpu_gemm.py -q -x 2>&1 | tail -5 | tee gpurun_out/r02_run22_pytest_gemm.log
echo "== gemm sweep"; timeout 600 python tools/gemm_exp.py 4096 8192 2>&1 | cut -c1-300 | tee gpurun_out/r02_run22_gemm_exp.jsonl
