#!/bin/bash
# r02 run 23 (1 GPU): LSTM fp16 engine with the recurrent weights in TENSOR MEMORY (A operand from TMEM): parity, timing, timeline
set -u
mkdir -p gpurun_out
echo "== lstm tests"; timeout 900 python -m pytest tests/test_gpu_dnn.py -q -x -k lstm 2>&1 | tail -8 | tee gpurun_out/r02_run23_pytest_lstm.log
for v in 1 0; do
  echo "== lstm timing rtmem=$v (pipe 1)"; TB200_LSTM_RTMEM=$v timeout 300 python tools/run_one.py lstm 2>&1 | tail -1 | tee -a gpurun_out/r02_run23_lstm_timing.log
  echo "== lstm timing rtmem=$v (pipe 0)"; TB200_LSTM_RTMEM=$v TB200_LSTM_PIPE=0 timeout 300 python tools/run_one.py lstm 2>&1 | tail -1 | tee -a gpurun_out/r02_run23_lstm_timing.log
done
echo "== lstm timeline (serial, rtmem)"; TB200_LSTM_PIPE=0 TB200_LSTM_TRACE=gpurun_out/lstm_h16_rtmem_trace_serial.bin timeout 300 python tools/run_one.py lstm > /dev/null 2>&1
python tools/lstm_trace.py gpurun_out/lstm_h16_rtmem_trace_serial.bin 100 h16 2>&1 | tee gpurun_out/r02_run23_lstm_h16_rtmem_timeline_serial.txt
