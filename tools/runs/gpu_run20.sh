#!/bin/bash
# r01 run 20: fp16 K-split LSTM engine + layer-pair pipeline — parity, timing, step timeline
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
echo "== pytest lstm"; timeout 900 python -m pytest tests/test_gpu_dnn.py -m gpu -q -x -k "lstm" 2>&1 | tail -15 | tee gpurun_out/run20_pytest_lstm.log
for cfg in "3 1" "4 0" "4 1"; do
  set -- $cfg
  echo "== lstm timing kernel=$1 pipe=$2"
  TB200_LSTM_KERNEL=$1 TB200_LSTM_PIPE=$2 timeout 300 python tools/run_one.py lstm 2>&1 | tail -3 | tee -a gpurun_out/run20_lstm_timing.log
done
for ch in 5 20; do
  echo "== lstm timing kernel=4 pipe=1 chunk=$ch"
  TB200_LSTM_KERNEL=4 TB200_LSTM_PIPE=1 TB200_LSTM_CHUNK=$ch timeout 300 python tools/run_one.py lstm 2>&1 | tail -1 | tee -a gpurun_out/run20_lstm_timing.log
done
echo "== trace (kernel 4, serial and pipelined)"
TB200_LSTM_KERNEL=4 TB200_LSTM_PIPE=0 TB200_LSTM_TRACE=gpurun_out/lstm_h16_trace_serial.bin timeout 300 python tools/run_one.py lstm > /dev/null 2>&1
python tools/lstm_trace.py gpurun_out/lstm_h16_trace_serial.bin 100 h16 2>&1 | tee gpurun_out/run20_lstm_h16_timeline_serial.txt
TB200_LSTM_KERNEL=4 TB200_LSTM_PIPE=1 TB200_LSTM_TRACE=gpurun_out/lstm_h16_trace_pipe.bin timeout 300 python tools/run_one.py lstm > /dev/null 2>&1
python tools/lstm_trace.py gpurun_out/lstm_h16_trace_pipe.bin 100 h16 2>&1 | tee gpurun_out/run20_lstm_h16_timeline_pipe.txt
rm -f gpurun_out/*.bin
echo "== full gpu suite"; timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/run20_pytest_all.log
