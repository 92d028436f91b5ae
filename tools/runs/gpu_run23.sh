#!/bin/bash
# r01 run 23 (session 5 re-entry): full -m gpu suite, bench line, LSTM engine timings + timeline, smoke
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/run23_pytest.log
for cfg in "3 1" "4 0" "4 1"; do
  set -- $cfg
  echo "== lstm timing kernel=$1 pipe=$2"
  TB200_LSTM_KERNEL=$1 TB200_LSTM_PIPE=$2 timeout 300 python tools/run_one.py lstm 2>&1 | tail -1 | tee -a gpurun_out/run23_lstm_timing.log
done
echo "== trace (kernel 4, serial and pipelined)"
TB200_LSTM_KERNEL=4 TB200_LSTM_PIPE=0 TB200_LSTM_TRACE=gpurun_out/lstm_h16_trace_serial.bin timeout 300 python tools/run_one.py lstm > /dev/null 2>&1
python tools/lstm_trace.py gpurun_out/lstm_h16_trace_serial.bin 100 h16 2>&1 | tee gpurun_out/run23_lstm_h16_timeline_serial.txt
TB200_LSTM_KERNEL=4 TB200_LSTM_PIPE=1 TB200_LSTM_TRACE=gpurun_out/lstm_h16_trace_pipe.bin timeout 300 python tools/run_one.py lstm > /dev/null 2>&1
python tools/lstm_trace.py gpurun_out/lstm_h16_trace_pipe.bin 100 h16 2>&1 | tee gpurun_out/run23_lstm_h16_timeline_pipe.txt
rm -f gpurun_out/*.bin
echo "== bench"; timeout 900 python bench.py > gpurun_out/run23_bench.json 2> gpurun_out/run23_bench.err; tail -c 3000 gpurun_out/run23_bench.json
