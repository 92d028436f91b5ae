#!/bin/bash
# r02 run 1: build check, smoke (now with conv + LSTM), full -m gpu suite, LSTM engine timing + timeline
# (block-scaled fp16 weights, per-CTA step flags), ncu --set full of lstm_layer_h16, bench line with all configs
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader; nproc
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 | tee gpurun_out/r02_run1_pytest.log
for cfg in "4 0" "4 1" "3 1"; do
  set -- $cfg
  echo "== lstm timing kernel=$1 pipe=$2"
  TB200_LSTM_KERNEL=$1 TB200_LSTM_PIPE=$2 timeout 300 python tools/run_one.py lstm 2>&1 | tail -1 | tee -a gpurun_out/r02_run1_lstm_timing.log
done
echo "== trace (kernel 4, serial)"
TB200_LSTM_KERNEL=4 TB200_LSTM_PIPE=0 TB200_LSTM_TRACE=gpurun_out/lstm_h16_trace_serial.bin timeout 300 python tools/run_one.py lstm > /dev/null 2>&1
python tools/lstm_trace.py gpurun_out/lstm_h16_trace_serial.bin 100 h16 2>&1 | tee gpurun_out/r02_run1_lstm_h16_timeline_serial.txt
rm -f gpurun_out/*.bin
echo "== ncu lstm_layer_h16 (one layer at a time: ncu serialises kernels)"
TB200_LSTM_KERNEL=4 TB200_LSTM_PIPE=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_layer_h16 -c 1 \
  -o gpurun_out/prof_lstm_h16 python tools/run_one.py lstm > gpurun_out/ncu_lstm.log 2>&1; tail -2 gpurun_out/ncu_lstm.log
echo "== bench"; timeout 900 python bench.py > gpurun_out/r02_run1_bench.json 2> gpurun_out/r02_run1_bench.err; tail -c 6000 gpurun_out/r02_run1_bench.json; tail -5 gpurun_out/r02_run1_bench.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference > gpurun_out/r02_run1_bench_ref.json 2>> gpurun_out/r02_run1_bench.err; head -c 1500 gpurun_out/r02_run1_bench_ref.json
