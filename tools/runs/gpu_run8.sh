#!/bin/bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
echo "== lstm tests" ; timeout 900 python -m pytest tests/test_gpu_dnn.py -m gpu -q -k lstm -x 2>&1 | tail -5 | tee gpurun_out/test_lstm.log
echo "== lstm kernel=3"; TB200_LSTM_KERNEL=3 timeout 300 python tools/run_one.py lstm 2>&1 | tail -1
NCU="ncu --set full --clock-control none --import-source on -f"
echo "== ncu lstm ksplit" ; TB200_LSTM_KERNEL=3 timeout 900 $NCU -k regex:lstm_layer_ksplit -s 1 -c 1 -o gpurun_out/prof_lstm_ksplit2 python tools/run_one.py lstm > gpurun_out/ncu_lstm.log 2>&1 ; tail -1 gpurun_out/ncu_lstm.log
