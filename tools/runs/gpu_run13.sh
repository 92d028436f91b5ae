#!/bin/bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
echo "== lstm tests" ; timeout 900 python -m pytest tests/test_gpu_dnn.py -m gpu -q -k lstm -x 2>&1 | tail -5 | tee gpurun_out/test_lstm.log
echo "== lstm kernel=3"; TB200_LSTM_KERNEL=3 timeout 300 python tools/run_one.py lstm 2>&1 | tail -1
echo "== lstm trace"; TB200_LSTM_TRACE=gpurun_out/lstm_trace2.bin TB200_LSTM_KERNEL=3 timeout 300 python tools/run_one.py lstm 2>&1 | tail -1
python tools/lstm_trace.py gpurun_out/lstm_trace2.bin 100 | head -12
