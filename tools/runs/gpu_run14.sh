#!/bin/bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
echo "== lstm tests" ; timeout 900 python -m pytest tests/test_gpu_dnn.py -m gpu -q -k lstm -x 2>&1 | tail -3
for ch in 1 2 4; do echo "== lstm kernel=3 chunks=$ch"; TB200_LSTM_DBG=$((ch*16)) TB200_LSTM_KERNEL=3 timeout 300 python tools/run_one.py lstm 2>&1 | tail -1 | cut -c1-120; done
echo "== lstm trace (chunks=2)"; TB200_LSTM_DBG=32 TB200_LSTM_TRACE=gpurun_out/lstm_trace3.bin TB200_LSTM_KERNEL=3 timeout 300 python tools/run_one.py lstm > /dev/null 2>&1
python tools/lstm_trace.py gpurun_out/lstm_trace3.bin 100 2>/dev/null | sed -n 10,19p
