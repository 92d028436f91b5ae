#!/bin/bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
for ch in 1 2 4; do echo "== lstm kernel=3 chunks=$ch"; TB200_LSTM_DBG=$((ch*16)) TB200_LSTM_KERNEL=3 timeout 300 python tools/run_one.py lstm 2>&1 | tail -1 | cut -c1-120; done
