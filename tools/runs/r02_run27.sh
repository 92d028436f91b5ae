#!/bin/bash
# r02 run 27 (run 26 with plain full-sector h stores) (1 GPU): LSTM fp16 engine with h exchanged IN THE DATA (h16ll: sentinel-polled operand fetch by the epilogue threads,
# no fence / counter on the critical path): parity, timing, timeline
set -u
mkdir -p gpurun_out
echo "== lstm tests"; timeout 900 python -m pytest tests/test_gpu_dnn.py -q -x -k lstm 2>&1 | tail -12 | tee gpurun_out/r02_run27_pytest_lstm.log
for v in 1 0; do
  echo "== lstm timing ll=$v (pipe 1)"; TB200_LSTM_LL=$v timeout 300 python tools/run_one.py lstm 2>&1 | tail -1 | tee -a gpurun_out/r02_run27_lstm_timing.log
  echo "== lstm timing ll=$v (pipe 0)"; TB200_LSTM_LL=$v TB200_LSTM_PIPE=0 timeout 300 python tools/run_one.py lstm 2>&1 | tail -1 | tee -a gpurun_out/r02_run27_lstm_timing.log
done
echo "== lstm timeline (serial, ll)"; TB200_LSTM_PIPE=0 TB200_LSTM_TRACE=gpurun_out/lstm_h16ll_trace_serial.bin timeout 300 python tools/run_one.py lstm > /dev/null 2>&1
python tools/lstm_trace.py gpurun_out/lstm_h16ll_trace_serial.bin 100 h16 2>&1 | tee gpurun_out/r02_run27_lstm_h16ll_timeline_serial.txt
echo "== lstm timeline (pipe, ll)"; TB200_LSTM_TRACE=gpurun_out/lstm_h16ll_trace_pipe.bin timeout 300 python tools/run_one.py lstm > /dev/null 2>&1
python tools/lstm_trace.py gpurun_out/lstm_h16ll_trace_pipe.bin 100 h16 2>&1 | tee gpurun_out/r02_run27_lstm_h16ll_timeline_pipe.txt
python - <<'PY'
import numpy as np
a = np.fromfile("gpurun_out/lstm_h16ll_trace_serial.bin", dtype=np.uint64).reshape(-1, 100, 8)
print("threads that patched a chunk after the bulk fetch (layer 0, 99 steps, 64 CTAs x 256 threads):", int(a[:, 0, 0].sum()))
PY
