#!/bin/bash
# r02 run 28 (1 GPU): full -m gpu suite at the round's state (R in tensor memory for the LSTM, last-wave k-split GEMM, the
# reference's sgemm / conv layer / vgg block / blurxy CPU generators through the front-end), LSTM barrier variant timing,
# generated DNN layer timings, bench.py (both arms) as the driver runs it
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest gpu"; timeout 1700 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 | tee gpurun_out/r02_run28_pytest_gpu.log
for v in 0 3; do echo "== lstm timing sync=$v"; TB200_LSTM_SYNC=$v timeout 300 python tools/run_one.py lstm 2>&1 | tail -1 | tee -a gpurun_out/r02_run28_lstm_timing.log; done
echo "== generated DNN layers"; timeout 600 python tools/gen_dnn_layers_timing.py 2>&1 | tail -4 | tee gpurun_out/r02_run28_gen_dnn_layers.log
echo "== bench (ours)"; timeout 900 python bench.py > gpurun_out/r02_run28_bench.json 2> gpurun_out/r02_run28_bench.err; tail -2 gpurun_out/r02_run28_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_run28_bench.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "clocks", "gpu_launches")}, d["e2e"]["value"], d["roofline"]["frac"], d["cpu_baseline"]["value"])
for k, v in (d.get("configs") or {}).items():
    print(k, {kk: vv for kk, vv in v.items() if kk in ("value", "ms", "error", "cublas_tf32_GFLOPs_same_run")}, v.get("roofline", {}).get("frac"), (v.get("e2e") or {}).get("value"))
PY
echo "== bench (reference arm)"; timeout 600 python bench.py --impl reference > gpurun_out/r02_run28_bench_ref.json 2>> gpurun_out/r02_run28_bench.err; head -c 600 gpurun_out/r02_run28_bench_ref.json
