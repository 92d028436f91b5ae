#!/bin/bash
# r02 run 30 (1 GPU): the round's final state — full -m gpu suite, generated DNN layers (register tile kept inside the
# reduction), ncu --set full of the LSTM layer kernel with R in tensor memory and of the persistent GEMM with the last-wave
# split (4096^3), bench.py (both arms) as the driver runs it
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest gpu"; timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee gpurun_out/r02_run30_pytest_gpu.log
echo "== generated DNN layers"; timeout 600 python tools/gen_dnn_layers_timing.py 2>&1 | tail -4 | tee gpurun_out/r02_run30_gen_dnn_layers.log
echo "== ncu lstm_layer_h16 (one layer at a time: ncu serialises kernels)"
TB200_LSTM_KERNEL=4 TB200_LSTM_PIPE=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_layer_h16 -c 1 \
  -o gpurun_out/prof_r02_lstm_h16_rtmem python tools/run_one.py lstm > gpurun_out/ncu_lstm.log 2>&1; tail -1 gpurun_out/ncu_lstm.log
echo "== ncu persistent GEMM 4096^3"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32_persistent -c 1 \
  -o gpurun_out/prof_r02_gemm4096_split python tools/run_one.py sgemm4096 > gpurun_out/ncu_gemm.log 2>&1; tail -1 gpurun_out/ncu_gemm.log
echo "== bench (ours)"; timeout 900 python bench.py > gpurun_out/r02_run30_bench.json 2> gpurun_out/r02_run30_bench.err; tail -2 gpurun_out/r02_run30_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_run30_bench.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "clocks", "gpu_launches")}, d["e2e"]["value"], d["roofline"]["frac"], d["cpu_baseline"]["value"])
for k, v in (d.get("configs") or {}).items():
    print(k, {kk: vv for kk, vv in v.items() if kk in ("value", "ms", "error", "cublas_tf32_GFLOPs_same_run")}, v.get("roofline", {}).get("frac"), (v.get("e2e") or {}).get("value"))
PY
echo "== bench (reference arm)"; timeout 600 python bench.py --impl reference > gpurun_out/r02_run30_bench_ref.json 2>> gpurun_out/r02_run30_bench.err; head -c 400 gpurun_out/r02_run30_bench_ref.json
