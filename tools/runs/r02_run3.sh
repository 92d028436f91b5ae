#!/bin/bash
# r02 run 3: split-K GEMM with cached tensor maps: tests, graph-replay timings, ncu --set full; advisor cases on the device
set -u
mkdir -p gpurun_out
echo "== gemm tests"; timeout 900 python -m pytest tests/test_gpu_gemm.py -q 2>&1 | tail -8 | tee gpurun_out/r02_run3_pytest_gemm.log
echo "== frontend advisor cases"; timeout 600 python -m pytest tests/test_frontend.py -q -m gpu -k "shared_by_two or unequal_fuse" 2>&1 | tail -5
echo "== gemm sweep"; timeout 600 python tools/gemm_exp.py 1024 2048 2>&1 | tee gpurun_out/r02_run3_gemm_exp.jsonl
echo "== ncu split-K 1024"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:splitk -c 1 -o gpurun_out/prof_sgemm1024_splitk python tools/run_one.py sgemm1024 > gpurun_out/ncu_sgemm1024.log 2>&1; tail -2 gpurun_out/ncu_sgemm1024.log
echo "== ncu launch list 1024 (gpu__time_duration)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_run3_sgemm1024_launches.csv python tools/run_one.py sgemm1024 > /dev/null 2>&1; tail -5 gpurun_out/r02_run3_sgemm1024_launches.csv
echo "== bench (sgemm only)"; timeout 600 python bench.py --only sgemm_1024,sgemm_2048 --no-cpu > gpurun_out/r02_run3_bench.json 2> gpurun_out/r02_run3_bench.err; python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_run3_bench.json"))
for k, v in (d.get("configs") or {}).items():
    print(k, {kk: vv for kk, vv in v.items() if kk in ("value", "ms", "ms_stream_launched", "ms_graph_replay", "launch", "graph_error", "cublas_tf32_GFLOPs_same_run")}, v.get("roofline", {}).get("frac"))
PY
tail -3 gpurun_out/r02_run3_bench.err
