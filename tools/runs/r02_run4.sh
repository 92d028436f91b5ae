#!/bin/bash
# r02 run 4: gpu Halide benchmarks through the front-end (parity + kernel times), blur with 128-bit stores (tests + bench), stencil tests
set -u
mkdir -p gpurun_out
echo "== halide gpu tests"; timeout 900 python -m pytest tests/test_frontend_halide_gpu.py -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r02_run4_pytest_halide.log
echo "== stencil tests"; timeout 900 python -m pytest tests/test_gpu_stencil.py -q 2>&1 | tail -5 | tee gpurun_out/r02_run4_pytest_stencil.log
echo "== halide gpu kernel times"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_run4_halide_gpu_launches.csv python tools/halide_gpu_run.py 2>&1 | tail -2
python tools/halide_gpu_run.py --report gpurun_out/r02_run4_halide_gpu_launches.csv | tee gpurun_out/r02_run4_halide_gpu_report.json
echo "== bench (headline only)"; timeout 600 python bench.py --no-others > gpurun_out/r02_run4_bench.json 2> gpurun_out/r02_run4_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_run4_bench.json"))
print(d["value"], d["ms_per_step"], json.dumps(d["kernels"])[:900]); print(d["clocks"], d["e2e"]["value"])
PY
tail -3 gpurun_out/r02_run4_bench.err
