#!/bin/bash
# r02 run 36 (1 GPU): the round's final commit — smoke, full -m gpu suite, bench.py both arms
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest gpu"; timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -10 | tee gpurun_out/r02_run36_pytest_gpu.log
echo "== bench (ours)"; timeout 900 python bench.py > gpurun_out/r02_run36_bench.json 2> gpurun_out/r02_run36_bench.err; tail -2 gpurun_out/r02_run36_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_run36_bench.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "clocks", "gpu_launches")}, d["e2e"]["value"], d["roofline"]["frac"], d["cpu_baseline"]["value"])
for k, v in (d.get("configs") or {}).items():
    print(k, {kk: vv for kk, vv in v.items() if kk in ("value", "ms", "error", "cublas_tf32_GFLOPs_same_run")}, v.get("roofline", {}).get("frac"), (v.get("e2e") or {}).get("value"))
PY
echo "== bench (reference arm)"; timeout 600 python bench.py --impl reference > gpurun_out/r02_run36_bench_ref.json 2>> gpurun_out/r02_run36_bench.err; head -c 300 gpurun_out/r02_run36_bench_ref.json
