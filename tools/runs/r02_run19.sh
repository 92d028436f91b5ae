#!/bin/bash
# r02 run 19 (1 GPU): full -m gpu suite at the round's state (staged-contraction pattern, host-loop CUDA graphs, gated GEMM,
# distribution lowering), the compiled LSTM generator's time with the graph replay, the reference's GPU sgemm generators
# through both lowerings, bench.py (both arms) as the driver runs it
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest gpu"; timeout 1700 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 | tee gpurun_out/r02_run19_pytest_gpu.log
echo "== lowered LSTM generator (config 4), graph replay vs direct"
timeout 300 python tools/lstm_gen_config4.py 2>&1 | tail -3 | tee gpurun_out/r02_run19_lstm_gen_timing.log
TB200_NO_HOST_GRAPH=1 timeout 300 python tools/lstm_gen_config4.py 2>&1 | tail -3 | tee -a gpurun_out/r02_run19_lstm_gen_timing.log
echo "== reference sgemm/gpu generators (3072^3): pattern vs faithful"
for g in 1 2 3; do for v in "" _generic; do
  exe=tests/compat/_bin/gen_sgemm_gpu_generator$g$v
  [ -x $exe ] && (cd /tmp && timeout 120 /root/repo/$exe 5 1 2>&1 | grep -i "tiramisu\|cublas\|match" | tail -4 | sed "s/^/gen$g$v: /")
done; done | tee gpurun_out/r02_run19_sgemm_gpu_generators.log
echo "== bench (ours)"; timeout 900 python bench.py > gpurun_out/r02_run19_bench.json 2> gpurun_out/r02_run19_bench.err; tail -2 gpurun_out/r02_run19_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_run19_bench.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "clocks", "gpu_launches")}, d["e2e"]["value"], d["roofline"]["frac"], d["cpu_baseline"]["value"])
for k, v in (d.get("configs") or {}).items():
    print(k, {kk: vv for kk, vv in v.items() if kk in ("value", "ms", "error", "cublas_tf32_GFLOPs_same_run")}, v.get("roofline", {}).get("frac"), (v.get("e2e") or {}).get("value"))
PY
echo "== bench (reference arm)"; timeout 600 python bench.py --impl reference > gpurun_out/r02_run19_bench_ref.json 2>> gpurun_out/r02_run19_bench.err; head -c 600 gpurun_out/r02_run19_bench_ref.json
