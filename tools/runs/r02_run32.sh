#!/bin/bash
# r02 run 32 (2 GPUs): multi-GPU sanity at the round's final state — tests/test_gpu_multi.py (dist_check parity, the reference's
# heat2d_dist wrapper, generated distributed functions over NCCL), bench.py at N=2 (ours), host<->device bandwidth with both
# ranks copying at the same time (what bounds the e2e legs at N > 1)
set -u
mkdir -p gpurun_out
echo "== pytest multi"; timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r02_run32_pytest_multi_2gpu.log
echo "== pcie probe, 1 rank then 2 ranks at once"
timeout 120 python tools/pcie_probe.py 2>&1 | tail -1 | tee gpurun_out/r02_run32_pcie_probe.jsonl
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29731 tools/pcie_probe.py 2>&1 | grep "^{" | tee -a gpurun_out/r02_run32_pcie_probe.jsonl
echo "== bench N=2 (ours)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29732 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_run32_bench_n2.json 2> gpurun_out/r02_run32_bench_n2.err
echo "rc=$?"; tail -2 gpurun_out/r02_run32_bench_n2.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02_run32_bench_n2.json").read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("value", "ms_per_step", "n_gpus", "halo", "clocks", "comm_nranks_ok")}, d["e2e"]["value"])
    for k, v in (d.get("configs") or {}).items():
        print(k, {kk: vv for kk, vv in v.items() if kk in ("value", "ms", "error")}, (v.get("e2e") or {}).get("value"))
except Exception as e:
    print("parse failed", e)
PY
