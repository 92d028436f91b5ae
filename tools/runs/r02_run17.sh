#!/bin/bash
# r02 run 17 (8 GPUs): root-push + double-buffered staging: parity at 8 ranks, overlap probe, bench.py as the driver launches it at N=8
set -u
mkdir -p gpurun_out
echo "== dist_check (8 ranks)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 tools/dist_check.py > gpurun_out/r02_run17_dist_check_8gpu.log 2>&1
echo "rc=$?"; grep -c " ok" gpurun_out/r02_run17_dist_check_8gpu.log; grep -v " ok$" gpurun_out/r02_run17_dist_check_8gpu.log | grep -v "^\*\|OMP_NUM\|^$" | tail -12
echo "== probe"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29721 tools/sgemm_bcast_probe.py 8192 16384 2>&1 | grep -v "^\*\|OMP_NUM" > gpurun_out/r02_run17_sgemm_bcast_probe.jsonl
python - <<'PY'
import json
for line in open('gpurun_out/r02_run17_sgemm_bcast_probe.jsonl'):
    line = line.strip()
    if line.startswith('{'):
        print(json.loads(line)['max_over_ranks_ms'])
    elif 'rror' in line:
        print(line[:300])
PY
echo "== bench N=8 (ours)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_run17_bench_n8.json 2> gpurun_out/r02_run17_bench_n8.err
echo "rc=$?"; tail -3 gpurun_out/r02_run17_bench_n8.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02_run17_bench_n8.json").read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("value", "ms_per_step", "n_gpus", "clocks", "comm_nranks_ok")}, d["e2e"]["value"])
    for k, v in (d.get("configs") or {}).items():
        print(k, {kk: vv for kk, vv in v.items() if kk in ("value", "ms", "ms_gemm_only", "ms_bcast_only", "bcast_GBps_per_rank", "ms_ncclBroadcast_form", "error")}, (v.get("e2e") or {}).get("value"))
except Exception as e:
    print("parse failed", e)
PY
