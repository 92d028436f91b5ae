#!/bin/bash
# r02 run 7 (2 GPUs): fused B broadcast column-panel form (dist_check 4b), heat2d_dist wrapper protocol,
# bench.py exactly as the driver launches it at N=2 (both arms), sgemm sweep line with the DMA-engine broadcast
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | head -3; nvidia-smi topo -m 2>/dev/null | head -6
echo "== dist_check (2 ranks)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 tools/dist_check.py > gpurun_out/r02_run7_dist_check_2gpu.log 2>&1
echo "rc=$?"; grep -c " ok" gpurun_out/r02_run7_dist_check_2gpu.log; grep -v " ok$" gpurun_out/r02_run7_dist_check_2gpu.log | tail -15
echo "== bench N=2 (ours)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_run7_bench_n2.json 2> gpurun_out/r02_run7_bench_n2.err
echo "rc=$?"; tail -3 gpurun_out/r02_run7_bench_n2.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02_run7_bench_n2.json").read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("value", "ms_per_step", "n_gpus", "halo", "clocks", "comm_nranks_ok")}, d["e2e"]["value"])
    for k, v in (d.get("configs") or {}).items():
        print(k, {kk: vv for kk, vv in v.items() if kk in ("value", "ms", "ms_gemm_only", "ms_bcast_only", "bcast_GBps_per_rank", "ms_ncclBroadcast_form", "ms_k_panel_form", "error")}, (v.get("e2e") or {}).get("value"))
except Exception as e:
    print("parse failed", e)
PY
