#!/bin/bash
# r02 run 12 (8 GPUs): every sharded path against the oracle at 8 ranks (dist_check: boxblur, heat2d NCCL + peer halos, 12 Jacobi
# sweeps, heat2d_dist_tiramisu ABI, MPI shims, fused sgemm broadcast), the reference's heat2d_dist wrapper at 4 ranks, the fused
# sgemm overlap probe, and bench.py exactly as the driver launches it at N=8
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | sort | uniq -c
echo "== pytest multi (dist_check at all GPUs + reference wrapper_heat2d_dist at 4 ranks)"
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r02_run12_pytest_multi.log
echo "== dist_check log (8 ranks)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 tools/dist_check.py > gpurun_out/r02_run12_dist_check_8gpu.log 2>&1
echo "rc=$?"; grep -c " ok" gpurun_out/r02_run12_dist_check_8gpu.log; grep -v " ok$" gpurun_out/r02_run12_dist_check_8gpu.log | grep -v "^\*\|OMP_NUM\|^$" | tail -12
echo "== probe"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29721 tools/sgemm_bcast_probe.py 8192 16384 2>&1 | grep -v "^\*\|OMP_NUM" | tee gpurun_out/r02_run12_sgemm_bcast_probe.jsonl | cut -c1-900
echo "== bench N=8 (ours)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_run12_bench_n8.json 2> gpurun_out/r02_run12_bench_n8.err
echo "rc=$?"; tail -3 gpurun_out/r02_run12_bench_n8.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02_run12_bench_n8.json").read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("value", "ms_per_step", "n_gpus", "clocks", "comm_nranks_ok")}, d["e2e"]["value"], d.get("kernels"))
    for k, v in (d.get("configs") or {}).items():
        print(k, {kk: vv for kk, vv in v.items() if kk in ("value", "ms", "ms_gemm_only", "ms_bcast_only", "bcast_GBps_per_rank", "ms_ncclBroadcast_form", "error")}, (v.get("e2e") or {}).get("value"))
except Exception as e:
    print("parse failed", e)
PY
