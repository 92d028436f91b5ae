#!/bin/bash
# r02 run 15 (8 GPUs): root-push scatter + flag fan-out kernels + SM re-layout: parity (dist_check), timeline, overlap probe
set -u
mkdir -p gpurun_out
echo "== dist_check (8 ranks)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 tools/dist_check.py > gpurun_out/r02_run15_dist_check_8gpu.log 2>&1
echo "rc=$?"; grep -c " ok" gpurun_out/r02_run15_dist_check_8gpu.log; grep -v " ok$" gpurun_out/r02_run15_dist_check_8gpu.log | grep -v "^\*\|OMP_NUM\|^$" | tail -12
echo "== timeline"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29741 tools/bcast_trace.py 16384 s4 f4 > gpurun_out/r02_run15_bcast_trace.log 2>&1
echo rc=$?; grep "rank 0\]\|rank 1\]" gpurun_out/r02_run15_bcast_trace.log | grep -v "p[123] " | head -60
echo "== probe"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29721 tools/sgemm_bcast_probe.py 8192 16384 2>&1 | grep -v "^\*\|OMP_NUM" > gpurun_out/r02_run15_sgemm_bcast_probe.jsonl
python - <<'PY'
import json
for line in open('gpurun_out/r02_run15_sgemm_bcast_probe.jsonl'):
    line = line.strip()
    if not line.startswith('{'):
        continue
    d = json.loads(line)
    print(d['max_over_ranks_ms'])
PY
