#!/bin/bash
# r01 run 25 (2 GPUs): peer-memory halos — parity (dist_check incl. 12 Jacobi sweeps with the in-kernel handshake)
# and the bench step with halos read in place vs the NCCL exchange
set -u
mkdir -p gpurun_out
N=${1:-2}
python -c "from tiramisu_b200 import build; build.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
echo "== dist_check ($N ranks)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29741 tools/dist_check.py 2>&1 | grep -v "^W\|^\*\*\*" | grep -v "exitcode\|error_file\|traceback\|time  \|host  " | tail -60 | tee gpurun_out/run25_dist_check_n$N.log
echo "== bench peer ($N ranks)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29742 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/run25_bench_n${N}_peer.json 2> gpurun_out/run25_bench_n${N}_peer.err
tail -c 2500 gpurun_out/run25_bench_n${N}_peer.json; tail -5 gpurun_out/run25_bench_n${N}_peer.err
