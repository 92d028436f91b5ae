#!/bin/bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
G=$(nvidia-smi -L | wc -l)
echo "== dist check ($G ranks)"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29701 tools/dist_check.py 2>&1 | grep -E "rank|DIST" | sort | tee gpurun_out/dist_check.log
echo "== pytest multi"; timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -3
for mode in nccl sag; do
echo "== sgemm sweep N=$G bcast=$mode"; TB200_SWEEP_BCAST=$mode timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29702 tools/bench_sgemm_sweep.py 8192 16384 2> gpurun_out/sweep_$mode.err | grep '^{' | tee gpurun_out/sweep_n${G}_$mode.jsonl
done
for p in 8 16; do
echo "== sgemm sweep N=$G bcast=sag panels=$p"; TB200_SWEEP_PANELS=$p timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29703 tools/bench_sgemm_sweep.py 16384 2>> gpurun_out/sweep_sag.err | grep '^{' | tee -a gpurun_out/sweep_n${G}_sag.jsonl
done
