#!/bin/bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
G=$(nvidia-smi -L | wc -l)
echo "GPUs: $G"
for n in 2 4 8; do
if [ "$G" -ge "$n" ]; then
echo "== bench N=$n"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --steps 20 --warmup 5 --no-others 2> gpurun_out/bench_n$n.err | grep '^{' > gpurun_out/bench_n$n.json; python -c "
import sys,json; d=json.load(open('gpurun_out/bench_n$n.json')); print({k:d[k] for k in ('value','ms_per_step','n_gpus','kernels','e2e')})"; grep -iE "error|assert" gpurun_out/bench_n$n.err | head -5
fi
done
if [ "$G" -ge 8 ]; then
echo "== sgemm sweep N=8"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29628 tools/bench_sgemm_sweep.py 4096 8192 16384 2> gpurun_out/sweep_n8.err | tee gpurun_out/sweep_n8.jsonl
fi
