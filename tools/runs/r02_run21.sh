#!/bin/bash
# r02 run 21 (1 GPU): stream-K in the persistent GEMM kernels: parity tests, timing sweep with / without; then the ncu captures of run 20
set -u
mkdir -p gpurun_out
echo "== gemm tests"; timeout 900 python -m pytest tests/test_gpu_gemm.py -q -x 2>&1 | tail -8 | tee gpurun_out/r02_run21_pytest_gemm.log
echo "== gemm sweep (default = stream-K where the last wave is partly filled)"
timeout 600 python tools/gemm_exp.py 2048 4096 8192 2>&1 | cut -c1-400 | tee gpurun_out/r02_run21_gemm_exp.jsonl
echo "== the same without stream-K"
TB200_GEMM_NO_STREAMK=1 timeout 600 python tools/gemm_exp.py 2048 4096 8192 2>&1 | cut -c1-400 | tee gpurun_out/r02_run21_gemm_exp_nostreamk.jsonl
bash tools/runs/r02_run20.sh
