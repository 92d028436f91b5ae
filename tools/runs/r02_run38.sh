#!/bin/bash
# r02 run 38 (1 GPU): front-end GPU tests at the final commit (own register-blocking / pieces generators on the device)
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_frontend.py tests/test_frontend_halide_gpu.py -m gpu -q 2>&1 | tail -5 | tee gpurun_out/r02_run38_pytest_frontend_gpu.log
