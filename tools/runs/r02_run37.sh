#!/bin/bash
# r02 run 37 (1 GPU): blurxy_tiramisu.cpp through the pattern (runtime blur kernel) on the device + the stencil / halide front-end tests
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_frontend_halide_gpu.py tests/test_gpu_stencil.py -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r02_run37_pytest.log
timeout 300 python -m pytest tests/test_frontend.py -m gpu -q -k "blur or heat2d" 2>&1 | tail -3 | tee -a gpurun_out/r02_run37_pytest.log
