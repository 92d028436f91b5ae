#!/bin/bash
# r02 run 39 (1 GPU): warp_affinegpu generator on the device + the other *gpu Halide benchmarks
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_frontend_halide_gpu.py -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r02_run39_pytest_halide_gpu.log
