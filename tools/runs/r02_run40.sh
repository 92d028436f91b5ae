#!/bin/bash
# r02 run 40 (1 GPU): tests/test_frontend.py on the device at the round's last front-end commit
set -u
mkdir -p gpurun_out
timeout 330 python -m pytest tests/test_frontend.py -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r02_run40_pytest_frontend_gpu.log
