#!/bin/bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
echo "== compat wrappers"; timeout 1200 python -m pytest tests/test_compat_wrappers.py -m gpu -q 2>&1 | tail -15 | tee gpurun_out/test_compat.log
echo "== sgemm gpu wrapper (benchmark mode: testN=5)"; (cd /tmp && timeout 300 $GRAFT_REPO_ROOT/tests/compat/_bin/sgemm_gpu_wrapper 5 0 3 2>&1 | tail -8) | tee gpurun_out/sgemm_gpu_wrapper.log
echo "== heat2d wrapper"; (cd /tmp && timeout 600 $GRAFT_REPO_ROOT/tests/compat/_bin/wrapper_heat2d 2>&1 | tail -5) | tee gpurun_out/wrapper_heat2d.log
echo "== sgemm cpu wrapper (slow)"; (cd /tmp && time timeout 2400 $GRAFT_REPO_ROOT/tests/compat/_bin/sgemm_wrapper 2>&1 | tail -40) 2>&1 | tee gpurun_out/sgemm_wrapper.log | tail -12
