#!/bin/bash
# r01 run 26: front-end tests incl. the re-targeted CPU schedules and the reference heat2d generator; smoke; e2e timing of
# the generated heat2d_tiramisu against the hand-registered one through the reference wrapper
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
tail -5 gpurun_out/build.log
echo "== frontend gpu tests"; timeout 1500 python -m pytest tests/test_frontend.py -m gpu -q 2>&1 | tail -25 | tee gpurun_out/run26_frontend.log
echo "== wrapper_heat2d: generated (pattern), generated (generic kernels), hand-registered"
for b in gen_heat2d gen_heat2d_generic wrapper_heat2d; do ( cd /tmp && timeout 600 $GRAFT_REPO_ROOT/tests/compat/_bin/$b 2>&1 | tail -3 ) | tee -a gpurun_out/run26_heat2d_wrappers.log; done
