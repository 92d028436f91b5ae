#!/bin/bash
# r01 run 24 (session 6): front-end / sm_100a lowering tests on the box, then the full -m gpu suite and a bench line
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
tail -4 gpurun_out/build.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== frontend gpu tests"; timeout 1200 python -m pytest tests/test_frontend.py -m gpu -q 2>&1 | tail -25 | tee gpurun_out/run24_frontend.log
echo "== sgemm gpu generators timing (reference wrapper, testN=5)"
for g in sgemm_gpu_generator1 sgemm_gpu_generator2 sgemm_gpu_generator3; do
  ( cd /tmp && timeout 300 $GRAFT_REPO_ROOT/tests/compat/_bin/gen_$g 5 0 3 2>&1 | tail -4 ) | tee -a gpurun_out/run24_sgemm_generators.log
done
echo "== pytest gpu (rest)"; timeout 1800 python -m pytest tests -m gpu -q -x --deselect tests/test_frontend.py 2>&1 | tail -8 | tee gpurun_out/run24_pytest.log
echo "== bench"; timeout 900 python bench.py > gpurun_out/run24_bench.json 2> gpurun_out/run24_bench.err; tail -c 1500 gpurun_out/run24_bench.json
