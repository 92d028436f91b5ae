#!/bin/bash
# r01 run 32: full -m gpu suite (incl. front-end tests) + smoke + bench line at the current head
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
tail -4 gpurun_out/build.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest gpu"; timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/run32_pytest.log
echo "== bench"; timeout 900 python bench.py > gpurun_out/run32_bench.json 2> gpurun_out/run32_bench.err; head -c 1600 gpurun_out/run32_bench.json; echo; tail -3 gpurun_out/run32_bench.err
