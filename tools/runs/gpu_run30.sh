#!/bin/bash
# r01 run 30: re-targeted CPU schedules after coarsening (8 rows per thread, 256-thread blocks): parity + kernel times
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
echo "== frontend gpu tests"; timeout 1500 python -m pytest tests/test_frontend.py -m gpu -q 2>&1 | tail -8 | tee gpurun_out/run30_frontend.log
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
cd /tmp
timeout 900 ncu --metrics $M --clock-control none -k regex:tiramisu_sm100 -c 6 --csv --log-file $GRAFT_REPO_ROOT/gpurun_out/run30_gen_heat2d_generic.csv $GRAFT_REPO_ROOT/tests/compat/_bin/gen_heat2d_generic > /dev/null 2>&1
grep -o '"tiramisu_sm100[^"]*".*"gpu__time_duration.sum","ns","[0-9]*"' $GRAFT_REPO_ROOT/gpurun_out/run30_gen_heat2d_generic.csv | sed 's/","[^t].*"ns","/ ns=/' | head -6
timeout 900 ncu --set full --clock-control none --import-source on -k regex:heat2d_s1 -s 1 -c 1 -o $GRAFT_REPO_ROOT/gpurun_out/prof_gen_heat2d_s1_v2 $GRAFT_REPO_ROOT/tests/compat/_bin/gen_heat2d_generic > /dev/null 2>&1
