#!/bin/bash
# r01 run 29: ncu evidence — launch list of the bench command, per-kernel device time / DRAM traffic of
# generated kernels (generic heat2d pair, the reference's hand-scheduled GEMM generator1, chain_mm pattern), ncu --set full of
# the generated heat2d kernel
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
echo "== launch list of bench.py"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/run29_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-others --no-cpu > gpurun_out/run29_bench_under_ncu.log 2>&1
grep -c tiramisu_sm100 gpurun_out/run29_bench_launches.csv
cd /tmp
echo "== generated heat2d (generic kernels), wrapper_heat2d 10000^2"
timeout 900 ncu --metrics $M --clock-control none -k regex:tiramisu_sm100 -c 6 --csv --log-file $GRAFT_REPO_ROOT/gpurun_out/run29_gen_heat2d_generic.csv $GRAFT_REPO_ROOT/tests/compat/_bin/gen_heat2d_generic > /dev/null 2>&1
echo "== generated heat2d (pattern -> runtime kernel)"
timeout 900 ncu --metrics $M --clock-control none -k regex:tiramisu_sm100 -c 6 --csv --log-file $GRAFT_REPO_ROOT/gpurun_out/run29_gen_heat2d_pattern.csv $GRAFT_REPO_ROOT/tests/compat/_bin/gen_heat2d > /dev/null 2>&1
echo "== sgemm/gpu/generator1 (faithful lowering) and chain_mm (pattern)"
timeout 900 ncu --metrics $M --clock-control none -k regex:tiramisu_sm100 -c 4 --csv --log-file $GRAFT_REPO_ROOT/gpurun_out/run29_gen_generator1.csv $GRAFT_REPO_ROOT/tests/compat/_bin/gen_sgemm_gpu_generator1 2 0 1 > /dev/null 2>&1
timeout 900 ncu --metrics $M --clock-control none -k regex:tiramisu_sm100 -c 8 --csv --log-file $GRAFT_REPO_ROOT/gpurun_out/run29_gen_chain_mm.csv $GRAFT_REPO_ROOT/tests/compat/_bin/gen_chain_mm 2 0 > /dev/null 2>&1
echo "== ncu --set full: generated generic heat2d_s1 kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:heat2d_s1 -s 1 -c 1 -o $GRAFT_REPO_ROOT/gpurun_out/prof_gen_heat2d_s1 $GRAFT_REPO_ROOT/tests/compat/_bin/gen_heat2d_generic > /dev/null 2>&1
ls -la $GRAFT_REPO_ROOT/gpurun_out/ | tail -12
