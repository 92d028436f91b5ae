#!/bin/bash
# First GPU round-trip: parity tests (separate processes so a trap in one suite cannot poison the
# next), tcgen05 GEMM diagnostics, bench line, ncu launch list + one full capture of the top kernel.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
echo "== stencil tests" ; timeout 900 python -m pytest tests/test_gpu_stencil.py -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/test_stencil.log
echo "== gemm diag" ; timeout 600 python tools/gemm_diag.py > gpurun_out/gemm_diag.log 2>&1 ; tail -5 gpurun_out/gemm_diag.log
echo "== gemm tests" ; timeout 900 python -m pytest tests/test_gpu_gemm.py -m gpu -q 2>&1 | tail -40 | tee gpurun_out/test_gemm.log
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== bench" ; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err ; cat gpurun_out/bench.json ; tail -3 gpurun_out/bench.err
echo "== ncu launches" ; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-others > gpurun_out/ncu_bench.log 2>&1 ; tail -2 gpurun_out/ncu_bench.log
echo "== ncu full heat2d" ; timeout 600 ncu --set full --clock-control none --import-source on -k regex:heat2d_v4 -s 3 -c 1 -o gpurun_out/prof_heat2d -f python bench.py --steps 2 --warmup 3 --no-cpu --no-others > gpurun_out/ncu_full.log 2>&1 ; tail -2 gpurun_out/ncu_full.log
echo "== ncu full blur" ; timeout 600 ncu --set full --clock-control none --import-source on -k regex:blur_f32_v4 -s 3 -c 1 -o gpurun_out/prof_blur -f python bench.py --steps 2 --warmup 3 --no-cpu --no-others > gpurun_out/ncu_full2.log 2>&1 ; tail -2 gpurun_out/ncu_full2.log
ls -la gpurun_out
