#!/bin/bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
echo "== stencil tests" ; timeout 900 python -m pytest tests/test_gpu_stencil.py -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/test_stencil.log
echo "== gemm diag" ; timeout 600 python tools/gemm_diag.py > gpurun_out/gemm_diag.log 2>&1 ; grep -E "probe|tB=0" gpurun_out/gemm_diag.log | cut -c1-120 | head -30
echo "== gemm tests" ; timeout 900 python -m pytest tests/test_gpu_gemm.py -m gpu -q 2>&1 | tail -25 | tee gpurun_out/test_gemm.log
echo "== dnn tests" ; timeout 1200 python -m pytest tests/test_gpu_dnn.py -m gpu -q 2>&1 | tail -40 | tee gpurun_out/test_dnn.log
echo "== tf32 peak" ; timeout 300 python tools/run_one.py tf32peak 2>&1 | tail -2
echo "== bench" ; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err ; cat gpurun_out/bench.json ; tail -3 gpurun_out/bench.err
echo "== ncu blur" ; timeout 600 ncu --set full --clock-control none --import-source on -k regex:blur_f32_v4 -s 2 -c 1 -o gpurun_out/prof_blur2 -f python tools/run_one.py blur > gpurun_out/ncu_blur.log 2>&1 ; tail -1 gpurun_out/ncu_blur.log
echo "== ncu sgemm" ; timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32 -s 2 -c 1 -o gpurun_out/prof_sgemm4096 -f python tools/run_one.py sgemm4096 > gpurun_out/ncu_sgemm.log 2>&1 ; tail -1 gpurun_out/ncu_sgemm.log
echo "== ncu conv" ; timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_block_tf32 -s 1 -c 1 -o gpurun_out/prof_conv -f python tools/run_one.py conv > gpurun_out/ncu_conv.log 2>&1 ; tail -1 gpurun_out/ncu_conv.log
echo "== ncu lstm launches" ; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/lstm_launches.csv python tools/run_one.py lstm > gpurun_out/ncu_lstm.log 2>&1 ; tail -1 gpurun_out/ncu_lstm.log
ls gpurun_out
