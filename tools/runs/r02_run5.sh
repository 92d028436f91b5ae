#!/bin/bash
# r02 run 5: thread-axis coalescing in the lowering (all front-end GPU tests + Halide gpu kernel times), LSTM gpu_lib / cpu_lib ABIs
set -u
mkdir -p gpurun_out
echo "== frontend + halide gpu + dnn tests"; timeout 1500 python -m pytest tests/test_frontend.py tests/test_frontend_halide_gpu.py tests/test_gpu_dnn.py tests/test_compat_wrappers.py -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r02_run5_pytest.log
echo "== halide gpu kernel times"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_run5_halide_gpu_launches.csv python tools/halide_gpu_run.py 2>&1 | tail -2
python tools/halide_gpu_run.py --report gpurun_out/r02_run5_halide_gpu_launches.csv | python -c "
import json,sys
d=json.load(sys.stdin)
for k,v in d.items(): print(k.split('(')[0][-60:], round(v['us'],1),'us', round(v['GBps']),'GB/s', round(v['frac_of_hbm_peak'],3))
" | tee gpurun_out/r02_run5_halide_gpu_report.txt
