#!/bin/bash
# r02 run 35 (equal row blocks; otherwise run 34) (1 GPU): generated accumulations with the register tile's accumulators in registers (conv layer, VGG block):
# device tests + timings + launch list
set -u
mkdir -p gpurun_out
echo "== pytest gpu (generated DNN layers)"; timeout 900 python -m pytest tests/test_frontend.py -m gpu -q -k "conv_layer or vgg_block or conv_block" 2>&1 | tail -4 | tee gpurun_out/r02_run35_pytest.log
echo "== generated DNN layers"; timeout 600 python tools/gen_dnn_layers_timing.py 2>&1 | tail -4 | tee gpurun_out/r02_run35_gen_dnn_layers.log
TIRAMISU_B200_NOOP=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_run35_gen_dnn_launches.csv python tools/gen_dnn_layers_timing.py > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r02_run35_gen_dnn_launches.csv")) if len(r) > 10 and r[0].isdigit()]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    name = r[4].split("(")[0]
    agg[name][0] += 1; agg[name][1] += float(r[-1].replace(",", "")) if r[-1] else 0.0
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:6]:
    print(f"{k[-60:]:60s} launches {n:4d}  avg {t/n/1e3:10.1f} us")
PY
