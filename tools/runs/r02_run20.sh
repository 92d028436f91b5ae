#!/bin/bash
# r02 run 20 (1 GPU): ncu launch list of the bench command (headline step) and ncu --set full of the two stencil kernels at the
# round's state (the blur kernel changed: 128-bit stores on both row parities) -> profiles/r02_*; halide gpu kernel times
set -u
mkdir -p gpurun_out
echo "== ncu launch list (bench headline, 2 steps)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_run20_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-others --no-cpu > gpurun_out/r02_run20_bench_under_ncu.log 2>&1
grep -c "tiramisu_sm100" gpurun_out/r02_run20_bench_launches.csv
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r02_run20_bench_launches.csv")) if len(r) > 10 and r[0].isdigit()]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    name = r[4].split("(")[0]
    agg[name][0] += 1; agg[name][1] += float(r[-1].replace(",", "")) if r[-1] else 0.0
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:8]:
    print(f"{k[:70]:70s} launches {n:4d}  total {t/1e3:10.1f} us  avg {t/n/1e3:8.1f} us")
PY
for k in blur heat2d; do
  echo "== ncu --set full $k"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$( [ $k = blur ] && echo blur_f32 || echo heat2d ) -c 1 -o gpurun_out/prof_r02_$k python tools/run_one.py $k > gpurun_out/ncu_$k.log 2>&1; tail -1 gpurun_out/ncu_$k.log
done
ls -la gpurun_out/*.ncu-rep
