#!/bin/bash
# r02 run 33 (1 GPU): the conv + ReLU + max-pool block generator without its whole-function idiom (reference wrapper check on
# the device), its time next to the idiom's, and the front-end GPU tests at the final state
set -u
mkdir -p gpurun_out
echo "== pytest gpu (front-end)"; timeout 1700 python -m pytest tests/test_frontend.py tests/test_frontend_halide_gpu.py -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r02_run33_pytest_frontend_gpu.log
for v in conv_block conv_block_noidiom; do
  echo "== $v (wrapper's own timing, 101 runs)"
  d=$(mktemp -d); python - "$d" <<'PY'
import sys, numpy as np
sys.path.insert(0, ".")
from oracle import oracle
from tests import refdata
inp, filt, bias = refdata.conv_inputs(32, 224, 3, 32)
want = oracle.conv_relu_maxpool(inp, filt, bias)
np.savetxt(sys.argv[1] + "/mkl_result.txt", np.ascontiguousarray(want.transpose(0, 1, 4, 2, 3)).reshape(-1), fmt="%.10g")
PY
  (cd $d && timeout 600 /root/repo/tests/compat/_bin/gen_$v 2>&1 | grep -i "time\|median\|correct" | tail -4) | tee -a gpurun_out/r02_run33_conv_block_idiom_vs_general.log
done
