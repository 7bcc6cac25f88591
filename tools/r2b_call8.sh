#!/bin/bash
# GPU call 8: the deferred walk of the lane kernel: parity, A/B, ncu.
cd "$(dirname "$0")/.."; mkdir -p gpurun_out; T=r2i
python -m pytest tests/test_gpu_zz_viterbi_variants.py -x -q 2>&1 | tail -3 | tee gpurun_out/${T}_pytest.txt
for df in 0 1; do
  python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --vl-defer $df 2>/dev/null | tail -1 > gpurun_out/${T}_bench_defer$df.json
  python -c "
import json; d = json.load(open('gpurun_out/${T}_bench_defer$df.json')); print('defer $df:', round(d['value']), 'Msamples/s', d['kernel_ms'])"
done
for df in 0 1; do SB200_VITERBI=v8 python - <<PY
import os, sys; sys.path.insert(0, os.getcwd())
# the three rates of the standalone decoder (bench_extra's configuration) with the walk at the trigger / deferred
import subprocess
PY
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_viterbi_lane -c 1 -f -o gpurun_out/${T}_viterbi_defer python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu --vl-defer 1 > /dev/null 2>&1
ls -la gpurun_out | grep ${T}
