#!/bin/bash
# GPU call 9: deferred walk + deeper soft-value prefetch: parity (variants + chains), A/B, crossover, ncu of the default.
cd "$(dirname "$0")/.."; mkdir -p gpurun_out; T=r2j
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/${T}_pytest_gpu.txt
for df in 0 1; do
  python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --vl-defer $df 2>/dev/null | tail -1 > gpurun_out/${T}_bench_defer$df.json
  python -c "
import json; d = json.load(open('gpurun_out/${T}_bench_defer$df.json')); print('defer $df:', round(d['value']), 'Msamples/s', d['kernel_ms'])"
done
python tools/vit_crossover.py 2>&1 | tee gpurun_out/${T}_vit_crossover.jsonl
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_viterbi_lane -c 1 -f -o gpurun_out/${T}_viterbi_default python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu > /dev/null 2>&1
ls -la gpurun_out | grep ${T}
