#!/bin/bash
# verification of the cleaned-up build: GPU suite, smoke, both bench arms (default flags)
cd "$(dirname "$0")/.."; mkdir -p gpurun_out; T=r2b_final2
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/${T}_pytest_gpu.txt
python __graft_entry__.py smoke 2>&1 | tail -1 | tee gpurun_out/${T}_smoke.txt
python bench.py --impl reference 2>/dev/null | tail -1 > gpurun_out/${T}_bench_reference.json
python bench.py 2>gpurun_out/${T}_bench.err | tail -1 > gpurun_out/${T}_bench.json
python -c "
import json; d = json.load(open('gpurun_out/${T}_bench.json')); print('value', round(d['value']), d['kernel_ms'], 'frac', round(d['roofline']['frac'], 4), 'traffic', d['roofline']['traffic'], 'e2e', d['e2e']['mode'], round(d['e2e']['value']), 'brick', d.get('e2e_brick', {}).get('value'), 'cpu', round(d['cpu_baseline']['value']))
r = json.load(open('gpurun_out/${T}_bench_reference.json')); print('reference arm', round(r['value']))"
