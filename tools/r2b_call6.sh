#!/bin/bash
# GPU call 6: start stagger of the lane kernel's warps (A/B), with and without the window prefetch; default bench; ncu of the default.
cd "$(dirname "$0")/.."; mkdir -p gpurun_out; T=r2g
python -m pytest tests/test_gpu_zz_viterbi_variants.py tests/test_gpu_rx11a.py -x -q 2>&1 | tail -3 | tee gpurun_out/${T}_pytest.txt
for cfg in "8 13" "8 5" "8 1" "6 13" "6 5"; do set -- $cfg
  python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --lane-min 0 --vl-hist-block $1 --vl-l2-hints $2 2>/dev/null | tail -1 > gpurun_out/${T}_bench_hb$1_f$2.json
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${T}_bench_hb$1_f$2.json")); f = $2
    print("lane kernel, $1-column blocks, stagger", "off" if f & 8 else "on", ", prefetch", "off" if f & 4 else "on", ":", round(d["value"]), "Msamples/s", round(d["kernel_ms"]["viterbi_descramble_crc"], 3), "ms")
except Exception as e: print("hb$1 flags $2 failed", e)
PY
done
python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu 2>/dev/null | tail -1 > gpurun_out/${T}_bench_default.json
python -c "
import json; d = json.load(open('gpurun_out/${T}_bench_default.json')); print('default:', round(d['value']), d['kernel_ms'], d['roofline']['kernel'])"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_viterbi_lane -c 1 -f -o gpurun_out/${T}_viterbi_default python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu > /dev/null 2>&1
python tools/vit_crossover.py 2>&1 | tee gpurun_out/${T}_vit_crossover.jsonl
ls -la gpurun_out | grep ${T}
