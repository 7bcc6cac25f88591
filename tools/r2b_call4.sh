#!/bin/bash
# GPU call 4: the lane kernel with 8-column history blocks at 16 / 12 / 10 / 8 resident warps per SM, against 6-column blocks; variant parity test; ncu of two points.
cd "$(dirname "$0")/.."; mkdir -p gpurun_out; T=r2e
python -m pytest tests/test_gpu_zz_viterbi_variants.py -x -q 2>&1 | tail -4 | tee gpurun_out/${T}_pytest_variants.txt
for hb in 8 6; do for pad in 0 18000 22000 28000; do
  python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --lane-min 0 --vl-hist-block $hb --vl-pad-smem $pad --vl-l2-hints 1 2>/dev/null | tail -1 > gpurun_out/${T}_bench_hb${hb}_pad$pad.json
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${T}_bench_hb${hb}_pad$pad.json")); print("lane kernel, $hb-column blocks, pad $pad, ring evict_last:", round(d["value"]), "Msamples/s", round(d["kernel_ms"]["viterbi_descramble_crc"], 3), "ms")
except Exception as e: print("hb$hb pad $pad failed", e)
PY
done; done
for pad in 0 28000; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_viterbi_lane -c 1 -f -o gpurun_out/${T}_viterbi_hb8_pad$pad python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu --lane-min 0 --vl-hist-block 8 --vl-pad-smem $pad --vl-l2-hints 1 > /dev/null 2>&1
done
ls -la gpurun_out | grep ${T}
