#!/bin/bash
# GPU call 3: the lane kernel (SB200_VITERBI=v8): parity tests, bench A/B, full-size ncu capture.
cd "$(dirname "$0")/.."; mkdir -p gpurun_out; T=r2d; V=${1:-v8}
SB200_VITERBI=$V timeout 600 python -m pytest tests/test_gpu_rx11a.py tests/test_gpu_rx11n.py tests/test_gpu_11n_qam.py tests/test_gpu_zz_viterbi_variants.py -x -q 2>&1 | tail -15 | tee gpurun_out/${T}_pytest_$V.txt
SB200_VITERBI=$V python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu 2>gpurun_out/${T}_bench_$V.err | tail -1 > gpurun_out/${T}_bench_$V.json
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${T}_bench_$V.json")); print("$V", round(d["value"]), "Msamples/s", d["kernel_ms"])
except Exception as e: print("$V failed", e)
PY
tail -3 gpurun_out/${T}_bench_$V.err
SB200_VITERBI=$V python bench_extra.py --config viterbi 2>/dev/null | tee gpurun_out/${T}_extra_viterbi_$V.jsonl | cut -c1-260
SB200_VITERBI=$V timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_viterbi_lane -c 1 -f -o gpurun_out/${T}_viterbi_$V python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu > /dev/null 2>&1
for pad in 18000 28000 37000; do
  python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --lane-min 0 --vl-pad-smem $pad 2>/dev/null | tail -1 > gpurun_out/${T}_bench_pad$pad.json
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${T}_bench_pad$pad.json")); print("lane kernel, pad $pad", round(d["value"]), "Msamples/s", round(d["kernel_ms"]["viterbi_descramble_crc"], 3), "ms")
except Exception as e: print("pad $pad failed", e)
PY
done
for pad in 0 28000; do
  python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --lane-min 0 --vl-hist-block 8 --vl-pad-smem $pad 2>/dev/null | tail -1 > gpurun_out/${T}_bench_hb8_pad$pad.json
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${T}_bench_hb8_pad$pad.json")); print("lane kernel, 8-column blocks, pad $pad", round(d["value"]), "Msamples/s", round(d["kernel_ms"]["viterbi_descramble_crc"], 3), "ms")
except Exception as e: print("hb8 pad $pad failed", e)
PY
done
for hints in 1 2 3; do
  python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --lane-min 0 --vl-l2-hints $hints 2>/dev/null | tail -1 > gpurun_out/${T}_bench_hints$hints.json
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${T}_bench_hints$hints.json")); print("lane kernel, L2 hints $hints", round(d["value"]), "Msamples/s", round(d["kernel_ms"]["viterbi_descramble_crc"], 3), "ms")
except Exception as e: print("hints $hints failed", e)
PY
done
python tools/vit_crossover.py 2>&1 | tee gpurun_out/${T}_vit_crossover.jsonl
SB200_TRACE=1 python bench.py --steps 5 --warmup 3 --no-cpu --e2e-wc 2>gpurun_out/${T}_bench_wc.err | tail -1 > gpurun_out/${T}_bench_wc.json
python - <<PY
import json
d = json.load(open("gpurun_out/${T}_bench_wc.json")); print("value", round(d["value"]), "e2e", d["e2e"]["mode"], round(d["e2e"]["value"]), {k: (round(v["value"]), v.get("chunks_gathered_on_host")) for k, v in d["e2e"]["modes"].items()})
PY
ls -la gpurun_out | grep ${T}
