#!/bin/bash
# GPU call 7: device-resident input cut into chunks so that the front end of chunk k+1 overlaps the (latency-bound) lane Viterbi of chunk k.
cd "$(dirname "$0")/.."; mkdir -p gpurun_out; T=r2h
for cd in 0 32768 21846 16384; do for lm in 32768 16384; do
  python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --chunk-device $cd --lane-min $lm 2>/dev/null | tail -1 > gpurun_out/${T}_bench_cd${cd}_lm$lm.json
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${T}_bench_cd${cd}_lm$lm.json")); print("chunk-device $cd, lane-min $lm:", round(d["value"]), "Msamples/s", round(d["ms_per_step"], 3), "ms/step")
except Exception as e: print("cd $cd failed", e)
PY
done; done
ls -la gpurun_out | grep ${T}
