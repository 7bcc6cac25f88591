#!/bin/bash
# Round-2 (second session) GPU call 1: GPU tests, Viterbi variant A/B (v3 shipping, v4 pair, v5 pair + global ring, v6 quad + global ring),
# host gather microbenchmark, e2e mode sweep, ncu captures of the two global-ring variants.  Output: gpurun_out/r2b_*.
cd "$(dirname "$0")/.."; mkdir -p gpurun_out; T=r2b
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/${T}_pytest_gpu.txt
for v in v3 v4 v5 v6; do
  SB200_VITERBI=$v python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu 2>gpurun_out/${T}_bench_$v.err | tail -1 > gpurun_out/${T}_bench_$v.json
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${T}_bench_$v.json")); print("$v", round(d["value"]), "Msamples/s", d["kernel_ms"])
except Exception as e: print("$v failed", e)
PY
done
for v in v5 v6; do SB200_VITERBI=$v timeout 300 python -m pytest tests/test_gpu_rx11a.py tests/test_gpu_rx11n.py -x -q 2>&1 | tail -2 | tee gpurun_out/${T}_pytest_$v.txt; done
( cd tools/microbench && g++ -O2 -std=c++17 -pthread -o host_gather_bench host_gather_bench.cpp ../../sora_b200/csrc/host_gather.cpp
  for t in 1 4 8 14 16 24 32; do ./host_gather_bench $t; SB200_GATHER=pf ./host_gather_bench $t; done ) 2>&1 | tee gpurun_out/${T}_host_gather.txt
SB200_TRACE=1 python bench.py --steps 5 --warmup 3 --no-cpu --e2e-sweep 2>gpurun_out/${T}_bench_sweep.err | tail -1 > gpurun_out/${T}_bench_sweep.json
grep "e2e sweep" gpurun_out/${T}_bench_sweep.err | tee gpurun_out/${T}_e2e_sweep.txt
grep "host_decimate:" gpurun_out/${T}_bench_sweep.err | tail -40 > gpurun_out/${T}_e2e_trace.txt
python - <<PY
import json
d = json.load(open("gpurun_out/r2b_bench_sweep.json")); print("value", round(d["value"]), "e2e", d["e2e"]["mode"], round(d["e2e"]["value"]), {k: round(v["value"]) for k, v in d["e2e"]["modes"].items()}, d.get("e2e_brick"))
PY
for v in v5 v6; do
  SB200_VITERBI=$v timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_viterbi_re -c 1 -f -o gpurun_out/${T}_viterbi_$v python bench.py --frames 16384 --steps 1 --warmup 0 --no-e2e --no-cpu > /dev/null 2>&1
done
ls -la gpurun_out | grep ${T}
