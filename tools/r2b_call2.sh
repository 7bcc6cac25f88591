#!/bin/bash
# GPU call 2: tests (default and SB200_VITERBI=v7: one lane per code block), Viterbi A/B v3 / v7, e2e with the gather of chunk k+1 running
# while chunk k is queued, full-size ncu captures of v5 and v7.
cd "$(dirname "$0")/.."; mkdir -p gpurun_out; T=r2c
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/${T}_pytest_gpu.txt
SB200_VITERBI=v7 timeout 600 python -m pytest tests/test_gpu_rx11a.py tests/test_gpu_rx11n.py tests/test_gpu_11n_qam.py -x -q 2>&1 | tail -5 | tee gpurun_out/${T}_pytest_v7.txt
for v in v3 v7; do
  SB200_VITERBI=$v python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu 2>gpurun_out/${T}_bench_$v.err | tail -1 > gpurun_out/${T}_bench_$v.json
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${T}_bench_$v.json")); print("$v", round(d["value"]), "Msamples/s", d["kernel_ms"])
except Exception as e: print("$v failed", e)
PY
  tail -3 gpurun_out/${T}_bench_$v.err
done
SB200_TRACE=1 python bench.py --steps 5 --warmup 3 --no-cpu 2>gpurun_out/${T}_bench_e2e.err | tail -1 > gpurun_out/${T}_bench_e2e.json
grep "host_decimate:" gpurun_out/${T}_bench_e2e.err | cut -c30- | tail -20 > gpurun_out/${T}_e2e_trace.txt; tail -9 gpurun_out/${T}_e2e_trace.txt
for ch in 2048 8192; do SB200_TRACE=1 python bench.py --steps 5 --warmup 3 --no-cpu --chunk $ch 2>/dev/null | tail -1 > gpurun_out/${T}_bench_e2e_chunk$ch.json; done
python - <<PY
import json
for f in ("e2e", "e2e_chunk2048", "e2e_chunk8192"):
    try:
        d = json.load(open("gpurun_out/r2c_bench_%s.json" % f)); print(f, "value", round(d["value"]), "e2e", d["e2e"]["mode"], round(d["e2e"]["value"]), {k: (round(v["value"]), v.get("chunks_gathered_on_host")) for k, v in d["e2e"]["modes"].items()})
    except Exception as e: print(f, "failed", e)
PY
for v in v5 v7; do
  SB200_VITERBI=$v timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_viterbi_re -c 1 -f -o gpurun_out/${T}_viterbi_$v python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu > /dev/null 2>&1
done
ls -la gpurun_out | grep ${T}
