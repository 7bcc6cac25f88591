#!/bin/bash
# Final capture of round 2 (second session) on the GPU box (one GPU): full GPU test suite, smoke, both bench arms, every bench_extra
# configuration, the launch list of the bench command and --set full captures of the Viterbi kernel the bench launch uses and of the front end.
# Everything lands in gpurun_out/ (scratch); summaries are copied to profiles/ afterwards (tools/ncu_summary.py).
cd "$(dirname "$0")/.."; mkdir -p gpurun_out; T=${1:-r2b_final}
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/${T}_pytest_gpu.txt
python __graft_entry__.py smoke 2>&1 | tail -1 | tee gpurun_out/${T}_smoke.txt
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/${T}_bench_reference.json
python bench.py --steps 10 --warmup 3 2>gpurun_out/${T}_bench.err | tail -1 > gpurun_out/${T}_bench.json
cut -c1-400 gpurun_out/${T}_bench.json
for c in viterbi 11b 11n; do python bench_extra.py --config $c 2>/dev/null > gpurun_out/${T}_extra_${c}.jsonl; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_viterbi_ -c 1 -f -o gpurun_out/${T}_viterbi python bench.py --frames 65536 --steps 1 --warmup 0 --no-e2e --no-cpu > /dev/null 2>&1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_zz_viterbi_variants.py "tests/test_gpu_rx11a.py::test_viterbi_wrap_and_ragged_stress" "tests/test_gpu_rx11a.py::test_chunked_pipeline_host_and_device" -x -q 2>&1 | tail -4 | tee gpurun_out/${T}_memcheck.txt
ls -la gpurun_out | grep ${T}
