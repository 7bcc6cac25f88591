#!/bin/bash
# End-of-round capture on the GPU box: full GPU test suite, sanitizer pass over the newest kernels, both bench arms, the launch list of
# the bench command and a --set full capture of the front end.  Everything lands in gpurun_out/ (scratch); summaries are copied to
# profiles/ by hand afterwards.
cd "$(dirname "$0")/.."; mkdir -p gpurun_out; T=${1:-r1_final2}
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/${T}_pytest_gpu.txt
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_tx11b.py "tests/test_gpu_rx11b.py::test_streams_match_oracle_driver" "tests/test_gpu_rx11n.py::test_streams_match_oracle_driver" -x -q -k "not loopback" 2>&1 | tail -4 | tee gpurun_out/${T}_memcheck.txt
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/${T}_bench_reference.json
python bench.py --steps 10 --warmup 3 2>/dev/null | tail -1 > gpurun_out/${T}_bench.json
cut -c1-260 gpurun_out/${T}_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_front11a -c 1 -f -o gpurun_out/${T}_front python bench.py --frames 16384 --steps 1 --warmup 0 --no-e2e --no-cpu > /dev/null 2>&1
ls -la gpurun_out | grep ${T}
