#!/bin/bash
# source-level capture of the OFDM front end (now 39 % of the step) and of the sink
cd "$(dirname "$0")/.."; mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_front11a -c 1 -f -o gpurun_out/r2k_front python bench.py --frames 16384 --steps 1 --warmup 0 --no-e2e --no-cpu > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_sink11a -c 1 -f -o gpurun_out/r2k_sink python bench.py --frames 65536 --steps 1 --warmup 0 --no-e2e --no-cpu > /dev/null 2>&1
ls -la gpurun_out | grep r2k
