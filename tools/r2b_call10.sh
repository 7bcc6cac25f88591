#!/bin/bash
# 2-GPU run of the bench (torchrun): weak scaling value, e2e, the NCCL scatter-decode-gather leg (16 384-slot pieces: the lane kernel), sharded check.
cd "$(dirname "$0")/.."; mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 2>gpurun_out/r2b_final_bench_2gpu.err | tail -1 > gpurun_out/r2b_final_bench_2gpu.json
python -c "
import json; d = json.load(open('gpurun_out/r2b_final_bench_2gpu.json')); print('2 GPUs: value', round(d['value']), 'e2e', d['e2e']['mode'], round(d['e2e']['value']), 'mgpu', d.get('mgpu', {}).get('value'), d.get('mgpu', {}).get('sharded_check'))"
tail -3 gpurun_out/r2b_final_bench_2gpu.err
