#!/bin/bash
# e2e (host buffers through the C ABI) against the pipeline chunk size
for c in 2048 4096 8192 16384; do
  echo -n "chunk $c: "; python bench.py --steps 5 --warmup 3 --no-cpu --chunk $c 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['e2e']['ms_per_step'], d['e2e']['value'])"
done
