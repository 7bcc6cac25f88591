#!/bin/bash
# occupancy sweep of the Viterbi kernel: pad the CTA's shared memory so that fewer CTAs fit per SM
for pad in 0 2000 4500 8000 12000 17000 26000; do
  echo -n "pad $pad: "; python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --vq-pad-smem $pad 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['kernel_ms']['viterbi_descramble_crc'])"
done
