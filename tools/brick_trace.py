import os, sys, subprocess, tempfile, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from sora_b200 import synth
from sora_b200.dumpfile import write_dump
iq, ps = synth.make_frames(256, psdu_len=1500, rate_kbps=54000, snr_db=30.0, lead=32, trail=32)
cap = iq.reshape(-1, 2); cap = cap[: len(cap) // 28 * 28]
p = '/tmp/bench.dmp'; write_dump(p, cap)
env = dict(os.environ, SB200_TRACE='1')
r = subprocess.run(['sora_b200/brick/demo_graph', p, '--threads', '16', '--repeat', '2'], capture_output=True, text=True, env=env)
print(r.stdout[-600:]); print(r.stderr[-3000:])
