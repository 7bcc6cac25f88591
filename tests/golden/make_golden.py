#!/usr/bin/env python3
"""Regenerates fsample-6.psdu.bin from fsample-6.dmp with the CPU oracle (run from the repo root)."""
import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_py
from sora_b200.dumpfile import load_dump
iq = load_dump(os.path.join(ROOT, "tests/golden/fsample-6.dmp"))
iq = (iq.astype(np.int32) << 2).astype(np.int16)
res, out = oracle_py.rx11a_run(iq)
assert res[0]["status"] == 1 and res[0]["length"] == 1392
out[0, :1392].tofile(os.path.join(ROOT, "tests/golden/fsample-6.psdu.bin"))
print("crc32 %08x" % res[0]["crc32"])
