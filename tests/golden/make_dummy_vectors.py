#!/usr/bin/env python3
"""Turns the reference's pre-modulated 802.11a frames (kernel/sample/mac/Dot11ADummy*.txt, C array initialisers that the
sample MAC transmits) into compact binary fixtures.  Data, not source: only the numbers are carried over.
Run in the build container (needs /root/reference):  python tests/golden/make_dummy_vectors.py

  Dot11ADummy.txt            SHORT[]      20 Msps  37 608 complex int16  -> dot11a_dummy_20m.i16
  Dot11ADummy_16.txt         COMPLEX16[]  40 Msps  75 215 complex int16  -> dot11a_dummy_16_40m.i16
  Dot11ADummy_8.txt          COMPLEX8[]   20 Msps     888 complex int8   -> dot11a_dummy_8_20m.i8
  Dot11ADummy_8_ack_40M.txt  COMPLEX8[]   40 Msps   1 827 complex int8   -> dot11a_dummy_8_ack_40m.i8
"""
import os, re, sys, numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/kernel/sample/mac"
FILES = [("Dot11ADummy.txt", "dot11a_dummy_20m.i16", np.int16), ("Dot11ADummy_16.txt", "dot11a_dummy_16_40m.i16", np.int16),
         ("Dot11ADummy_8.txt", "dot11a_dummy_8_20m.i8", np.int8), ("Dot11ADummy_8_ack_40M.txt", "dot11a_dummy_8_ack_40m.i8", np.int8)]

def parse(path):
    t = open(path).read(); body = t[t.index("{") + 1:t.rindex("}")]
    return np.array([int(x, 0) for x in re.findall(r"-?(?:0x[0-9a-fA-F]+|\d+)", body)], dtype=np.int64)

if __name__ == "__main__":
    for src, dst, dt in FILES:
        v = parse(os.path.join(SRC, src))
        assert len(v) % 2 == 0 and v.min() >= np.iinfo(dt).min and v.max() <= np.iinfo(dt).max, src
        v.astype(dt).tofile(os.path.join(HERE, dst))
        print("%-28s %6d complex samples, |max| %5d -> %s" % (src, len(v) // 2, np.abs(v).max(), dst))
