#!/usr/bin/env python3
"""Known answers of the reference's legacy 802.11b transmit filter, made by the reference's own code: oracle/_ref/libfir37_ref.so is the body of
FIR37SSE_INTRINSIC + its coefficient table compiled from /root/reference/kernel/bb/dot11b/bbb_fir.c by oracle/build_ref.sh.
Writes tests/golden/fir37/fir37_<name>.in.i8 / .out.i8 (COMPLEX8, re/im interleaved).  Run in the build container (needs oracle/_ref)."""
import os, sys, numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_py

def inputs():
    rng = np.random.default_rng(20090707)
    yield "random", rng.integers(-128, 128, (2048, 2)).astype(np.int8)
    x = np.full((512, 2), 127, np.int8); x[256:] = -128; x[100:110, 1] = -128                               # lane tree saturates at both rails
    yield "saturating", x
    c = np.zeros((4096, 2), np.int8); c[::4, 0] = np.where(rng.integers(0, 2, 1024) > 0, 127, -128)      # DBPSK chips, 4x zero-stuffed (bbb_dbpsk.c)
    yield "dbpsk_chips", c
    q = np.zeros((4096, 2), np.int8); k = rng.integers(0, 4, 1024)                                          # QPSK / CCK chips on the axes
    q[::4, 0] = np.array([127, 0, -128, 0], np.int8)[k]; q[::4, 1] = np.array([0, 127, 0, -128], np.int8)[k]
    yield "qpsk_chips", q

if __name__ == "__main__":
    assert oracle_py.ref_fir37_available(), "build oracle/_ref first (oracle/build_ref.sh)"
    for name, x in inputs():
        y = oracle_py.ref_fir37(x)
        x.tofile(os.path.join(HERE, "fir37", f"fir37_{name}.in.i8")); y.tofile(os.path.join(HERE, "fir37", f"fir37_{name}.out.i8"))
        print(name, x.shape, int(np.abs(y.astype(int)).max()))
