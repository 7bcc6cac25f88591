"""Sora RX dump files (`*.dmp`): 128-byte RX_BLOCKs = 16-byte descriptor + 28 COMPLEX16 samples.

Mirrors LoadSoraDumpFile (reference kernel/brick/inc/brickutil.h:21-59; RX_BLOCK in
kernel/core/inc/_rx_manager.h:79-113).  `sign_extend_14` applies the legacy 14-bit sample fix
(RX_COMPLEX16_INVALID_BITS 2, kernel/core/inc/const.h:73) that old captures such as
kernel/test-data/fsample-6.dmp need.
"""
import numpy as np

SAMPLES_PER_BLOCK = 28
BLOCK_BYTES = 128
DESC_BYTES = 16

def load_dump(path, sign_extend_14=False):
    raw = np.fromfile(path, dtype=np.uint8)
    nblk = len(raw) // BLOCK_BYTES
    blk = raw[: nblk * BLOCK_BYTES].reshape(nblk, BLOCK_BYTES)
    iq = blk[:, DESC_BYTES:].copy().view("<i2").reshape(nblk * SAMPLES_PER_BLOCK, 2)
    if sign_extend_14:
        iq = ((iq.astype(np.int32) << 2).astype(np.int16) >> 2).astype(np.int16)
    return np.ascontiguousarray(iq)

def write_dump(path, iq):
    """Inverse of load_dump, padding to whole blocks (descriptor `01 00 70 00 ...`, modulate11a.cpp:137)."""
    iq = np.asarray(iq, dtype=np.int16).reshape(-1, 2)
    pad = (-len(iq)) % SAMPLES_PER_BLOCK
    if pad:
        iq = np.concatenate([iq, np.zeros((pad, 2), np.int16)])
    nblk = len(iq) // SAMPLES_PER_BLOCK
    out = np.zeros((nblk, BLOCK_BYTES), np.uint8)
    out[:, 0] = 1; out[:, 2] = 0x70
    out[:, DESC_BYTES:] = iq.reshape(nblk, -1).view(np.uint8)
    out.tofile(path)
