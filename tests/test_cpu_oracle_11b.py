"""CPU suite, 802.11b: the oracle against the reference's own TX fixtures (kernel/HWTest/exe/tx samples) and a float
modulator at all four rates."""
import os, numpy as np, pytest
import oracle_py
from sora_b200 import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")

def _fixture(name):
    """`*.mf.bin` = COMPLEX8 TX waveform at 44 Msps; converted like ConvertModFile2DumpFile_8b (modulate11a.cpp:131-190):
    << 8, padded to whole 28-sample blocks plus silence."""
    b = np.fromfile(os.path.join(GOLD, name), dtype=np.int8).reshape(-1, 2)
    iq = b.astype(np.int16) << 8
    pad = (-len(iq)) % 28 + 56
    return np.ascontiguousarray(np.concatenate([np.zeros((280, 2), np.int16), iq, np.zeros((pad, 2), np.int16)]))

FRAME = np.array([int(x, 16) for x in open(os.path.join(GOLD, "frame.txt")).read().split()], np.uint8)

@pytest.mark.parametrize("name,rate", [("1long44.mf.bin", 1000), ("2long44.mf.bin", 2000)])
def test_reference_tx_fixtures_decode_to_frame_txt(name, rate):
    """SURVEY.md §8c pin (2): byte-exact decode of the reference's 11b waveforms (114-byte MPDU incl. FCS 7B 74 B2 5A)."""
    res, out = oracle_py.rx11b_run(_fixture(name))
    assert len(res) == 1 and res[0]["status"] == 1 and res[0]["rate_kbps"] == rate and res[0]["length"] == 114
    assert (out[0, :113] == FRAME[:113]).all()                  # the 4th FCS byte is never delivered (PHY_11b.hpp:721-739)
    assert res[0]["crc32"] == 0xB2747B

@pytest.mark.parametrize("rate", [1000, 2000, 5500, 11000])
def test_roundtrip_all_rates(rate):
    iq, ps = synth.make_frames_11b(3, psdu_len=130, rate_kbps=rate, snr_db=40, gain=0.3, seed0=rate)
    F, slot, _ = iq.shape
    res, out = oracle_py.rx11b_batch(iq.reshape(-1, 2), np.arange(F) * slot, np.full(F, slot))
    assert (res["status"] == 1).all() and (res["rate_kbps"] == rate).all() and (res["length"] == 130).all()
    assert (out[:, :129] == ps[:, :129]).all()

def test_no_signal_and_garbage():
    res, _ = oracle_py.rx11b_run(np.zeros((5600, 2), np.int16)); assert len(res) == 0
    rng = np.random.default_rng(2)
    res, _ = oracle_py.rx11b_run(rng.normal(0, 3000, (28 * 400, 2)).astype(np.int16), max_frames=64)
    assert all(r["status"] != 1 for r in res)
