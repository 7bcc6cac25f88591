"""GPU parity tests, 802.11b: CUDA path through the C ABI vs the CPU oracle on the same IQ."""
import os, numpy as np, pytest
import oracle_py
from sora_b200 import api, synth
from test_cpu_oracle_11b import _fixture, FRAME

pytestmark = pytest.mark.gpu

@pytest.fixture(scope="module")
def eng():
    return api.Engine(0)

def _cmp(eng, iq2, off, ln):
    res, out = eng.rx11b_batch(iq2, off, ln)
    ores, oout = oracle_py.rx11b_batch(iq2, off, ln, out_stride=out.shape[1])
    assert (res["status"] == ores["status"]).all(), (res["status"], ores["status"])
    ev = ores["status"] != oracle_py.E_NO_FRAME
    for k in ("rate_kbps", "length", "crc32", "sample_index", "detect_vec"):
        assert (res[k][ev] == ores[k][ev]).all(), (k, res[k], ores[k])
    for i in np.nonzero(ev)[0]:
        if ores["status"][i] in (1, oracle_py.E_CRC32_FAIL):
            L = int(ores["length"][i]) - 1
            assert (out[i, :L] == oout[i, :L]).all(), f"slot {i}"
    return res, out

def test_reference_fixtures(eng):
    for name, rate in (("1long44.mf.bin", 1000), ("2long44.mf.bin", 2000)):
        iq = _fixture(name)
        res, out = _cmp(eng, iq, np.array([0], np.uint64), np.array([len(iq)], np.uint32))
        assert res[0]["status"] == 1 and res[0]["rate_kbps"] == rate and (out[0, :113] == FRAME[:113]).all()

@pytest.mark.parametrize("rate", [1000, 2000, 5500, 11000])
def test_all_rates(eng, rate):
    for snr, gain in ((None, 1.0), (40, 0.3), (28, 0.25), (18, 0.2)):
        iq, ps = synth.make_frames_11b(5, psdu_len=96, rate_kbps=rate, snr_db=snr, gain=gain, seed0=rate + 7)
        F, slot, _ = iq.shape
        res, out = _cmp(eng, iq.reshape(-1, 2), np.arange(F, dtype=np.uint64) * slot, np.full(F, slot, np.uint32))
        if snr in (None, 40) and not (rate == 11000 and snr is None):
            assert (res["status"] == 1).all() and (out[:, :95] == ps[:, :95]).all()

def test_edges(eng):
    rng = np.random.default_rng(4)
    noise = rng.normal(0, 2500, (3, 28 * 300, 2)).astype(np.int16)
    _cmp(eng, noise.reshape(-1, 2), np.arange(3, dtype=np.uint64) * 8400, np.full(3, 8400, np.uint32))
    iq, _ = synth.make_frames_11b(4, psdu_len=200, rate_kbps=11000, snr_db=40, gain=0.3)
    F, slot, _ = iq.shape
    ln = np.array([slot, slot // 2, 3000, 27], np.uint32)          # truncated / tiny slots
    _cmp(eng, iq.reshape(-1, 2), np.arange(F, dtype=np.uint64) * slot, ln)
    z = np.zeros((2, 5600, 2), np.int16)
    _cmp(eng, z.reshape(-1, 2), np.arange(2, dtype=np.uint64) * 5600, np.full(2, 5600, np.uint32))
