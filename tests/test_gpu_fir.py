"""GPU tests of the anti-alias FIR decimator (sb200_fir_decimate2) and of the 20 Msps entry of the 802.11a chain.  The decimator is an
extension without a reference counterpart (the reference's graph only drops every other sample): its oracle is the arithmetic stated in
include/sora_b200.h, restated here in numpy with 64-bit integers."""
import numpy as np, pytest
import oracle_py
from sora_b200 import api, synth

pytestmark = pytest.mark.gpu
HALF_BAND_31 = np.array([-121, 0, 209, 0, -381, 0, 644, 0, -1056, 0, 1759, 0, -3278, 0, 10391, 16434, 10391, 0, -3278, 0, 1759, 0, -1056, 0, 644, 0, -381, 0, 209, 0, -121], np.int64)

@pytest.fixture(scope="module")
def eng():
    return api.Engine(0)

def fir_model(iq, taps):
    taps = np.asarray(taps, np.int64); c = len(taps) // 2; n = len(iq); m = (n + 1) // 2
    x = np.zeros((n + 2 * c + 2, 2), np.int64); x[c:c + n] = iq
    out = np.zeros((m, 2), np.int64)
    for k, t in enumerate(taps):
        if t: out += t * x[k: k + 2 * m: 2][:m]
    return np.clip((out + (1 << 14)) >> 15, -32768, 32767).astype(np.int16)

@pytest.mark.parametrize("n", [1, 2, 7, 31, 4095, 4096, 4097, 8191, 12345, 70001])
def test_fir_decimate2_matches_the_stated_arithmetic(eng, n):
    rng = np.random.default_rng(n)
    iq = rng.integers(-32768, 32768, (n, 2)).astype(np.int16)                  # full-scale noise: saturation and rounding both occur
    assert (eng.fir_decimate2(iq) == fir_model(iq, HALF_BAND_31)).all()
    for nt in (1, 3, 17, 63):
        taps = rng.integers(-3000, 3000, nt).astype(np.int16); taps[nt // 2] = 20000
        assert (eng.fir_decimate2(iq, taps) == fir_model(iq, taps)).all(), (n, nt)

def test_fir_then_20msps_chain_decodes_and_rejects_an_adjacent_channel(eng):
    """A frame under a strong interferer 11.5 .. 18.5 MHz off the carrier: dropping every other sample (the reference's TDownSample2) folds it onto the
    channel, the FIR decimator removes it first.  Also: the even samples handed to the 20 Msps entry give what the 40 Msps chain gives."""
    iq, ps = synth.make_frames(4, psdu_len=300, rate_kbps=36000, snr_db=30, seed0=0xF1, lead=120, trail=136)
    F, slot, _ = iq.shape
    flat = iq.reshape(-1, 2)
    off = np.arange(F, dtype=np.uint64) * slot; ln = np.full(F, slot, np.uint32)
    ref, refo = eng.rx11a_batch(flat, off, ln)
    pick = flat[::2].copy()
    r20, o20 = eng.rx11a_batch(pick, off // 2, ln // 2, sample_rate_mhz=20)
    for k in ("status", "rate_kbps", "length", "crc32", "nsym", "detect_index", "cfo_est"): assert (r20[k] == ref[k]).all(), k
    assert (o20 == refo).all() and (ref["status"] == 1).all()
    rng = np.random.default_rng(5); N = len(flat)                              # band-limited interferer 11.5 .. 18.5 MHz off the carrier, stronger than the frame
    S = (rng.normal(size=N) + 1j * rng.normal(size=N)); fr = np.fft.fftfreq(N, 1 / 40e6); S[(fr < 11.5e6) | (fr > 18.5e6)] = 0
    w = np.fft.ifft(S); w *= 6000 / np.sqrt((np.abs(w) ** 2).mean() / 2)
    jam = np.clip(flat.astype(np.int32) + np.stack([w.real, w.imag], 1).round().astype(np.int32), -32768, 32767).astype(np.int16)
    rj, _ = eng.rx11a_batch(jam, off, ln)                                      # plain decimation folds the interferer onto the channel
    y = eng.fir_decimate2(jam)
    rf, of = eng.rx11a_batch(y, off // 2, ln // 2, sample_rate_mhz=20)
    assert (rf["status"] == 1).all() and (of[:, :300] == ps).all()
    assert (rj["status"] != 1).any()
