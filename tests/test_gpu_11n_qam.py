"""GPU parity tests for 802.11n MCS 11..14 (16-QAM / 64-QAM, two streams; pytest -m gpu): engine option ht_mcs_limit = 15 opens the HT-SIG gate
the reference keeps shut (PHY_11n.hpp:496-501) and the frames go through the 16-/64-QAM branches of its graphs — CUDA against the CPU oracle on
the same IQ, receive and transmit, and the default (refuse, like the reference)."""
import numpy as np, pytest
import oracle_py
from sora_b200 import api, synth

pytestmark = pytest.mark.gpu

@pytest.fixture(scope="module")
def eng():
    e = api.Engine(0); e.set_option("ht_mcs_limit", 15)
    oracle_py.set_ht_mcs_limit(15)
    yield e
    oracle_py.set_ht_mcs_limit(11)

def _compare(eng, iq0, iq1, off, ln):
    res, out = eng.rx11n_batch(iq0, iq1, off, ln)
    ores, oout = oracle_py.rx11n_batch(iq0, iq1, off, ln, out_stride=out.shape[1])
    assert (res["status"] == ores["status"]).all(), (res["status"], ores["status"])
    m = ores["status"] != oracle_py.E_NO_FRAME
    for k in ("mcs", "length", "crc32", "nsym", "detect_index", "cfo_est", "lsig_length"):
        assert (res[k][m] == ores[k][m]).all(), (k, res[k], ores[k])
    for i in range(len(res)):
        if ores["status"][i] in (1, oracle_py.E_CRC32_FAIL):
            L = int(ores["length"][i]); assert (out[i, :L] == oout[i, :L]).all(), f"slot {i} bytes differ"
    return res, out

@pytest.mark.parametrize("mcs", [11, 12, 13, 14])
def test_qam_mcs_clean_and_noisy(eng, mcs):
    for L, snr, cfo, chan in ((40, None, 0.0, ((1, 0), (0, 1))), (257, 38, 20e3, ((1.0, 0.3j), (-0.2, 0.9))), (1500, 34, -35e3, ((0.8, -0.4), (0.3j, 1.0))), (999, 20, 5e3, ((1.0, 0.3j), (-0.2, 0.9)))):
        iq0, iq1, ps = synth.make_frames_11n(6, psdu_len=L, mcs=mcs, snr_db=snr, lead=400, trail=200, cfo_hz=cfo, chan=chan)
        F, slot, _ = iq0.shape
        res, out = _compare(eng, iq0.reshape(-1, 2), iq1.reshape(-1, 2), np.arange(F) * slot, np.full(F, slot))
        if snr is None or snr >= 34:
            assert (res["status"] == 1).all() and (out[:, :L] == ps).all()

def test_qam_stage_taps(eng):
    for mcs, L in ((11, 120), (12, 300), (13, 555), (14, 1111)):
        iq0, iq1, ps = synth.make_frames_11n(1, psdu_len=L, mcs=mcs, snr_db=30, lead=400, trail=200, cfo_hz=12e3)
        o = oracle_py.rx11n_taps(iq0[0], iq1[0]); nd = o["ndata"]
        g = eng.rx11n_taps(iq0[0], iq1[0], [0], [iq0.shape[1]], max_sym=nd)
        assert g["res"]["status"][0] == o["res"]["status"] == 1
        assert (g["hinv"][0] == o["hinv"]).all() and (g["theta"][0, :nd] == o["theta"]).all() and (g["eq"][0][:, :nd] == o["eq"]).all()
        ns = len(o["soft"]); assert ns == nd * 104 * synth.HT_MCS[mcs][0] and (g["soft"][0, :ns] == o["soft"]).all()

def test_mixed_mcs_batch_and_truncation(eng):
    """All seven MCS in one batch (three Viterbi code rates), one slot cut short, one with a broken FCS."""
    s0, s1, want = [], [], []
    for mcs in (8, 9, 10, 11, 12, 13, 14, 13, 12):
        a, b, ps = synth.make_frames_11n(1, psdu_len=200 + 37 * mcs, mcs=mcs, snr_db=36, lead=400, trail=200, seed0=0x5000 + mcs)
        s0.append(a[0]); s1.append(b[0]); want.append(ps[0])
    s0[7] = s0[7][: len(s0[7]) // 2 // 28 * 28]; s1[7] = s1[7][: len(s0[7])]                    # truncated
    s0[8] = s0[8].copy(); s0[8][2600:2700] = 0; s1[8] = s1[8].copy(); s1[8][2600:2700] = 0       # damaged
    off = np.cumsum([0] + [len(s) for s in s0[:-1]]); ln = np.array([len(s) for s in s0])
    res, out = _compare(eng, np.concatenate(s0), np.concatenate(s1), off, ln)
    assert (res["status"][:7] == 1).all() and (res["mcs"][:7] == np.arange(8, 15)).all()
    for i in range(7): assert (out[i, :len(want[i])] == want[i]).all()
    assert res["status"][7] == oracle_py.E_NO_FRAME and res["status"][8] != 1

@pytest.mark.parametrize("mcs", [11, 12, 13, 14])
def test_tx_matches_oracle_bit_exact(eng, mcs):
    rng = np.random.default_rng(mcs + 40)
    lens = [1, 2, 3, 13, 14, 37, 200, 333, 1496, 57, 1000, 2000]
    pay = [rng.integers(0, 256, L).astype(np.uint8) for L in lens]
    seeds = np.array([0xAB, 0x5B, 0x02, 0x80, 0xFE, 0x13, 0xFF, 0x6D, 0x00, 0x01, 0xAB, 0x7F], np.uint8)
    o0, o1, ns = eng.tx11n_batch(pay, mcs, seeds=seeds)
    for i, p in enumerate(pay):
        w0, w1 = oracle_py.tx11n_modulate(p, mcs, int(seeds[i]))
        assert ns[i] == len(w0), (i, ns[i], len(w0))
        for got, want, name in ((o0[i], w0, "stream 0"), (o1[i], w1, "stream 1")):
            bad = np.nonzero((got[:len(want)] != want).any(1))[0]
            assert len(bad) == 0, (mcs, lens[i], name, bad[:10], got[bad[:4]], want[bad[:4]])
            assert (got[len(want):] == 0).all()

@pytest.mark.parametrize("mcs,L,F", [(11, 1496, 64), (12, 300, 64), (13, 1496, 128), (14, 1496, 128)])
def test_loopback_tx_to_rx_on_device(eng, mcs, L, F):
    import torch
    rng = np.random.default_rng(mcs)
    pay = rng.integers(0, 256, (F, L)).astype(np.uint8)
    d_pay = torch.from_numpy(pay.reshape(-1)).cuda()
    d_off = torch.arange(F, dtype=torch.int64, device="cuda") * L; d_len = torch.full((F,), L, dtype=torch.int32, device="cuda")
    nsym = -(-((L + 4) * 8 + 22) // synth.HT_MCS[mcs][2]) + 1
    slot = (400 + 1600 + 160 * nsym + 300 + 27) // 28 * 28
    d0 = torch.empty((F, slot, 2), dtype=torch.int16, device="cuda"); d1 = torch.empty_like(d0); st = torch.cuda.current_stream().cuda_stream
    eng.tx11n_raw(d_pay.data_ptr(), F * L, d_off.data_ptr(), d_len.data_ptr(), 0, F, mcs, 400, d0.data_ptr(), d1.data_ptr(), slot, 0, st)
    s_off = torch.arange(F, dtype=torch.int64, device="cuda") * slot; s_len = torch.full((F,), slot, dtype=torch.int32, device="cuda")
    d_out = torch.zeros((F, 1536), dtype=torch.uint8, device="cuda"); d_res = torch.zeros((F, 7), dtype=torch.int32, device="cuda")
    eng.rx11n_raw(d0.data_ptr(), d1.data_ptr(), F * slot, s_off.data_ptr(), s_len.data_ptr(), F, d_out.data_ptr(), 1536, d_res.data_ptr(), st)
    torch.cuda.synchronize()
    res = d_res.cpu().numpy()
    assert (res[:, 0] == 1).all() and (res[:, 1] == mcs).all() and (res[:, 2] == L + 4).all(), res[:4]
    assert (d_out.cpu().numpy()[:, :L] == pay).all()

def test_default_engine_refuses_like_the_reference():
    e = api.Engine(0)                                   # ht_mcs_limit 11: PHY_11n.hpp:496-501
    oracle_py.set_ht_mcs_limit(11)
    try:
        iq0, iq1, ps = synth.make_frames_11n(2, psdu_len=300, mcs=13, snr_db=36, lead=400, trail=200)
        F, slot, _ = iq0.shape
        res, _ = e.rx11n_batch(iq0.reshape(-1, 2), iq1.reshape(-1, 2), np.arange(F) * slot, np.full(F, slot))
        ores, _ = oracle_py.rx11n_batch(iq0.reshape(-1, 2), iq1.reshape(-1, 2), np.arange(F) * slot, np.full(F, slot), out_stride=1536)
        assert (res["status"] == oracle_py.E_PLCP_FAIL).all() and (res["status"] == ores["status"]).all() and (res["mcs"] == ores["mcs"]).all()
        with pytest.raises(api.Sb200Error): e.set_option("ht_mcs_limit", 16)
    finally:
        oracle_py.set_ht_mcs_limit(15)
