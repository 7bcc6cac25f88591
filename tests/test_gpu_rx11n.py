"""GPU parity tests for the 802.11n 2x2 path (pytest -m gpu): CUDA through the C ABI against the CPU oracle on the same IQ."""
import numpy as np, pytest
import oracle_py
from sora_b200 import api, synth

pytestmark = pytest.mark.gpu

@pytest.fixture(scope="module")
def eng():
    return api.Engine(0)

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

@pytest.mark.parametrize("mcs", [8, 9, 10])
def test_all_mcs_clean_and_noisy(eng, mcs):
    for L, snr, cfo, chan in ((40, None, 0.0, ((1, 0), (0, 1))), (257, 30, 20e3, ((1.0, 0.3j), (-0.2, 0.9))), (1500, 24, -35e3, ((0.8, -0.4), (0.3j, 1.0))), (999, 14, 5e3, ((1.0, 0.3j), (-0.2, 0.9)))):
        iq0, iq1, ps = synth.make_frames_11n(6, psdu_len=L, mcs=mcs, snr_db=snr, lead=400, trail=200, cfo_hz=cfo, chan=chan)
        F, slot, _ = iq0.shape
        res, out = _compare(eng, iq0.reshape(-1, 2), iq1.reshape(-1, 2), np.arange(F) * slot, np.full(F, slot))
        if snr is None or snr >= 24:
            assert (res["status"] == 1).all() and (out[:, :L] == ps).all()

def test_stage_taps_11n(eng):
    for mcs, L in ((8, 120), (9, 300), (10, 555)):
        iq0, iq1, ps = synth.make_frames_11n(1, psdu_len=L, mcs=mcs, snr_db=22, lead=400, trail=200, cfo_hz=12e3)
        o = oracle_py.rx11n_taps(iq0[0], iq1[0]); nd = o["ndata"]
        g = eng.rx11n_taps(iq0[0], iq1[0], [0], [iq0.shape[1]], max_sym=nd)
        assert g["res"]["status"][0] == o["res"]["status"] == 1
        assert (g["sig"][0] == o["sig"]).all()
        assert (g["siso"][0] == o["siso"]).all()
        assert (g["hinv"][0] == o["hinv"]).all()
        assert (g["theta"][0, :nd] == o["theta"]).all()
        assert (g["eq"][0][:, :nd] == o["eq"]).all()
        ns = len(o["soft"]); assert (g["soft"][0, :ns] == o["soft"]).all()

def test_failures_and_edges_11n(eng):
    rng = np.random.default_rng(11)
    iq0, iq1, ps = synth.make_frames_11n(8, psdu_len=300, mcs=9, snr_db=28, lead=400, trail=200)
    F, slot, _ = iq0.shape
    a, b = iq0.copy(), iq1.copy()
    a[0, 3000:3300] = rng.integers(-4000, 4000, (300, 2)); b[0, 3000:3300] = rng.integers(-4000, 4000, (300, 2))        # payload hit: CRC
    s = 400 + 640 + 160
    a[1, s:s + 160] = rng.integers(-4000, 4000, (160, 2)); b[1, s:s + 160] = rng.integers(-4000, 4000, (160, 2))          # HT-SIG1 hit: PLCP
    a[2, 400 + 640:400 + 800] = rng.integers(-4000, 4000, (160, 2)); b[2, 400 + 640:400 + 800] = rng.integers(-4000, 4000, (160, 2))   # L-SIG hit
    a[3] = 0; b[3] = 0                                                                                                      # silence
    a[4] = rng.normal(0, 3000, a[4].shape).astype(np.int16); b[4] = rng.normal(0, 3000, b[4].shape).astype(np.int16)        # noise only
    a[5, 400 + 640 + 6 * 160:] = 0; b[5, 400 + 640 + 6 * 160:] = 0                                                          # signal stops after the HT-LTFs
    off = np.arange(F) * slot; ln = np.full(F, slot); ln[6] = slot // 2; ln[7] = 27                                         # truncated capture, shorter than one block
    res, _ = _compare(eng, a.reshape(-1, 2), b.reshape(-1, 2), off, ln)
    assert res["status"][0] == oracle_py.E_CRC32_FAIL and res["status"][1] == oracle_py.E_PLCP_FAIL
    # a legacy 802.11a frame: L-SIG parses, HT-SIG does not
    iq, _ = synth.make_frames(2, psdu_len=200, rate_kbps=6000, snr_db=30, lead=400, trail=200, gain=0.6)
    F, slot, _ = iq.shape
    res, _ = _compare(eng, iq.reshape(-1, 2), iq.reshape(-1, 2), np.arange(F) * slot, np.full(F, slot))
    assert (res["status"] == oracle_py.E_PLCP_FAIL).all()

def test_11n_device_resident_and_many_slots(eng):
    import torch
    iq0, iq1, ps = synth.make_frames_11n(16, psdu_len=700, mcs=10, snr_db=26, lead=400, trail=200)
    F, slot, _ = iq0.shape; rep = 24; n = F * rep
    d0 = torch.from_numpy(iq0.reshape(F, -1)).cuda().repeat(rep, 1).contiguous(); d1 = torch.from_numpy(iq1.reshape(F, -1)).cuda().repeat(rep, 1).contiguous()
    d_off = torch.arange(n, dtype=torch.int64, device="cuda") * slot; d_len = torch.full((n,), slot, dtype=torch.int32, device="cuda")
    d_out = torch.zeros((n, 1536), dtype=torch.uint8, device="cuda"); d_res = torch.zeros((n, 7), dtype=torch.int32, device="cuda")
    eng.rx11n_raw(d0.data_ptr(), d1.data_ptr(), n * slot, d_off.data_ptr(), d_len.data_ptr(), n, d_out.data_ptr(), 1536, d_res.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ores, oout = oracle_py.rx11n_batch(iq0.reshape(-1, 2), iq1.reshape(-1, 2), np.arange(F) * slot, np.full(F, slot), out_stride=1536)
    st = d_res[:, 0].cpu().numpy().astype(np.uint32).reshape(rep, F)
    assert (st == ores["status"][None, :]).all() and (ores["status"] == 1).all()
    assert (d_out.cpu().numpy().reshape(rep, F, 1536)[:, :, :700] == oout[None, :, :700]).all()


def _capture_11n(seed, specs):
    """A continuous two-antenna capture: HT-MF frames (mcs, psdu_len, snr, lead, trail, damage) back to back; damage = None,
    'data' (a stretch of a data symbol negated: CRC failure) or 'sig' (HT-SIG flattened: header refused)."""
    a, b = [], []
    for i, (mcs, L, snr, lead, trail, damage) in enumerate(specs):
        i0, i1, _ = synth.make_frames_11n(1, psdu_len=L, mcs=mcs, seed0=seed * 1000 + i, snr_db=snr, lead=lead, trail=trail)
        i0, i1 = i0[0].copy(), i1[0].copy()
        if damage == "data": i0[lead + 1700:lead + 1800] = -i0[lead + 1700:lead + 1800]; i1[lead + 1700:lead + 1800] = -i1[lead + 1700:lead + 1800]
        if damage == "sig": i0[lead + 840:lead + 1100] //= 16; i1[lead + 840:lead + 1100] //= 16
        a.append(i0); b.append(i1)
    return np.concatenate(a), np.concatenate(b)

def test_streams_match_oracle_driver(eng):
    """Continuous captures through sb200_rx11n_streams == the restated RxThread loop (oracle Rx11n::run): same events, same order, same
    positions, same bytes, with TCCA11n / MimoAutoCorr's history carried across frames exactly as the never-reset bricks keep it."""
    caps = [_capture_11n(1, [(8, 200, 30, 400, 300, None), (9, 500, 28, 250, 420, None), (10, 120, 30, 333, 200, None), (8, 60, 30, 401, 500, None)]),
            _capture_11n(2, [(9, 300, 26, 500, 180, "data"), (8, 100, 30, 180, 300, None), (10, 700, 30, 222, 100, "sig"), (9, 90, 30, 300, 300, None)]),
            _capture_11n(3, [(10, 1500, 24, 777, 64, None), (10, 40, 30, 140, 900, None)]),
            (np.zeros((3000, 2), np.int16), np.zeros((3000, 2), np.int16))]
    off = np.cumsum([0] + [(len(c[0]) + 3) // 4 * 4 for c in caps[:-1]]).astype(np.uint64)
    total = int(off[-1]) + len(caps[-1][0]); iq0 = np.zeros((total, 2), np.int16); iq1 = np.zeros((total, 2), np.int16)
    for o, c in zip(off, caps): iq0[int(o):int(o) + len(c[0])] = c[0]; iq1[int(o):int(o) + len(c[1])] = c[1]
    ln = np.array([len(c[0]) for c in caps], np.uint32)
    res, out, sidx, cnt = eng.rx11n_streams(iq0, iq1, off, ln, max_frames=8)
    nok = 0
    for s, c in enumerate(caps):
        ores, oout = oracle_py.rx11n_run(c[0], c[1], max_frames=8, out_stride=1536)
        assert cnt[s] == len(ores), (s, cnt[s], res[s, :cnt[s]], ores)
        vec_base = 0; prev = 0
        for k in range(len(ores)):
            for fld in ("status", "mcs", "length", "nsym", "cfo_est", "lsig_length"):
                assert res[s, k][fld] == ores[k][fld], (s, k, fld, res[s, k], ores[k])
            assert sidx[s, k] == ores[k]["sample_index"], (s, k, sidx[s, k], ores[k]["sample_index"])
            assert vec_base + res[s, k]["detect_index"] == ores[k]["detect_index"], (s, k)          # the oracle counts 20 Msps vectors since the capture began
            vec_base += 4 * ((int(sidx[s, k]) - prev) // 8); prev = int(sidx[s, k])
            if ores[k]["status"] in (1, oracle_py.E_CRC32_FAIL):
                L = int(ores[k]["length"]); assert res[s, k]["crc32"] == ores[k]["crc32"] and (out[s, k, :L] == oout[k, :L]).all(), (s, k)
            nok += int(ores[k]["status"] == 1)
    assert nok >= 8 and cnt[3] == 0
    # single-capture form of the same call and a max_frames cut-off
    r1, _, s1, c1 = eng.rx11n_streams(caps[1][0], caps[1][1], [0], [len(caps[1][0])], max_frames=2)
    assert c1[0] == 2 and (r1[0, :2] == res[1, :2]).all() and (s1[0, :2] == sidx[1, :2]).all()
