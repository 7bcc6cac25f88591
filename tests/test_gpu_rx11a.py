"""GPU parity tests (pytest -m gpu): the CUDA path, called through the C ABI, against the CPU oracle on the same IQ."""
import numpy as np, pytest
import oracle_py
from sora_b200 import api, synth
from sora_b200.dumpfile import load_dump
import os

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")

@pytest.fixture(scope="module")
def eng():
    return api.Engine(0)

def _slots(iq):
    F, slot, _ = iq.shape
    return iq.reshape(-1, 2), np.arange(F, dtype=np.uint64) * slot, np.full(F, slot, np.uint32)

def _compare(eng, iq2, off, ln, expect_ok=None):
    res, out = eng.rx11a_batch(iq2, off, ln)
    ores, oout = oracle_py.rx11a_batch(iq2, off, ln, out_stride=out.shape[1])
    for k in ("status", "rate_kbps", "length", "crc32", "nsym", "detect_index", "cfo_est"):
        m = (ores["status"] != oracle_py.E_NO_FRAME) if k != "status" else np.ones(len(res), bool)
        assert (res[k][m] == ores[k][m]).all(), (k, res[k], ores[k])
    for i in range(len(res)):
        if ores["status"][i] in (1, oracle_py.E_CRC32_FAIL):
            L = int(ores["length"][i])
            assert (out[i, :L] == oout[i, :L]).all(), f"frame {i} bytes differ"
    if expect_ok is not None:
        assert (res["status"] == 1).sum() >= expect_ok
    return res, out

@pytest.mark.parametrize("rate", sorted(synth.RATES))
def test_all_rates_clean_and_noisy(eng, rate):
    for snr in (None, 22):
        iq, ps = synth.make_frames(6, psdu_len=257, rate_kbps=rate, snr_db=snr, seed0=0x1000 + rate)
        res, out = _compare(eng, *_slots(iq), expect_ok=6 if snr is None else 0)
        if snr is None:
            for i in range(6):
                assert (out[i, :257] == ps[i]).all()

@pytest.mark.parametrize("stage", [1, 2])
def test_front_end_staging_variants(eng, stage):
    """Option front_stage: register double buffer (1) and bulk asynchronous copy into shared memory (2) feed the same samples to the same
    arithmetic as the direct loads (0): identical results on all rates, odd symbol counts, truncated and misaligned slots."""
    caps = []
    for i, rate in enumerate(sorted(synth.RATES)):
        iq, _ = synth.make_frames(2, psdu_len=61 + 97 * i, rate_kbps=rate, snr_db=25, seed0=0x5000 + rate, lead=37 + 3 * i, trail=50 + i)
        caps += [iq[0], iq[1][: len(iq[1]) - 400 * (i % 3)]]
    off = np.concatenate([[0], np.cumsum([len(c) + 1 for c in caps])[:-1]]).astype(np.uint64)     # + 1: slot starts at every alignment mod 4 words
    flat = np.zeros((int(off[-1]) + len(caps[-1]) + 8, 2), np.int16)
    for o, c in zip(off, caps): flat[int(o): int(o) + len(c)] = c
    ln = np.array([len(c) for c in caps], np.uint32)
    ref, refo = eng.rx11a_batch(flat, off, ln)
    eng.set_option("front_stage", stage)
    try:
        res, out = eng.rx11a_batch(flat, off, ln)
        assert (res == ref).all() and (out == refo).all()
        assert (ref["status"] == 1).sum() >= 8
    finally:
        eng.set_option("front_stage", 0)

def test_fsample6_golden(eng):
    iq = load_dump(os.path.join(GOLD, "fsample-6.dmp"))
    iq = (iq.astype(np.int32) << 2).astype(np.int16)     # 14-bit sample sign adjust (arx_fd.c:530 xmmAdjustSignBit)
    off = np.array([0], np.uint64); ln = np.array([len(iq)], np.uint32)
    res, out = _compare(eng, iq, off, ln, expect_ok=1)
    assert res["rate_kbps"][0] == 6000 and res["length"][0] == 1392

def test_low_snr_and_garbage(eng):
    rng = np.random.default_rng(7)
    iq, _ = synth.make_frames(16, psdu_len=120, rate_kbps=36000, snr_db=9, seed0=0x77)
    _compare(eng, *_slots(iq))
    noise = rng.normal(0, 3000, (4, 6000, 2)).astype(np.int16)
    _compare(eng, *_slots(noise))
    zeros = np.zeros((2, 2800, 2), np.int16)
    _compare(eng, *_slots(zeros))

def test_ragged_and_truncated(eng):
    iq, _ = synth.make_frames(5, psdu_len=400, rate_kbps=24000, snr_db=28, seed0=0x99)
    F, slot, _ = iq.shape
    flat = iq.reshape(-1, 2)
    off = np.arange(F, dtype=np.uint64) * slot
    ln = np.array([slot, slot - 1000, slot // 2, 700, 27], np.uint32)
    _compare(eng, flat, off, ln)

def test_cfo_and_gain(eng):
    for cfo in (-120e3, 40e3, 200e3):
        for gain in (0.5, 1.0):
            iq, _ = synth.make_frames(4, psdu_len=333, rate_kbps=54000, snr_db=30, cfo_hz=cfo, gain=gain, seed0=int(abs(cfo)) + 5)
            _compare(eng, *_slots(iq))

def test_stage_taps(eng):
    iq, _ = synth.make_frames(3, psdu_len=500, rate_kbps=54000, snr_db=27, cfo_hz=55e3, seed0=0x4242)
    flat, off, ln = _slots(iq)
    t = eng.rx11a_taps(flat, off, ln, max_sym=40)
    for i in range(3):
        o = oracle_py.rx11a_taps(iq[i], max_sym=40)
        ns = o["nsym"]
        assert (t["freq_coeffs"][i] == o["freq_coeffs"]).all()
        assert (t["chan_coeffs"][i] == o["chan_coeffs"]).all()
        assert (t["fft_out"][i, :ns] == o["fft_out"]).all()
        assert (t["equalized"][i, :ns] == o["equalized"]).all()
        used = [b for b in range(64) if (1 <= b <= 26) or (38 <= b <= 63)]
        assert (t["tracked"][i, :ns][:, used] == o["tracked"][:, used]).all()
        nsoft = len(o["soft"]) - 48          # oracle soft includes the 48 SIGNAL values first
        assert (t["soft"][i, :nsoft] == o["soft"][48:]).all()

def test_viterbi_standalone(eng):
    rng = np.random.default_rng(3)
    for cr, (num, den) in ((api.CR_12, (1, 2)), (api.CR_23, (2, 3)), (api.CR_34, (3, 4))):
        L = 700
        nbits = 8 * L + 16 + 6
        nbits += (-nbits) % (6 * 8)
        bits = rng.integers(0, 2, (5, nbits)).astype(np.uint8); bits[:, 8 * L + 16:] = 0
        A, B = synth.conv_encode(bits)
        coded = synth.puncture(A, B, (num, den))
        for flip in (0.0, 0.06, 0.5):
            soft = np.where(coded > 0, rng.integers(5, 8, coded.shape), rng.integers(0, 3, coded.shape)).astype(np.uint8)
            noise = rng.random(coded.shape) < flip
            soft = np.where(noise, rng.integers(0, 8, coded.shape), soft).astype(np.uint8)
            g = eng.viterbi_k7(soft, cr, L)
            o = oracle_py.viterbi_blocks(soft, cr, L)
            assert (g == o).all(), (cr, flip)
            if flip == 0.0:
                assert (np.unpackbits(g, axis=1, bitorder="little")[:, :8 * L + 16] == bits[:, :8 * L + 16]).all()

def test_viterbi_wrap_and_ragged_stress(eng):
    """Soft values that make the uint8 path metrics wrap (uniform garbage, all-ones, alternating extremes) and block lengths that end
    anywhere in a 6-step phase cycle / 8-step normalisation period: the device decoder must follow the reference's wrap, tie-break and
    windowing bit for bit at all three code rates and both windows (256/24 of 802.11a, 192/36 of 802.11n)."""
    rng = np.random.default_rng(11)
    for cr, per in ((api.CR_12, 2), (api.CR_23, 3), (api.CR_34, 4)):
        steps_per = {2: 1, 3: 2, 4: 3}[per]
        for L in (1, 2, 3, 5, 17, 40, 101, 333, 1000):
            nbits = 8 * L + 16 + 6
            ngroups = -(-nbits // steps_per) + int(rng.integers(0, 5))          # a few groups past the end as well
            ns = ngroups * per
            pats = [rng.integers(0, 8, (6, ns)), np.full((1, ns), 7), np.zeros((1, ns), int), np.tile([0, 7, 7, 0, 7], ns)[None, :ns],
                    rng.integers(3, 5, (2, ns))]
            soft = np.concatenate(pats).astype(np.uint8)
            for depth, look in ((256, 24), (192, 36)):
                g = eng.viterbi_k7(soft, cr, L, depth, look)
                o = oracle_py.viterbi_blocks(soft, cr, L, depth, look)
                assert (g == o).all(), (cr, L, depth, np.argwhere(g != o)[:4])

def test_brick_adaptor_graph(eng):
    """The header-only GPU brick inside a CREATE_BRICK_* graph driven like RxThread (sora_b200/brick/demo_graph.cpp)."""
    import subprocess, json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "sora_b200", "brick", "demo_graph")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.dirname(exe)])
    p = subprocess.run([exe, os.path.join(GOLD, "fsample-6.dmp"), "legacy14"], capture_output=True, text=True, timeout=120)
    ev = [json.loads(l) for l in p.stdout.splitlines() if l.startswith("{")]
    assert ev and ev[0]["error_code"] == "0x00000001" and ev[0]["rate_kbps"] == 6000 and ev[0]["length"] == 1392
    assert ev[0]["crc32"] == "0x80EF9B11" and ev[0]["bytes_out"] == 1392

def test_brick_adaptor_finds_every_frame_of_a_capture(tmp_path):
    """A dump file with several frames (one damaged, one with a broken SIGNAL) through the brick graph of demo_graph.cpp driven like RxThread:
    the adaptor decodes the window in continuous-capture mode and hands the driver one event per poll — the same events, in the same order,
    with the same fields as the CPU oracle's RxThread run.  Then 6 graph instances in 6 threads: their windows are decoded together."""
    import subprocess, json
    from sora_b200.dumpfile import write_dump
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "sora_b200", "brick", "demo_graph")
    if not os.path.exists(exe): subprocess.check_call(["make", "-C", os.path.dirname(exe)])
    cap = _mixed_stream(21)
    cap = cap[: len(cap) // 28 * 28]
    p = tmp_path / "multi.dmp"; write_dump(str(p), cap)
    ores, _ = oracle_py.rx11a_run(cap, max_frames=32, out_stride=2560)
    for extra in ([], ["--window", "20000"]):
        out = subprocess.run([exe, str(p)] + extra, capture_output=True, text=True, timeout=120).stdout
        ev = [json.loads(l) for l in out.splitlines() if l.startswith("{")]
        if not extra:                                   # one window = the whole file: exactly RxThread's event list
            assert len(ev) == len(ores) >= 6
            for e, o in zip(ev, ores):
                assert int(e["error_code"], 16) == int(o["status"]) and e["rate_kbps"] == o["rate_kbps"] and e["length"] == o["length"]
        else:                                           # windowed: the carrier sense restarts at window cuts, every good frame is still found
            good = [o for o in ores if o["status"] == 1]
            assert [e["length"] for e in ev if int(e["error_code"], 16) == 1] == [int(o["length"]) for o in good]
    iq, _ = synth.make_frames(70, psdu_len=30, rate_kbps=24000, snr_db=30, seed0=0xB00, lead=200, trail=220)   # more frames than one window's event list holds
    many = iq.reshape(-1, 2); many = many[: len(many) // 28 * 28]
    p2 = tmp_path / "many.dmp"; write_dump(str(p2), many)
    for extra in ([], ["--max-events", "16"]):          # the second run fills the event list of a window five times over: the window is re-armed, no frame is lost
        ev = [json.loads(l) for l in subprocess.run([exe, str(p2)] + extra, capture_output=True, text=True, timeout=120).stdout.splitlines() if l.startswith("{")]
        assert len(ev) == 70 and all(int(e["error_code"], 16) == 1 and e["length"] == 30 for e in ev)
    out = subprocess.run([exe, str(p), "--threads", "6"], capture_output=True, text=True, timeout=180).stdout
    summ = json.loads([l for l in out.splitlines() if l.startswith("{")][-1])
    assert summ["graph_instances"] == 6 and summ["events"] == 6 * len(ores) and summ["frames_ok"] == 6 * int((ores["status"] == 1).sum())

def test_chunked_pipeline_host_and_device(eng):
    """Large calls are cut into chunks pipelined over three streams (copy | front end | Viterbi): same results."""
    import torch
    iq, ps = synth.make_frames(11, psdu_len=180, rate_kbps=48000, snr_db=27, seed0=0xC0)
    flat, off, ln = _slots(iq)
    ref, refo = eng.rx11a_batch(flat, off, ln)
    eng.set_option("chunk_frames", 3); eng.set_option("chunk_frames_device", 4)
    try:
        res, out = eng.rx11a_batch(flat, off, ln)                      # host IQ: staged through the double buffer
        assert (res == ref).all() and (out == refo).all()
        for nthreads in (1, 3):                                        # host-side TDownSample2: only the even samples cross PCIe
            eng.set_option("host_decimate", nthreads)
            res, out = eng.rx11a_batch(flat, off, ln)
            assert (res == ref).all() and (out == refo).all()
            lo = ln.copy(); lo[::2] -= 29; lo[1] = 27; lo[3] = 700     # ragged (odd) slot lengths, a slot shorter than one block, a truncated one
            r1, o1 = eng.rx11a_batch(flat, off, lo)
            eng.set_option("host_decimate", 0); r0, o0 = eng.rx11a_batch(flat, off, lo); eng.set_option("host_decimate", nthreads)
            assert (r1 == r0).all() and (o1 == o0).all()
        # the mix of gathered chunks and chunks sent as they are: forced alternation, the adaptive rule (twice: the second call has its
        # estimates), every chunk gathered — same results, and sb200_last_transfer reports what really crossed the link
        slot_bytes = int(ln[0]) * 4; half_bytes = (int(ln[0]) + 1) // 2 * 4
        for mix, ngath in ((2, 2), (1, None), (1, None), (0, 4)):
            eng.set_option("host_decimate_mix", mix)
            res, out = eng.rx11a_batch(flat, off, ln)
            assert (res == ref).all() and (out == refo).all()
            nbytes, chunks, gathered = eng.last_transfer()
            assert chunks == 4 and (ngath is None or gathered == ngath)
            if mix == 0: assert nbytes == 11 * half_bytes
            if mix == 2: assert nbytes == (3 + 3) * slot_bytes + (3 + 2) * half_bytes          # chunks of 3, 3, 3, 2 slots: 0 and 2 as they are, 1 and 3 gathered
            assert 11 * half_bytes <= nbytes <= 11 * slot_bytes
        eng.set_option("host_decimate_mix", 1)
        eng.set_option("host_decimate", 0)
        res, out = eng.rx11a_batch(flat, off, ln); assert eng.last_transfer() == (11 * slot_bytes, 4, 0)
        dev = torch.device("cuda", 0)
        t_iq = torch.from_numpy(flat).to(dev); t_off = torch.from_numpy(off.astype(np.int64)).to(dev); t_len = torch.from_numpy(ln.astype(np.int32)).to(dev)
        t_out = torch.zeros((11, 256), dtype=torch.uint8, device=dev); t_res = torch.zeros((11, 7), dtype=torch.int32, device=dev)
        eng.set_option("slot_table_immutable", 1)
        for _ in range(2):                                             # second call hits the cached slot table
            eng.rx11a_raw(t_iq.data_ptr(), flat.shape[0], t_off.data_ptr(), t_len.data_ptr(), 11, t_out.data_ptr(), 256, t_res.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert (t_res.cpu().numpy().view(api.RESULT_DTYPE).reshape(-1) == ref).all()
        assert (t_out.cpu().numpy()[:, :180] == ps).all()
    finally:
        eng.set_option("chunk_frames", 4096); eng.set_option("chunk_frames_device", 0); eng.set_option("slot_table_immutable", 0); eng.set_option("host_decimate", 0); eng.set_option("host_decimate_mix", 1)

def test_44msps_capture(eng):
    """sb200_rx11a_batch_ex(sample_rate_mhz=44): on-device 11:10 resampler + the 40 Msps chain == oracle doing the same."""
    from test_cpu_oracle import _capture_44
    caps = [_capture_44(rate, 150 + 10 * i, 50 + i, snr_db=(None if i % 2 else 30)) for i, rate in enumerate((6000, 12000, 36000, 54000))]
    L = max(len(c[0]) for c in caps)
    flat = np.zeros((len(caps), L, 2), np.int16); ln = np.zeros(len(caps), np.uint32)
    for i, (iq, _) in enumerate(caps): flat[i, :len(iq)] = iq; ln[i] = len(iq)
    off = np.arange(len(caps), dtype=np.uint64) * L
    res, out = eng.rx11a_batch(flat.reshape(-1, 2), off, ln, sample_rate_mhz=44)
    for i, (iq, ps) in enumerate(caps):
        ores, oout = oracle_py.rx11a_run(oracle_py.resample_44_40(iq), max_frames=1)
        assert len(ores) == 1
        for k in ("status", "rate_kbps", "length", "crc32", "nsym", "detect_index", "cfo_est"):
            assert res[k][i] == ores[k][0], (i, k, res[k][i], ores[k][0])
        n = int(ores["length"][0]); assert (out[i, :n] == oout[0, :n]).all()
        if ores["status"][0] == 1: assert (out[i, :n] == ps).all()

def _mixed_stream(seed=5, dc=(180, -140)):
    """A continuous capture: frames of different rates / lengths back to back with gaps, a DC offset (so the carried
    CF_VecDC matters), one frame with a corrupted payload (CRC fail) and a bad SIGNAL field (PLCP fail)."""
    rng = np.random.default_rng(seed); parts = []
    specs = [(54000, 700), (6000, 120), (24000, 1500), (36000, 64), (48000, 2300), (9000, 333), (12000, 40), (18000, 999)]
    for i, (rate, L) in enumerate(specs):
        iq, _ = synth.make_frames(1, psdu_len=L, rate_kbps=rate, seed0=0x77000000 + i, snr_db=32, lead=int(rng.integers(40, 400)), trail=int(rng.integers(60, 500)), gain=0.5)
        s = iq[0].copy()
        if i == 2: s[5000:5200] = rng.integers(-3000, 3000, (200, 2))       # payload hit -> CRC32 fail
        if i == 4:
            lead_guess = np.nonzero(np.abs(s[:, 0].astype(np.int32)) > 400)[0][0]
            s[lead_guess + 640:lead_guess + 800] = rng.integers(-3000, 3000, (160, 2))   # SIGNAL symbol hit
        parts.append(s)
    st = np.concatenate(parts).astype(np.int32) + np.array(dc, np.int32)
    return np.clip(st, -32768, 32767).astype(np.int16)

def test_stream_mode_matches_rxthread(eng):
    for seed in (5, 6):
        st = _mixed_stream(seed)
        ores, oout = oracle_py.rx11a_run(st, max_frames=16, out_stride=2560)
        res, out, sidx = eng.rx11a_stream(st, max_frames=16)
        assert len(ores) >= 7 and len(res) == len(ores), (len(res), len(ores))
        for k in ("status", "rate_kbps", "length", "crc32", "nsym", "cfo_est"):
            assert (res[k] == ores[k]).all(), (k, res[k], ores[k])
        assert (sidx == ores["sample_index"]).all(), (sidx, ores["sample_index"])
        # the oracle counts 20 Msps vectors since the start of the capture, the library since the restart: each earlier
        # segment of n 40 Msps samples contributed 4*floor(n/8) of them (the decimator's queue is dropped on reset)
        seg = np.diff(np.concatenate([[0], ores["sample_index"].astype(np.int64)]))
        cum = np.concatenate([[0], np.cumsum(4 * (seg // 8))[:-1]])
        assert (res["detect_index"] + cum == ores["detect_index"]).all(), (res["detect_index"], cum, ores["detect_index"])
        for i in range(len(res)):
            if ores["status"][i] in (1, oracle_py.E_CRC32_FAIL):
                L = int(ores["length"][i]); assert (out[i, :L] == oout[i, :L]).all(), i
        assert (res["status"] == 1).sum() >= 5
    # max_frames bound and an empty / too-short capture
    res, _, _ = eng.rx11a_stream(_mixed_stream(5), max_frames=3); assert len(res) == 3
    res, _, _ = eng.rx11a_stream(np.zeros((2000, 2), np.int16)); assert len(res) == 0
    noise = np.random.default_rng(1).normal(0, 4000, (20000, 2)).astype(np.int16)      # spurious detections, no frames
    ores, _ = oracle_py.rx11a_run(noise, max_frames=32); res, _, sidx = eng.rx11a_stream(noise, max_frames=32)
    assert len(res) == len(ores) and (res["status"] == ores["status"]).all() and (sidx == ores["sample_index"]).all()

def test_rxblock_ingest_on_device(eng):
    """*.dmp bytes -> device gather (+ legacy 14-bit shift) -> continuous-capture decode: the golden frame comes out."""
    raw = np.fromfile(os.path.join(GOLD, "fsample-6.dmp"), dtype=np.uint8)
    iq = eng.rxblocks_unpack(raw, left_shift=2)
    ref = (load_dump(os.path.join(GOLD, "fsample-6.dmp")).astype(np.int32) << 2).astype(np.int16)
    assert iq.shape == ref.shape and (iq == ref).all()
    assert (eng.rxblocks_unpack(raw, 0) == load_dump(os.path.join(GOLD, "fsample-6.dmp"))).all()
    res, out, sidx = eng.rx11a_stream(iq, max_frames=4)
    assert len(res) == 1 and res["status"][0] == 1 and res["rate_kbps"][0] == 6000 and res["length"][0] == 1392 and res["crc32"][0] == 0x80EF9B11
    assert (out[0, :1392] == np.fromfile(os.path.join(GOLD, "fsample-6.psdu.bin"), dtype=np.uint8)).all()

def test_rxblock_descriptors(eng):
    """VStreamBits and TimeStamp of every RX_BLOCK (___RX_DESC) come back as the file holds them."""
    raw = np.fromfile(os.path.join(GOLD, "fsample-6.dmp"), dtype=np.uint8)
    vb, ts = eng.rxblocks_desc(raw)
    blk = raw[: len(raw) // 128 * 128].reshape(-1, 128)
    assert (vb == blk[:, 0:4].copy().view("<u4")[:, 0]).all() and (ts == blk[:, 12:16].copy().view("<u4")[:, 0]).all()
    rng = np.random.default_rng(2); syn = rng.integers(0, 256, (1000, 128)).astype(np.uint8)
    vb, ts = eng.rxblocks_desc(syn.reshape(-1))
    assert (vb == syn[:, 0:4].copy().view("<u4")[:, 0]).all() and (ts == syn[:, 12:16].copy().view("<u4")[:, 0]).all()

def test_legacy_c_api_shim():
    """CsFrameDemod's loop (kernel/bb/demod11/demod11a.cpp:81-200) written against include/sora_b200_legacy.h via ctypes."""
    import ctypes as C
    lib = api.load_library()
    class Stream(C.Structure): _fields_ = [("start", C.c_void_p), ("size", C.c_uint32), ("end", C.c_void_p), ("scan", C.c_void_p), ("mask", C.c_uint32)]
    class Ctx(C.Structure):
        _fields_ = [("SampleRate", C.c_uint), ("thr", C.c_uint32), ("maxblk", C.c_uint), ("minblk", C.c_uint), ("work", C.c_void_p), ("frame", C.c_void_p),
                    ("framemax", C.c_uint), ("framesize", C.c_uint), ("datarate", C.c_uint), ("frametype", C.c_uint), ("engine", C.c_void_p), ("events", C.c_void_p), ("shift", C.c_uint)]
    for f in ("BB11ARxCarrierSense", "BB11ARxFrameDemod"): getattr(lib, f).restype = C.c_int32
    # a capture with three frames (one of them damaged), stored as RX_BLOCKs
    parts = []
    for i, (rate, L) in enumerate(((24000, 300), (54000, 1000), (12000, 80))):
        iq, _ = synth.make_frames(1, psdu_len=L, rate_kbps=rate, seed0=0x99000000 + i, snr_db=30, lead=300, trail=300, gain=0.5)
        parts.append(iq[0])
    parts[1] = parts[1].copy(); parts[1][3000:3200] = 0
    cap = np.concatenate(parts); cap = cap[: len(cap) // 28 * 28]
    blocks = np.zeros((len(cap) // 28, 128), np.uint8); blocks[:, 0] = 1; blocks[:, 16:] = cap.reshape(-1, 28 * 2).view(np.uint8)
    oracle_py.lib().sbo_set_cca_threshold(C.c_uint32(250000))             # rxThreshold reaches the carrier sense of the engine: same value in the oracle
    try: ores, oout = oracle_py.rx11a_run(cap, max_frames=8, out_stride=2560)
    finally: oracle_py.lib().sbo_set_cca_threshold(C.c_uint32(0))
    work = C.c_uint32(1); frame = np.zeros(4096, np.uint8)
    st = Stream(); ctx = Ctx()
    lib.SoraGenRadioRxStreamOffline(C.byref(st), C.c_void_p(blocks.ctypes.data), C.c_uint32(blocks.size))
    lib.BB11ARxContextInit(C.byref(ctx), 40, 250000, 150, 112, C.byref(work))
    got = []
    for _ in range(10000):
        hr = lib.BB11ARxCarrierSense(C.byref(ctx), C.byref(st))
        if hr == 0x201:                                   # BB11A_OK_POWER_DETECTED
            lib.BB11APrepareRx(C.byref(ctx), C.c_void_p(frame.ctypes.data), 4096)
            hr = lib.BB11ARxFrameDemod(C.byref(ctx), C.byref(st)) & 0xFFFFFFFF
            got.append((hr, ctx.framesize, ctx.datarate, frame[:ctx.framesize].copy()))
        if st.scan == st.start and got: break             # wrapped: the offline loop ends here
    lib.BB11ARxContextCleanup(C.byref(ctx))
    assert len(got) == len(ores) == 3
    codes = {1: 0x202, oracle_py.E_CRC32_FAIL: 0x80006004, oracle_py.E_PLCP_FAIL: 0x80006002}
    for (hr, n, rate, by), o, ob in zip(got, ores, oout):
        assert hr == codes[int(o["status"])] and n == o["length"] and rate == o["rate_kbps"]
        if o["status"] in (1, oracle_py.E_CRC32_FAIL): assert (by == ob[:n]).all()
    assert got[0][0] == 0x202 and got[1][0] == 0x80006004 and got[2][0] == 0x202

def test_device_slot_table_is_checked_every_call(eng):
    """A device-resident slot table rewritten in place between two calls (same pointers): the second call sees the new table — a slot that
    now leaves the buffer is refused, a longer valid slot gets workspaces of its own size (no stale cached max length)."""
    import torch
    iq, ps = synth.make_frames(3, psdu_len=90, rate_kbps=12000, snr_db=30, seed0=0xD0)
    big, psb = synth.make_frames(1, psdu_len=1400, rate_kbps=12000, snr_db=30, seed0=0xD1)
    flat = np.concatenate([iq.reshape(-1, 2), big.reshape(-1, 2)]); slot = iq.shape[1]
    dev = torch.device("cuda", 0); st = torch.cuda.current_stream().cuda_stream
    t_iq = torch.from_numpy(flat).to(dev)
    t_off = torch.tensor([0, slot, 2 * slot], dtype=torch.int64, device=dev); t_len = torch.full((3,), slot, dtype=torch.int32, device=dev)
    t_out = torch.zeros((3, 1500), dtype=torch.uint8, device=dev); t_res = torch.zeros((3, 7), dtype=torch.int32, device=dev)
    call = lambda: eng.rx11a_raw(t_iq.data_ptr(), flat.shape[0], t_off.data_ptr(), t_len.data_ptr(), 3, t_out.data_ptr(), 1500, t_res.data_ptr(), st)
    call(); torch.cuda.synchronize()
    assert (t_res.cpu().numpy()[:, 0] == 1).all() and (t_out.cpu().numpy()[:, :90] == ps).all()
    t_off[2] = 3 * slot; t_len[2] = big.shape[1]                       # same pointers, a much longer (valid) third slot
    call(); torch.cuda.synchronize()
    r = t_res.cpu().numpy(); assert (r[:, 0] == 1).all() and r[2, 2] == 1400 and (t_out.cpu().numpy()[2, :1400] == psb[0]).all()
    t_len[1] = flat.shape[0]                                            # slot 1 now runs past the end of the buffer
    with pytest.raises(api.Sb200Error): call()
    t_len[1] = slot; t_off[0] = -2 ** 63                                # read as uint64 2^63: offset + length would wrap a 64-bit sum
    with pytest.raises(api.Sb200Error): call()

def test_44msps_more_than_65535_slots(eng):
    """sb200_rx11a_batch_ex at 44 Msps with more slots than a grid's y extent allows (65535): empty slots, one real capture among them."""
    from test_cpu_oracle import _capture_44
    cap, ps = _capture_44(24000, 120, 91, snr_db=30)
    n = 66000; ln = np.full(n, 56, np.uint32); off = np.arange(n, dtype=np.uint64) * 56
    flat = np.zeros((n * 56 + len(cap), 2), np.int16); flat[n * 56:] = cap
    off[65999] = n * 56; ln[65999] = len(cap)
    res, out = eng.rx11a_batch(flat, off, ln, sample_rate_mhz=44)
    assert res["status"][65999] == 1 and res["length"][65999] == 120 and (out[65999, :120] == ps).all()
    assert (res["status"][:65999] == api.FRAME_NONE).all()

def test_legacy_shim_more_events_than_one_pass_holds():
    """300 short frames in one RX stream: the legacy entry points keep finding them after the shim's 256-event pass (RxThread has no such limit)."""
    import ctypes as C
    lib = api.load_library()
    class Stream(C.Structure): _fields_ = [("start", C.c_void_p), ("size", C.c_uint32), ("end", C.c_void_p), ("scan", C.c_void_p), ("mask", C.c_uint32)]
    class Ctx(C.Structure):
        _fields_ = [("SampleRate", C.c_uint), ("thr", C.c_uint32), ("maxblk", C.c_uint), ("minblk", C.c_uint), ("work", C.c_void_p), ("frame", C.c_void_p),
                    ("framemax", C.c_uint), ("framesize", C.c_uint), ("datarate", C.c_uint), ("frametype", C.c_uint), ("engine", C.c_void_p), ("events", C.c_void_p), ("shift", C.c_uint)]
    for f in ("BB11ARxCarrierSense", "BB11ARxFrameDemod"): getattr(lib, f).restype = C.c_int32
    iq, ps = synth.make_frames(300, psdu_len=30, rate_kbps=24000, snr_db=30, seed0=0xE000, lead=200, trail=220)
    cap = iq.reshape(-1, 2); cap = cap[: len(cap) // 28 * 28]
    blocks = np.zeros((len(cap) // 28, 128), np.uint8); blocks[:, 0] = 1; blocks[:, 16:] = cap.reshape(-1, 28 * 2).view(np.uint8)
    work = C.c_uint32(1); frame = np.zeros(4096, np.uint8); st = Stream(); ctx = Ctx()
    lib.SoraGenRadioRxStreamOffline(C.byref(st), C.c_void_p(blocks.ctypes.data), C.c_uint32(blocks.size))
    lib.BB11ARxContextInit(C.byref(ctx), 40, 0, 150, 112, C.byref(work))
    got = []
    for _ in range(100000):
        hr = lib.BB11ARxCarrierSense(C.byref(ctx), C.byref(st))
        if hr == 0x201:
            lib.BB11APrepareRx(C.byref(ctx), C.c_void_p(frame.ctypes.data), 4096)
            hr = lib.BB11ARxFrameDemod(C.byref(ctx), C.byref(st)) & 0xFFFFFFFF
            got.append((hr, frame[:ctx.framesize].copy()))
        if st.scan == st.start and got: break
    lib.BB11ARxContextCleanup(C.byref(ctx))
    assert len(got) == 300 and all(hr == 0x202 for hr, _ in got)
    for (hr, by), p in zip(got, ps): assert (by == p).all()

def test_ofdm_bin_golden(eng):
    """The reference's own 24 Mbps modulator output decodes to 200 x 0x31 + FCS on the GPU as well."""
    raw = np.fromfile(os.path.join(GOLD, "ofdm.bin"), dtype=np.int8).reshape(-1, 2)
    iq = np.concatenate([np.zeros((400, 2), np.int16), raw.astype(np.int16) << 8, np.zeros((400, 2), np.int16)])
    res, out = _compare(eng, iq, [0], [len(iq)])
    assert res["status"][0] == 1 and res["rate_kbps"][0] == 24000 and res["length"][0] == 204
    assert (out[0, :200] == 0x31).all() and bytes(out[0, 200:204]) == bytes.fromhex("388d4983")

def test_reference_dummy_frames(eng):
    """kernel/sample/mac/Dot11ADummy*.txt (four waveforms of the reference's own modulator) in one batch of ragged slots: the GPU
    returns what the oracle returns, the two long ones give fsample-6's PSDU, the two short ones the known ACK bytes."""
    import golden_vectors as gv
    vec = gv.dummy_vectors(); names = sorted(vec)
    caps = [vec[n][0] for n in names]
    off = np.concatenate([[0], np.cumsum([len(c) for c in caps])[:-1]]).astype(np.uint64); ln = np.array([len(c) for c in caps], np.uint32)
    res, out = _compare(eng, np.concatenate(caps), off, ln, expect_ok=4)
    for i, n in enumerate(names):
        want = vec[n][2] if vec[n][2] is not None else bytes(gv.fsample6_psdu())
        assert res["rate_kbps"][i] == 6000 and res["length"][i] == len(want) and bytes(out[i, :len(want)]) == want, n

def test_legacy_tx_vectors_decode_on_the_device(eng):
    """tests/golden/legacy_tx/: frames made from the reference's own transmit tables at all eight rates (54 Mbps / 64-QAM / R = 3/4 included),
    one ragged batch: the device returns the oracle's results and every frame body."""
    rates = (6000, 9000, 12000, 18000, 24000, 36000, 48000, 54000)
    caps, bodies = [], []
    for kbps in rates:
        w = np.fromfile(os.path.join(GOLD, "legacy_tx", f"legacy_tx_{kbps}.i8"), np.int8).reshape(-1, 2)
        caps.append(np.concatenate([np.zeros((400, 2), np.int16), w.astype(np.int16) << 8, np.zeros((428, 2), np.int16)]))
        bodies.append(np.fromfile(os.path.join(GOLD, "legacy_tx", f"legacy_tx_{kbps}.bin"), np.uint8))
    off = np.concatenate([[0], np.cumsum([len(c) for c in caps])[:-1]]).astype(np.uint64); ln = np.array([len(c) for c in caps], np.uint32)
    res, out = _compare(eng, np.concatenate(caps), off, ln, expect_ok=8)
    for i, kbps in enumerate(rates):
        assert res["rate_kbps"][i] == kbps and res["length"][i] == len(bodies[i]) + 4 and (out[i, :len(bodies[i])] == bodies[i]).all()

def test_many_streams_at_once(eng):
    """sb200_rx11a_streams: a batch of continuous captures, each decoded with RxThread's sequential semantics."""
    caps = [_mixed_stream(s, dc=(40 * s - 200, 17 * s)) for s in range(7, 19)]
    caps.append(np.zeros((3000, 2), np.int16)); caps.append(caps[0][:5000].copy()); caps.append(np.zeros((10, 2), np.int16))
    off = np.concatenate([[0], np.cumsum([len(c) for c in caps])[:-1]]); ln = np.array([len(c) for c in caps])
    res, out, sidx, cnt = eng.rx11a_streams(np.concatenate(caps), off, ln, max_frames=6)
    for s, c in enumerate(caps):
        ores, oout = oracle_py.rx11a_run(c, max_frames=6, out_stride=2560)
        assert cnt[s] == len(ores), (s, cnt[s], len(ores))
        for k in ("status", "rate_kbps", "length", "crc32", "nsym", "cfo_est"):
            assert (res[s, :cnt[s]][k] == ores[k]).all(), (s, k)
        assert (sidx[s, :cnt[s]] == ores["sample_index"]).all()
        for i in range(cnt[s]):
            if ores["status"][i] in (1, oracle_py.E_CRC32_FAIL):
                L = int(ores["length"][i]); assert (out[s, i, :L] == oout[i, :L]).all()
    assert cnt.max() == 6 and cnt[-1] == 0 and cnt[-3] == 0

def test_page_locked_capture_buffer(eng):
    """sb200_host_alloc: a capture buffer the device can DMA from (the role of SoraURadioMapRxSampleBuf's mapping); decoding from it gives the
    same results as from ordinary host memory, and the brick batcher stages its windows in one."""
    import ctypes as C
    lib = api.load_library()
    lib.sb200_host_alloc.restype = C.c_void_p; lib.sb200_host_alloc.argtypes = [C.c_size_t]; lib.sb200_host_free.argtypes = [C.c_void_p]
    iq, ps = synth.make_frames(5, psdu_len=333, rate_kbps=36000, snr_db=28, seed0=0xA110C)
    flat, off, ln = _slots(iq)
    ref, refo = eng.rx11a_batch(flat, off, ln)
    nbytes = flat.size * 2
    p = lib.sb200_host_alloc(nbytes); assert p
    try:
        C.memmove(p, flat.ctypes.data, nbytes)
        res = np.zeros(len(off), dtype=api.RESULT_DTYPE); out = np.zeros((len(off), 2560), np.uint8)
        eng.rx11a_raw(p, flat.shape[0], off.ctypes.data, ln.ctypes.data, len(off), out.ctypes.data, 2560, res.ctypes.data)
        for k in ("status", "rate_kbps", "length", "crc32", "nsym", "detect_index", "cfo_est"): assert (res[k] == ref[k]).all(), k
        assert (out == refo).all() and (res["status"] == 1).all()
    finally:
        lib.sb200_host_free(p)
    assert lib.sb200_host_alloc(0) is None
