#!/usr/bin/env python3
"""bench_extra.py — the BASELINE.json configurations that are not the headline line of bench.py:

  --config viterbi   config #5: standalone K=7 soft Viterbi, rates 1/2, 2/3, 3/4, independent blocks of 20 022 information
                     bits (max 11a frame, 2500 B), device-resident soft values; coded bits/s and decoded Mbit/s
  --config 11b       config #3: 802.11b 11 Mbps CCK RX chain, PSDU 1500 B, 44 Msps, one frame per slot
  --config tx11a     SURVEY.md §8(f) rank 2: the 802.11a modulator on the device, 54 Mbps / 1500 B frames into config #2's slots
  --config tx11b     SURVEY.md §8(f) rank 2: the 802.11b modulator on the device, 11 Mbps CCK / 1500 B frames at 44 Msps
  --config tx11n     the 802.11n two-stream modulator on the device, MCS 8 / 9 / 10, 1500 B frames: the input of config #4 made on the device
  --config 11n       config #4: 802.11n HT-MF 2x2 RX chain at MCS 8, 9, 10, PSDU 1500 B, 2 x 40 Msps, fixed 2x2 channel

Each prints one JSON line per measurement (same timing rules as bench.py: >= 3 warm-ups, CUDA events on the launch
stream, inputs larger than L2).  Results are checked against the CPU oracle on a sample before timing.
"""
import argparse, json, os, sys, time
import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

def peaks():
    try: return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception: return 6650.0

def bench_viterbi(args):
    import torch, oracle_py
    from sora_b200 import api, synth
    eng = api.Engine(0); dev = torch.device("cuda", 0); st = torch.cuda.current_stream()
    L = 2500
    for cr, rate, name in ((api.CR_12, (1, 2), "1/2"), (api.CR_23, (2, 3), "2/3"), (api.CR_34, (3, 4), "3/4")):
        nbits = 8 * L + 16 + 6; nbits += (-nbits) % 48
        U = 64
        rng = np.random.default_rng(cr)
        bits = rng.integers(0, 2, (U, nbits)).astype(np.uint8); bits[:, 8 * L + 16:] = 0
        A, B = synth.conv_encode(bits); coded = synth.puncture(A, B, rate)
        soft = np.where(coded > 0, rng.integers(5, 8, coded.shape), rng.integers(0, 3, coded.shape)).astype(np.uint8)
        flip = rng.random(coded.shape) < 0.03
        soft = np.where(flip, rng.integers(0, 8, coded.shape), soft).astype(np.uint8)
        nsoft = soft.shape[1]; stride = (nsoft + 15) // 16 * 16
        NB = args.blocks
        sp = np.zeros((U, stride), np.uint8); sp[:, :nsoft] = soft
        d_soft = torch.from_numpy(sp).to(dev).repeat((NB + U - 1) // U, 1)[:NB].contiguous()
        d_out = torch.zeros((NB, L + 2 + 14), dtype=torch.uint8, device=dev)
        def step(): eng.viterbi_raw(d_soft.data_ptr(), stride, nsoft, NB, cr, L, d_out.data_ptr(), d_out.shape[1], stream=st.cuda_stream)
        step(); torch.cuda.synchronize()
        ref = oracle_py.viterbi_blocks(soft, cr, L)
        got = d_out[:U, :L + 2].cpu().numpy()
        assert (got == ref).all(), "GPU Viterbi differs from the oracle"
        for _ in range(3): step()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(st)
        for _ in range(args.steps): step()
        e1.record(st); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        coded_bits = NB * nsoft
        t0 = time.perf_counter(); ncpu = os.cpu_count() or 1
        rep = np.tile(soft, (max(1, 4096 // U), 1))
        oracle_py.viterbi_blocks(rep, cr, L, nthreads=ncpu); dt = time.perf_counter() - t0
        alg = coded_bits * (1.0 + (rate[0] / rate[1]) / 8.0)
        print(json.dumps({"metric": "standalone K=7 soft Viterbi throughput", "code_rate": name, "value": coded_bits / (ms * 1e-3) / 1e9, "unit": "G coded bits/s",
                          "decoded_mbit_s": NB * (8 * L + 16) / (ms * 1e-3) / 1e6, "ms_per_step": ms, "blocks": NB, "info_bits_per_block": 8 * L + 22,
                          "coded_bits_per_step": coded_bits, "n_gpus": 1, "dtype": "uint8 path metrics",
                          "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peaks(), "unit": "GB/s", "frac": alg / (ms * 1e-3) / 1e9 / peaks()},
                          "cpu_baseline": {"value": rep.shape[0] * nsoft / dt / 1e9, "unit": "G coded bits/s", "cores": ncpu, "kind": "port", "sample": f"{rep.shape[0]} blocks"},
                          "parity": "bit-exact vs oracle on %d blocks" % U}))

def bench_11b(args):
    import torch, oracle_py
    from sora_b200 import api, synth
    eng = api.Engine(0); dev = torch.device("cuda", 0); st = torch.cuda.current_stream()
    # The reference's CCK decoder is a pruned search (cck.hpp:262-769) and drops isolated symbols on some band-limited
    # waveforms even without noise; the timed set is made of slots the CPU oracle decodes FRAME_OK, so that the whole
    # chain (all 1500 bytes + CRC) is exercised.  GPU == oracle is asserted on all of them either way.
    iq, ps = synth.make_frames_11b(32, psdu_len=1500, rate_kbps=11000, snr_db=40, gain=0.15, lead=392, trail=200)
    F0, slot, _ = iq.shape
    ores0, _ = oracle_py.rx11b_batch(iq.reshape(-1, 2), np.arange(F0) * slot, np.full(F0, slot), out_stride=1504)
    keep = np.nonzero(ores0["status"] == 1)[0][:16]
    iq, ps = iq[keep], ps[keep]; U = len(keep)
    assert U >= 4, "too few decodable 11b slots"
    F0, slot, _ = iq.shape
    F = args.frames
    flat = iq.reshape(F0, -1)
    d_iq = torch.from_numpy(flat).to(dev).repeat((F + U - 1) // U, 1)[:F].contiguous()
    d_off = torch.arange(F, dtype=torch.int64, device=dev) * slot
    d_len = torch.full((F,), slot, dtype=torch.int32, device=dev)
    d_out = torch.zeros((F, 1504), dtype=torch.uint8, device=dev); d_res = torch.zeros((F, 6), dtype=torch.int32, device=dev)
    def step(): eng.rx11b_raw(d_iq.data_ptr(), F * slot, d_off.data_ptr(), d_len.data_ptr(), F, d_out.data_ptr(), 1504, d_res.data_ptr(), st.cuda_stream)
    step(); torch.cuda.synchronize()
    ores, oout = oracle_py.rx11b_batch(iq.reshape(-1, 2), np.arange(U) * slot, np.full(U, slot), out_stride=1504)
    assert (d_res[:U, 0].cpu().numpy().astype(np.uint32) == ores["status"]).all() and (ores["status"] == 1).all()
    assert (d_out[:U, :1499].cpu().numpy() == oout[:, :1499]).all() and (oout[:, :1499] == ps[:, :1499]).all()
    assert (d_res[:, 0].cpu().numpy() == 1).all()
    for _ in range(3): step()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(st)
    for _ in range(args.steps): step()
    e1.record(st); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    ncpu = os.cpu_count() or 1
    n = 2048
    t0 = time.perf_counter(); oracle_py.rx11b_batch(iq.reshape(-1, 2), (np.arange(n) % U) * slot, np.full(n, slot), out_stride=1504, nthreads=ncpu); dt = time.perf_counter() - t0
    alg = F * (slot * 4.0 + 1516)
    print(json.dumps({"metric": "802.11b 11 Mbps CCK RX PHY Msamples/s (IQ in, bits out)", "value": F * slot / (ms * 1e-3) / 1e6, "unit": "Msamples/s", "ms_per_step": ms,
                      "n_gpus": 1, "config": {"workload": "802.11b 11 Mbps CCK long preamble, PSDU 1500 B, 44 Msps (BASELINE config #3)", "slots_per_step": F, "samples_per_slot": int(slot), "unique_slots": U},
                      "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peaks(), "unit": "GB/s", "frac": alg / (ms * 1e-3) / 1e9 / peaks()},
                      "cpu_baseline": {"value": n * slot / dt / 1e6, "unit": "Msamples/s", "cores": ncpu, "kind": "port", "sample": f"{n} slots"},
                      "parity": "bytes and verdicts identical to the oracle on the %d unique slots" % U}))

def bench_11n(args):
    import torch, oracle_py
    from sora_b200 import api, synth
    eng = api.Engine(0); dev = torch.device("cuda", 0); st = torch.cuda.current_stream()
    for mcs in (8, 9, 10):
        U = 32
        iq0, iq1, ps = synth.make_frames_11n(U, psdu_len=1500, mcs=mcs, snr_db=30, lead=400, trail=200)      # fixed 2x2 channel [[1, 0.3j], [-0.2, 0.9]]
        F0, slot, _ = iq0.shape
        F = args.frames
        d0 = torch.from_numpy(iq0.reshape(U, -1)).to(dev).repeat((F + U - 1) // U, 1)[:F].contiguous()
        d1 = torch.from_numpy(iq1.reshape(U, -1)).to(dev).repeat((F + U - 1) // U, 1)[:F].contiguous()
        d_off = torch.arange(F, dtype=torch.int64, device=dev) * slot; d_len = torch.full((F,), slot, dtype=torch.int32, device=dev)
        d_out = torch.zeros((F, 1536), dtype=torch.uint8, device=dev); d_res = torch.zeros((F, 7), dtype=torch.int32, device=dev)
        def step(): eng.rx11n_raw(d0.data_ptr(), d1.data_ptr(), F * slot, d_off.data_ptr(), d_len.data_ptr(), F, d_out.data_ptr(), 1536, d_res.data_ptr(), st.cuda_stream)
        step(); torch.cuda.synchronize()
        ores, oout = oracle_py.rx11n_batch(iq0.reshape(-1, 2), iq1.reshape(-1, 2), np.arange(U) * slot, np.full(U, slot), out_stride=1536)
        assert (ores["status"] == 1).all() and (oout[:, :1500] == ps).all()
        assert (d_res[:, 0].cpu().numpy() == 1).all() and (d_out[:U, :1500].cpu().numpy() == oout[:, :1500]).all()
        for _ in range(3): step()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(st)
        for _ in range(args.steps): step()
        e1.record(st); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        kt = eng.last_kernel_times()
        ncpu = os.cpu_count() or 1; n = 1024
        t0 = time.perf_counter(); oracle_py.rx11n_batch(iq0.reshape(-1, 2), iq1.reshape(-1, 2), (np.arange(n) % U) * slot, np.full(n, slot), out_stride=1536, nthreads=ncpu); dt = time.perf_counter() - t0
        alg = F * (slot * 8.0 + 1516)
        print(json.dumps({"metric": "802.11n 2x2 RX PHY Msample-pairs/s (2 x IQ in, bits out)", "mcs": mcs, "value": F * slot / (ms * 1e-3) / 1e6, "unit": "Msample-pairs/s", "ms_per_step": ms,
                          "n_gpus": 1, "config": {"workload": "802.11n HT-MF 2x2, 20 MHz, PSDU 1500 B, 40 Msps per antenna, AWGN 30 dB, channel [[1,0.3j],[-0.2,0.9]] (BASELINE config #4)",
                                                   "slots_per_step": F, "sample_pairs_per_slot": int(slot), "unique_slots": U},
                          "kernel_ms": dict(zip(("carrier_sense", "ofdm_front_end", "viterbi_descramble_crc", "pack"), kt)) if kt is not None else None,
                          "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peaks(), "unit": "GB/s", "frac": alg / (ms * 1e-3) / 1e9 / peaks()},
                          "cpu_baseline": {"value": n * slot / dt / 1e6, "unit": "Msample-pairs/s", "cores": ncpu, "kind": "port", "sample": f"{n} slots"},
                          "parity": "bytes and verdicts identical to the oracle on the %d unique slots" % U}))

def bench_tx11a(args):
    import torch, oracle_py
    from sora_b200 import api
    eng = api.Engine(0); dev = torch.device("cuda", 0); st = torch.cuda.current_stream()
    F, L, rate = args.frames, 1496, 54000                    # PSDU 1500 B incl. FCS, as BASELINE config #2's frames
    rng = np.random.default_rng(7)
    pay = rng.integers(0, 256, (256, L)).astype(np.uint8)
    d_pay = torch.from_numpy(pay).to(dev).repeat((F + 255) // 256, 1)[:F].contiguous()
    d_off = torch.arange(F, dtype=torch.int64, device=dev) * L; d_len = torch.full((F,), L, dtype=torch.int32, device=dev)
    nsym = -(-(L + 7) * 8 // 216); slot = 32 + 640 + 160 * (1 + nsym) + 32                    # 9824 samples: config #2's slot
    d_iq = torch.zeros((F, slot, 2), dtype=torch.int16, device=dev)
    def step(): eng.tx11a_raw(d_pay.data_ptr(), F * L, d_off.data_ptr(), d_len.data_ptr(), 0, F, rate, 32, 16, d_iq.data_ptr(), slot, 0, st.cuda_stream)
    step(); torch.cuda.synchronize()
    got = d_iq[:4, 32:32 + 640 + 160 * (1 + nsym)].cpu().numpy()
    for i in range(4):
        assert (got[i] == oracle_py.tx11a_modulate(pay[i], rate, 0xFF, 0).astype(np.int16) << 8).all(), "GPU modulator differs from the oracle"
    # the receive path decodes what the transmit path made
    s_off = torch.arange(F, dtype=torch.int64, device=dev) * slot; s_len = torch.full((F,), slot, dtype=torch.int32, device=dev)
    d_out = torch.zeros((F, 1500), dtype=torch.uint8, device=dev); d_res = torch.zeros((F, 7), dtype=torch.int32, device=dev)
    eng.rx11a_raw(d_iq.data_ptr(), F * slot, s_off.data_ptr(), s_len.data_ptr(), F, d_out.data_ptr(), 1500, d_res.data_ptr(), st.cuda_stream); torch.cuda.synchronize()
    assert (d_res[:, 0] == 1).all() and (d_out[:256, :L].cpu().numpy() == pay).all()
    for _ in range(3): step()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(st)
    for _ in range(args.steps): step()
    e1.record(st); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    ncpu = os.cpu_count() or 1; n = 512
    import concurrent.futures as cf
    t0 = time.perf_counter()
    with cf.ThreadPoolExecutor(ncpu) as ex: list(ex.map(lambda i: oracle_py.tx11a_modulate(pay[i % 256], rate, 0xFF, 0), range(n)))
    dt = time.perf_counter() - t0
    alg = F * (L + slot * 4.0)
    print(json.dumps({"metric": "802.11a TX PHY Msamples/s (bytes in, IQ out)", "value": F * slot / (ms * 1e-3) / 1e6, "unit": "Msamples/s", "ms_per_step": ms, "n_gpus": 1,
                      "config": {"workload": "802.11a 54 Mbps modulator, PSDU 1500 B, COMPLEX16 slots of 9824 samples at 40 Msps (the input of BASELINE config #2 made on the device)", "frames_per_step": F},
                      "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peaks(), "unit": "GB/s", "frac": alg / (ms * 1e-3) / 1e9 / peaks()},
                      "cpu_baseline": {"value": n * slot / dt / 1e6, "unit": "Msamples/s", "cores": ncpu, "kind": "port", "sample": f"{n} frames (python threads around the C oracle)"},
                      "parity": "bit-exact vs the transmit oracle on 4 frames; every slot decodes FRAME_OK through the receive path"}))

def bench_tx11b(args):
    import torch, oracle_py
    from sora_b200 import api
    eng = api.Engine(0); dev = torch.device("cuda", 0); st = torch.cuda.current_stream()
    F, L, rate = args.frames, 1496, 11000                    # PSDU 1500 B incl. FCS at 11 Mbps CCK, as BASELINE config #3's frames
    rng = np.random.default_rng(11)
    pay = rng.integers(0, 256, (256, L)).astype(np.uint8)
    d_pay = torch.from_numpy(pay).to(dev).repeat((F + 255) // 256, 1)[:F].contiguous()
    d_off = torch.arange(F, dtype=torch.int64, device=dev) * L; d_len = torch.full((F,), L, dtype=torch.int32, device=dev)
    lead = 392; ns = ((24 * 88 + (L + 4) * 8 + 5) * 4 + 7) // 8 * 8
    slot = (lead + ns + 200 + 55) // 56 * 56                 # whole 28-sample blocks, multiple of 8
    d_iq = torch.zeros((F, slot, 2), dtype=torch.int16, device=dev)
    def step(): eng.tx11b_raw(d_pay.data_ptr(), F * L, d_off.data_ptr(), d_len.data_ptr(), F, rate, 0, lead, 16, d_iq.data_ptr(), slot, 0, st.cuda_stream)
    step(); torch.cuda.synchronize()
    got = d_iq[:4, lead:lead + ns].cpu().numpy()
    for i in range(4):
        assert (got[i] == oracle_py.tx11b_modulate(pay[i], rate).astype(np.int16) << 8).all(), "GPU modulator differs from the oracle"
    # the receive path decodes what the transmit path made
    s_off = torch.arange(F, dtype=torch.int64, device=dev) * slot; s_len = torch.full((F,), slot, dtype=torch.int32, device=dev)
    d_out = torch.zeros((F, 1504), dtype=torch.uint8, device=dev); d_res = torch.zeros((F, 6), dtype=torch.int32, device=dev)
    eng.rx11b_raw(d_iq.data_ptr(), F * slot, s_off.data_ptr(), s_len.data_ptr(), F, d_out.data_ptr(), 1504, d_res.data_ptr(), st.cuda_stream); torch.cuda.synchronize()
    ok = float((d_res[:, 0] == 1).float().mean())
    assert ok == 1.0 and (d_out[:256, :L].cpu().numpy() == pay).all(), ok
    for _ in range(3): step()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(st)
    for _ in range(args.steps): step()
    e1.record(st); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    ncpu = os.cpu_count() or 1; n = 512
    import concurrent.futures as cf
    t0 = time.perf_counter()
    with cf.ThreadPoolExecutor(ncpu) as ex: list(ex.map(lambda i: oracle_py.tx11b_modulate(pay[i % 256], rate), range(n)))
    dt = time.perf_counter() - t0
    alg = F * (L + slot * 4.0)
    print(json.dumps({"metric": "802.11b TX PHY Msamples/s (bytes in, IQ out)", "value": F * slot / (ms * 1e-3) / 1e6, "unit": "Msamples/s", "ms_per_step": ms, "n_gpus": 1,
                      "config": {"workload": "802.11b 11 Mbps CCK modulator, long preamble, PSDU 1500 B, COMPLEX16 slots of %d samples at 44 Msps (the input of BASELINE config #3 made on the device)" % slot, "frames_per_step": F},
                      "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peaks(), "unit": "GB/s", "frac": alg / (ms * 1e-3) / 1e9 / peaks()},
                      "cpu_baseline": {"value": n * slot / dt / 1e6, "unit": "Msamples/s", "cores": ncpu, "kind": "port", "sample": f"{n} frames (python threads around the C oracle)"},
                      "parity": "bit-exact vs the transmit oracle on 4 frames; every slot decodes FRAME_OK through the 802.11b receive path"}))

def bench_tx11n(args):
    import torch, oracle_py
    from sora_b200 import api
    eng = api.Engine(0); dev = torch.device("cuda", 0); st = torch.cuda.current_stream()
    F, L = args.frames, 1496
    rng = np.random.default_rng(12)
    pay = rng.integers(0, 256, (256, L)).astype(np.uint8)
    d_pay = torch.from_numpy(pay).to(dev).repeat((F + 255) // 256, 1)[:F].contiguous()
    d_off = torch.arange(F, dtype=torch.int64, device=dev) * L; d_len = torch.full((F,), L, dtype=torch.int32, device=dev)
    for mcs in (8, 9, 10):
        nd = {8: 52, 9: 104, 10: 156}[mcs]; nsym = -(-((L + 4) * 8 + 22) // nd) + 1
        lead = 400; slot = (lead + 1600 + 160 * nsym + 200 + 27) // 28 * 28
        d0 = torch.zeros((F, slot, 2), dtype=torch.int16, device=dev); d1 = torch.zeros_like(d0)
        def step(): eng.tx11n_raw(d_pay.data_ptr(), F * L, d_off.data_ptr(), d_len.data_ptr(), 0, F, mcs, lead, d0.data_ptr(), d1.data_ptr(), slot, 0, st.cuda_stream)
        step(); torch.cuda.synchronize()
        for i in range(3):
            w0, w1 = oracle_py.tx11n_modulate(pay[i], mcs)
            assert (d0[i, lead:lead + len(w0)].cpu().numpy() == w0).all() and (d1[i, lead:lead + len(w1)].cpu().numpy() == w1).all(), "GPU modulator differs from the oracle"
        s_off = torch.arange(F, dtype=torch.int64, device=dev) * slot; s_len = torch.full((F,), slot, dtype=torch.int32, device=dev)
        d_out = torch.zeros((F, 1536), dtype=torch.uint8, device=dev); d_res = torch.zeros((F, 7), dtype=torch.int32, device=dev)
        eng.rx11n_raw(d0.data_ptr(), d1.data_ptr(), F * slot, s_off.data_ptr(), s_len.data_ptr(), F, d_out.data_ptr(), 1536, d_res.data_ptr(), st.cuda_stream); torch.cuda.synchronize()
        assert bool((d_res[:, 0] == 1).all()) and (d_out[:256, :L].cpu().numpy() == pay).all()
        for _ in range(3): step()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(st)
        for _ in range(args.steps): step()
        e1.record(st); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        alg = F * (L + 2 * slot * 4.0)
        print(json.dumps({"metric": "802.11n 2-stream TX PHY Msample-pairs/s (bytes in, 2 x IQ out)", "mcs": mcs, "value": F * slot / (ms * 1e-3) / 1e6, "unit": "Msample-pairs/s", "ms_per_step": ms, "n_gpus": 1,
                          "config": {"workload": "802.11n HT-MF two-stream modulator, PSDU 1500 B, two COMPLEX16 slots of %d samples at 40 Msps per frame" % slot, "frames_per_step": F},
                          "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peaks(), "unit": "GB/s", "frac": alg / (ms * 1e-3) / 1e9 / peaks()},
                          "parity": "bit-exact vs the transmit oracle on 3 frames; every slot pair decodes FRAME_OK through the 802.11n receive path"}))
        del d0, d1

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", choices=["viterbi", "11b", "11n", "tx11a", "tx11b", "tx11n"], required=True)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--blocks", type=int, default=32768)
    ap.add_argument("--frames", type=int, default=32768)
    a = ap.parse_args()
    {"viterbi": bench_viterbi, "11b": bench_11b, "11n": bench_11n, "tx11a": bench_tx11a, "tx11b": bench_tx11b, "tx11n": bench_tx11n}[a.config](a)
