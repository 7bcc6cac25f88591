"""ctypes binding of oracle/libsora_oracle.so — TEST INFRASTRUCTURE (checker only, never the product path)."""
import ctypes as C, os, subprocess, numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None

class FrameResult(C.Structure):
    _fields_ = [("status", C.c_uint32), ("rate_kbps", C.c_uint32), ("length", C.c_uint32), ("crc32", C.c_uint32),
                ("nsym", C.c_uint32), ("sample_index", C.c_uint32), ("detect_index", C.c_uint32),
                ("cfo_est", C.c_int16), ("peak_index", C.c_uint16)]

RES_DTYPE = np.dtype([("status", "<u4"), ("rate_kbps", "<u4"), ("length", "<u4"), ("crc32", "<u4"), ("nsym", "<u4"),
                      ("sample_index", "<u4"), ("detect_index", "<u4"), ("cfo_est", "<i2"), ("peak_index", "<u2")])

E_FRAME_OK, E_CRC32_FAIL, E_PLCP_FAIL, E_NO_FRAME = 1, 0x80000006, 0x80000005, 0x8000F001

def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(ROOT, "oracle", "libsora_oracle.so")
        srcs = [os.path.join(ROOT, "oracle", f) for f in os.listdir(os.path.join(ROOT, "oracle")) if f.endswith((".cpp", ".h", ".inc"))]
        stale = lambda: not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
        if stale():                                     # several ranks / test workers may get here at once: one builds, the others wait for it
            import fcntl
            with open(os.path.join(ROOT, "oracle", ".build.lock"), "w") as lk:
                fcntl.flock(lk, fcntl.LOCK_EX)
                try:
                    if stale(): subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
                finally: fcntl.flock(lk, fcntl.LOCK_UN)
        _LIB = C.CDLL(so)
        _LIB.sbo_viterbi_block.restype = C.c_uint64
        _LIB.sbo_uatan2.restype = C.c_int16; _LIB.sbo_usin.restype = C.c_int16; _LIB.sbo_ucos.restype = C.c_int16
        _LIB.sbo_uatan2.argtypes = [C.c_int, C.c_int]; _LIB.sbo_usin.argtypes = [C.c_int16]; _LIB.sbo_ucos.argtypes = [C.c_int16]
        _LIB.sbo_sts_pattern.restype = C.POINTER(C.c_int16)
        _LIB.sbo_crc32.restype = C.c_uint32
        _LIB.sbo_viterbi_signal.restype = C.c_uint32
    return _LIB

def _p(a, t=C.c_void_p):
    return a.ctypes.data_as(t)

def rx11a_run(iq, max_frames=16, out_stride=4096):
    """iq: int16 array [n,2] (40 Msps). Returns (results structured array, bytes [nf, stride])."""
    iq = np.ascontiguousarray(iq, dtype=np.int16)
    res = np.zeros(max_frames, dtype=RES_DTYPE); out = np.zeros((max_frames, out_stride), dtype=np.uint8)
    n = lib().sbo_rx11a_run(_p(iq), C.c_uint64(iq.shape[0]), C.c_int(max_frames), _p(res), _p(out), C.c_uint64(out_stride))
    return res[:n], out[:n]

def rx11a_batch(iq, off, length, out_stride=2048, nthreads=1):
    iq = np.ascontiguousarray(iq, dtype=np.int16)
    off = np.ascontiguousarray(off, dtype=np.uint64); length = np.ascontiguousarray(length, dtype=np.uint32)
    nf = len(off)
    res = np.zeros(nf, dtype=RES_DTYPE); out = np.zeros((nf, out_stride), dtype=np.uint8)
    lib().sbo_rx11a_batch(_p(iq), _p(off), _p(length), C.c_uint32(nf), _p(res), _p(out), C.c_uint64(out_stride), C.c_int(nthreads))
    return res, out

def rx11a_batch_2t(iq, off, length, out_stride=2048, npipes=1):
    """The same batch through the reference's two-thread topology (front end | Viterbi thread per pipeline), npipes pipelines."""
    iq = np.ascontiguousarray(iq, dtype=np.int16)
    off = np.ascontiguousarray(off, dtype=np.uint64); length = np.ascontiguousarray(length, dtype=np.uint32)
    nf = len(off)
    res = np.zeros(nf, dtype=RES_DTYPE); out = np.zeros((nf, out_stride), dtype=np.uint8)
    lib().sbo_rx11a_batch_2t(_p(iq), _p(off), _p(length), C.c_uint32(nf), _p(res), _p(out), C.c_uint64(out_stride), C.c_int(npipes))
    return res, out

def rx11a_taps(iq, max_sym=600):
    iq = np.ascontiguousarray(iq, dtype=np.int16)
    res = np.zeros(1, dtype=RES_DTYPE)
    fc = np.zeros((64, 2), np.int16); cc = np.zeros((64, 2), np.int16)
    fo = np.zeros((max_sym, 64, 2), np.int16); eq = np.zeros_like(fo); tr = np.zeros_like(fo)
    soft = np.zeros(max_sym * 288, np.uint8); so = np.zeros(max_sym + 1, np.uint32)
    ns = lib().sbo_rx11a_taps(_p(iq), C.c_uint64(iq.shape[0]), _p(res), _p(fc), _p(cc), _p(fo), _p(eq), _p(tr), _p(soft), _p(so), C.c_int(max_sym))
    return dict(res=res[0], freq_coeffs=fc, chan_coeffs=cc, fft_out=fo[:ns], equalized=eq[:ns], tracked=tr[:ns],
                soft=soft[:so[ns]], soft_off=so[:ns + 1], nsym=ns)

def viterbi_block(soft, code_rate, frame_len_bytes, depth=256, lookahead=24):
    soft = np.ascontiguousarray(soft, dtype=np.uint8)
    out = np.zeros(frame_len_bytes + 2 + 64, dtype=np.uint8)
    n = lib().sbo_viterbi_block(_p(soft), C.c_uint64(len(soft)), C.c_int(code_rate), C.c_uint32(frame_len_bytes),
                                C.c_uint32(depth), C.c_uint32(lookahead), _p(out))
    return out[:n]

def viterbi_blocks(soft2d, code_rate, frame_len_bytes, depth=256, lookahead=24, nthreads=1):
    soft2d = np.ascontiguousarray(soft2d, dtype=np.uint8)
    nb, ns = soft2d.shape
    stride = frame_len_bytes + 2
    out = np.zeros((nb, stride), dtype=np.uint8)
    lib().sbo_viterbi_blocks(_p(soft2d), C.c_uint64(ns), C.c_uint32(nb), C.c_int(code_rate), C.c_uint32(frame_len_bytes),
                             C.c_uint32(depth), C.c_uint32(lookahead), _p(out), C.c_uint64(stride), C.c_int(nthreads))
    return out

def fft64(x):
    x = np.ascontiguousarray(x, dtype=np.int16); o = np.zeros((64, 2), np.int16); lib().sbo_fft64(_p(x), _p(o)); return o
def ifft64(x):
    x = np.ascontiguousarray(x, dtype=np.int16); o = np.zeros((64, 2), np.int16); lib().sbo_ifft64(_p(x), _p(o)); return o
def crc32(b):
    b = np.ascontiguousarray(b, dtype=np.uint8); return int(lib().sbo_crc32(_p(b), C.c_uint64(len(b))))

RES11B_DTYPE = np.dtype([("status", "<u4"), ("rate_kbps", "<u4"), ("length", "<u4"), ("crc32", "<u4"), ("sample_index", "<u4"), ("detect_vec", "<u4")])

def rx11b_run(iq, max_frames=8, out_stride=4096):
    iq = np.ascontiguousarray(iq, dtype=np.int16)
    res = np.zeros(max_frames, dtype=RES11B_DTYPE); out = np.zeros((max_frames, out_stride), dtype=np.uint8)
    n = lib().sbo_rx11b_run(_p(iq), C.c_uint64(iq.shape[0]), C.c_int(max_frames), _p(res), _p(out), C.c_uint64(out_stride))
    return res[:n], out[:n]

def rx11b_batch(iq, off, length, out_stride=4096, nthreads=1):
    iq = np.ascontiguousarray(iq, dtype=np.int16)
    off = np.ascontiguousarray(off, dtype=np.uint64); length = np.ascontiguousarray(length, dtype=np.uint32)
    nf = len(off)
    res = np.zeros(nf, dtype=RES11B_DTYPE); out = np.zeros((nf, out_stride), dtype=np.uint8)
    lib().sbo_rx11b_batch(_p(iq), _p(off), _p(length), C.c_uint32(nf), _p(res), _p(out), C.c_uint64(out_stride), C.c_int(nthreads))
    return res, out

def resample_44_40(iq44):
    iq44 = np.ascontiguousarray(iq44, dtype=np.int16)
    out = np.zeros_like(iq44)
    lib().sbo_resample_44_40.restype = C.c_uint64
    n = lib().sbo_resample_44_40(_p(iq44), C.c_uint64(iq44.shape[0]), _p(out))
    return out[:n].copy()


# ---- 802.11n 2x2 ------------------------------------------------------------------------------------------------------
RES11N_DTYPE = np.dtype([("status", "<u4"), ("mcs", "<u4"), ("length", "<u4"), ("crc32", "<u4"), ("nsym", "<u4"),
                         ("sample_index", "<u4"), ("detect_index", "<u4"), ("cfo_est", "<i2"), ("lsig_length", "<u2")])

def rx11n_run(iq0, iq1, max_frames=8, out_stride=2048):
    iq0 = np.ascontiguousarray(iq0, dtype=np.int16); iq1 = np.ascontiguousarray(iq1, dtype=np.int16)
    res = np.zeros(max_frames, dtype=RES11N_DTYPE); out = np.zeros((max_frames, out_stride), dtype=np.uint8)
    n = lib().sbo_rx11n_run(_p(iq0), _p(iq1), C.c_uint64(iq0.reshape(-1, 2).shape[0]), C.c_int(max_frames), _p(res), _p(out), C.c_uint64(out_stride))
    return res[:n], out[:n]

def rx11n_batch(iq0, iq1, off, length, out_stride=1536, nthreads=1):
    iq0 = np.ascontiguousarray(iq0, dtype=np.int16); iq1 = np.ascontiguousarray(iq1, dtype=np.int16)
    off = np.ascontiguousarray(off, dtype=np.uint64); length = np.ascontiguousarray(length, dtype=np.uint32); nf = len(off)
    res = np.zeros(nf, dtype=RES11N_DTYPE); out = np.zeros((nf, out_stride), dtype=np.uint8)
    lib().sbo_rx11n_batch(_p(iq0), _p(iq1), _p(off), _p(length), C.c_uint32(nf), _p(res), _p(out), C.c_uint64(out_stride), C.c_int(nthreads))
    return res, out

def rx11n_taps(iq0, iq1, max_sym=300):
    iq0 = np.ascontiguousarray(iq0, dtype=np.int16); iq1 = np.ascontiguousarray(iq1, dtype=np.int16)
    res = np.zeros(1, dtype=RES11N_DTYPE)
    siso = np.zeros((2, 64, 2), np.int16); hinv = np.zeros((4, 64, 2), np.int16)
    fo = np.zeros((2, max_sym, 64, 2), np.int16); eq = np.zeros((2, max_sym, 64, 2), np.int16)
    soft = np.zeros(max_sym * 624, np.uint8); nsoft = C.c_uint32(0); theta = np.zeros(max_sym, np.int16); sig = np.zeros(9, np.uint8); nfft = C.c_int(0)
    nd = lib().sbo_rx11n_taps(_p(iq0), _p(iq1), C.c_uint64(iq0.reshape(-1, 2).shape[0]), _p(res), _p(siso), _p(hinv), _p(fo), _p(eq), _p(soft),
                              C.byref(nsoft), _p(theta), _p(sig), C.c_int(max_sym), C.byref(nfft))
    return dict(res=res[0], siso=siso, hinv=hinv, fft_out=fo[:, :nfft.value], eq=eq[:, :nd], soft=soft[:nsoft.value], theta=theta[:nd], sig=sig, ndata=nd)

def tables11n():
    sc = np.zeros((65536, 2), np.int16); at = np.zeros(4097, np.int16); dm = np.zeros(256, np.uint8); c8 = np.zeros(256, np.uint8)
    di = np.zeros((4, 2, 312), np.uint16); ls = np.zeros(64, np.uint8); hs = np.zeros(64, np.uint8)
    lib().sbo_tables11n(_p(sc), _p(at), _p(dm), _p(c8), _p(di), _p(ls), _p(hs))
    d16 = np.zeros((2, 256), np.uint8); d64 = np.zeros((3, 288), np.uint8); lib().sbo_tables11n_qam(_p(d16), _p(d64))
    return dict(sincos=sc, atan=at, demap=dm, crc8=c8, deint=di, lltf_sign=ls, htltf_sign=hs, demap16=d16, demap64=d64)

def set_ht_mcs_limit(first_refused=11):
    """11 = the reference's HT-SIG parser as shipped (MCS 8..10); 15 = the 16-/64-QAM branches enabled (MCS 8..14)."""
    lib().sbo_set_ht_mcs_limit(C.c_uint32(first_refused))


# ---- 802.11a transmit (brick modulator restatement) ---------------------------------------------------------------------
def tx11a_modulate(payload, rate_kbps, seed=0xFF, tail_zeros=32):
    """payload: MPDU bytes without FCS.  Returns int8 [n, 2] complex samples at 40 Msps (what `demod11 -m` writes)."""
    payload = np.ascontiguousarray(payload, dtype=np.uint8); L = lib(); L.sbo_tx11a_modulate.restype = C.c_uint64
    nsym = L.sbo_tx11a_nsym(C.c_uint32(len(payload)), C.c_uint32(rate_kbps)); cap = 640 + 160 * (1 + nsym) + tail_zeros
    out = np.zeros((cap, 2), np.int8)
    n = L.sbo_tx11a_modulate(_p(payload), C.c_uint32(len(payload)), C.c_uint32(rate_kbps), C.c_uint8(seed), _p(out), C.c_uint64(cap), C.c_uint32(tail_zeros))
    assert n == cap, (n, cap)
    return out

def tx11a_legacy_modulate(body, rate_kbps, append_crc=True):
    """The reference's LEGACY transmitter (BB11ATxFrameMod; oracle/tx11a_legacy.cpp).  Returns int8 [n, 2]: 640 + 160 (1 + nsym) + 8 samples at 40 Msps."""
    body = np.ascontiguousarray(body, dtype=np.uint8); L = lib(); L.sbo_tx11a_legacy_modulate.restype = C.c_uint64
    pre = np.fromfile(os.path.join(ROOT, "tests", "golden", "preamble40_11a.i16"), np.int16)
    nsym = L.sbo_tx11a_legacy_nsym(C.c_uint32(len(body) + (4 if append_crc else 0)), C.c_uint32(rate_kbps)); assert nsym, "not an 802.11a rate"
    cap = 640 + 160 * (1 + nsym) + 8
    out = np.zeros((cap, 2), np.int8)
    n = L.sbo_tx11a_legacy_modulate(_p(body), C.c_uint32(len(body)), C.c_int(1 if append_crc else 0), C.c_uint32(rate_kbps), _p(pre), _p(out), C.c_uint64(cap))
    assert n == cap, (n, cap)
    return out

def ifft128(x):
    x = np.ascontiguousarray(x, dtype=np.int16); o = np.zeros((128, 2), np.int16); lib().sbo_ifft128(_p(x), _p(o)); return o


# ---- 802.11b transmit (brick modulator restatement) ---------------------------------------------------------------------
def tx11b_modulate(payload, rate_kbps, init_phase=0, return_phase=False):
    """payload: MPDU bytes without FCS.  Returns int8 [n, 2] complex samples at 44 Msps (the COMPLEX8 buffer TModSink fills)."""
    payload = np.ascontiguousarray(payload, dtype=np.uint8); L = lib(); L.sbo_tx11b_modulate.restype = C.c_uint64
    cap = L.sbo_tx11b_nsamples(C.c_uint32(len(payload)), C.c_uint32(rate_kbps)); assert cap, "not an 802.11b rate"
    out = np.zeros((cap, 2), np.int8)
    fin = C.c_uint32(0)
    n = L.sbo_tx11b_modulate(_p(payload), C.c_uint32(len(payload)), C.c_uint32(rate_kbps), C.c_uint32(init_phase), _p(out), C.c_uint64(cap), C.byref(fin))
    assert n == cap, (n, cap)
    return (out, fin.value) if return_phase else out

def tx11b_taps():
    o = np.zeros(20, np.int16); lib().sbo_tx11b_taps(_p(o)); return o


# ---- 802.11n transmit (brick modulator restatement) ---------------------------------------------------------------------
def tx11n_modulate(payload, mcs, seed=0xAB):
    """payload: MPDU bytes without FCS.  Returns two int16 [n, 2] streams at 40 Msps (what `demod11` writes to *_0.dmp / *_1.dmp)."""
    payload = np.ascontiguousarray(payload, dtype=np.uint8); L = lib(); L.sbo_tx11n_modulate.restype = C.c_uint64
    sig = C.c_uint32(0); nsym = L.sbo_tx11n_nsym(C.c_uint32(len(payload)), C.c_uint32(mcs), C.byref(sig)); assert nsym, "mcs must be 8 .. 14"
    cap = 640 + 480 + 480 + 160 * nsym
    o0 = np.zeros((cap, 2), np.int16); o1 = np.zeros((cap, 2), np.int16)
    n = L.sbo_tx11n_modulate(_p(payload), C.c_uint32(len(payload)), C.c_uint32(mcs), C.c_uint8(seed), _p(o0), _p(o1), C.c_uint64(cap))
    assert n == cap, (n, cap)
    return o0, o1

def tx11n_preamble_tables():
    a = np.zeros((320, 2), np.int16); b = np.zeros((320, 2), np.int16); c = np.zeros((160, 2), np.int16); d = np.zeros((160, 2), np.int16)
    lib().sbo_tx11n_preamble_tables(_p(a), _p(b), _p(c), _p(d)); return a, b, c, d


def fir37_legacy(chips, variant=0):
    """oracle/tx11b_legacy.cpp: BB11BPMDSpreadFIR4SSE (variant 0) / BB11BPMDSpreadFIR4ASM (variant 1) restated.  int8 [n,2] -> int8 [n,2] (n % 8 == 0)."""
    x = np.ascontiguousarray(chips, dtype=np.int8).reshape(-1, 2); out = np.zeros_like(x)
    lib().sbo_fir37_legacy(C.c_void_p(x.ctypes.data), C.c_uint32(len(x)), C.c_int(variant), C.c_void_p(out.ctypes.data))
    return out

_REF_FIR = None
def ref_fir37_available():
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libfir37_ref.so"))
def ref_fir37(chips):
    """oracle/_ref/libfir37_ref.so: the reference's own FIR37SSE_INTRINSIC body + coefficient table, compiled by oracle/build_ref.sh."""
    global _REF_FIR
    if _REF_FIR is None: _REF_FIR = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libfir37_ref.so"))
    x = np.ascontiguousarray(chips, dtype=np.int8).reshape(-1, 2); out = np.zeros_like(x)
    _REF_FIR.ref_fir37(C.c_void_p(x.ctypes.data), C.c_uint(len(x)), C.c_void_p(out.ctypes.data))
    return out
