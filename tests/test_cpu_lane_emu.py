"""The one-lane-per-code-block Viterbi kernel (sora_b200/csrc/viterbi_k7_lane.cuh) WITHOUT a GPU: its device source is compiled for the host
by tests/cpp/lane_emu.cpp (the few CUDA intrinsics it uses are written out in C++ there, lanes run one after another) and compared with the
CPU oracle, bit for bit, on the cases the GPU tests use: clean and noisy code words, wrap-around garbage, block lengths that end anywhere in
the 6-step trellis phase / 8-step normalisation period / 6-column history block, both traceback windows, and code blocks of different lengths
side by side (the receive chains' per-frame path).  The GPU run of the same kernel is in tests/test_gpu_rx11a.py."""
import ctypes as C, os, shutil, subprocess
import numpy as np, pytest
import oracle_py
from sora_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "lane_emu.cpp"); CSRC = os.path.join(ROOT, "sora_b200", "csrc")
CR_12, CR_23, CR_34 = 0, 1, 2
pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="no host compiler")

@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("lane_emu") / "lane_emu.so")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-DSB_HOST_EMU", "-Wno-unknown-pragmas", "-I", CSRC, "-o", so, SRC])
    lib = C.CDLL(so)
    lib.lane_emu_viterbi.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint64,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    lib.lane_emu_viterbi.restype = C.c_int
    def run(soft, cr, L, depth=256, look=24, lens=None, nsofts=None, hb=6):
        soft = np.ascontiguousarray(soft, dtype=np.uint8); nb, ns = soft.shape
        stride = (ns + 15) // 16 * 16 + 16                              # the kernel fetches whole chunks: rows padded like the library's soft rows
        sp = np.zeros((nb, stride), np.uint8); sp[:, :ns] = soft
        ostride = L + 2 + 14
        out = np.zeros((nb, ostride), np.uint8); nraw = np.zeros(nb, np.uint32)
        lp = np.ascontiguousarray(lens, dtype=np.uint32) if lens is not None else None
        npp = np.ascontiguousarray(nsofts, dtype=np.uint32) if nsofts is not None else None
        rc = lib.lane_emu_viterbi(sp.ctypes.data, stride, ns, nb, cr, L, depth, look, out.ctypes.data, ostride, nraw.ctypes.data,
                                  lp.ctypes.data if lp is not None else None, npp.ctypes.data if npp is not None else None, hb)
        assert rc == 0
        return out[:, :L + 2], nraw
    return run

def _coded(rng, nblocks, L, rate):
    nbits = 8 * L + 16 + 6; nbits += (-nbits) % 48
    bits = rng.integers(0, 2, (nblocks, nbits)).astype(np.uint8); bits[:, 8 * L + 16:] = 0
    A, B = synth.conv_encode(bits)
    return bits, synth.puncture(A, B, rate)

HB = pytest.mark.parametrize("hb", (6, 8, 9))          # columns per history block: 6, 8, and 9 = 8 with the walk spread over the step loop

@HB
def test_code_words_clean_and_noisy(emu, hb):
    rng = np.random.default_rng(3)
    for cr, rate in ((CR_12, (1, 2)), (CR_23, (2, 3)), (CR_34, (3, 4))):
        L = 700
        bits, coded = _coded(rng, 5, L, rate)
        for flip in (0.0, 0.06, 0.5):
            soft = np.where(coded > 0, rng.integers(5, 8, coded.shape), rng.integers(0, 3, coded.shape)).astype(np.uint8)
            soft = np.where(rng.random(coded.shape) < flip, rng.integers(0, 8, coded.shape), soft).astype(np.uint8)
            g, nraw = emu(soft, cr, L, hb=hb)
            assert (g == oracle_py.viterbi_blocks(soft, cr, L)).all(), (cr, flip)
            assert (nraw == L + 2).all()
            if flip == 0.0:
                assert (np.unpackbits(g, axis=1, bitorder="little")[:, :8 * L + 16] == bits[:, :8 * L + 16]).all()

@HB
def test_wrap_and_ragged_stress(emu, hb):
    """tests/test_gpu_rx11a.py::test_viterbi_wrap_and_ragged_stress on the emulated kernel."""
    rng = np.random.default_rng(11)
    for cr, per in ((CR_12, 2), (CR_23, 3), (CR_34, 4)):
        steps_per = {2: 1, 3: 2, 4: 3}[per]
        for L in (1, 2, 3, 5, 17, 40, 101, 333, 1000):
            nbits = 8 * L + 16 + 6
            ngroups = -(-nbits // steps_per) + int(rng.integers(0, 5))
            ns = ngroups * per
            pats = [rng.integers(0, 8, (6, ns)), np.full((1, ns), 7), np.zeros((1, ns), int), np.tile([0, 7, 7, 0, 7], ns)[None, :ns], rng.integers(3, 5, (2, ns))]
            soft = np.concatenate(pats).astype(np.uint8)
            for depth, look in ((256, 24), (192, 36)):
                g, _ = emu(soft, cr, L, depth, look, hb=hb)
                o = oracle_py.viterbi_blocks(soft, cr, L, depth, look)
                assert (g == o).all(), (cr, L, depth, np.argwhere(g != o)[:4])

@HB
def test_blocks_of_different_lengths_side_by_side(emu, hb):
    """The receive chains hand every code block its own length and soft-byte count (FrameInfo): 37 blocks — a full warp and part of the next —
    whose triggers and ends fall at different times; a short input (a truncated frame) among them decodes what it has, like the oracle."""
    rng = np.random.default_rng(5)
    for cr, per, steps_per in ((CR_12, 2, 1), (CR_23, 3, 2), (CR_34, 4, 3)):
        lens = np.array([1 + (37 * i) % 411 for i in range(37)], np.uint32)
        nsofts = np.array([-(-(8 * int(L) + 22) // steps_per) * per for L in lens], np.uint32)
        nsofts[7] = nsofts[7] // (2 * per) * per                          # truncated: half of the soft values of block 7 are missing
        ns = int(nsofts.max())
        soft = rng.integers(0, 8, (37, ns)).astype(np.uint8)
        Lmax = int(lens.max())
        g, nraw = emu(soft, cr, Lmax, lens=lens, nsofts=nsofts, hb=hb)
        for i in range(37):
            L = int(lens[i]); n = int(nsofts[i])
            o = oracle_py.viterbi_blocks(soft[i:i + 1, :n], cr, L)[0]
            k = int(nraw[i])
            assert k <= L + 2 and (i == 7 or k == L + 2), (cr, i, k, L)
            assert (g[i, :k] == o[:k]).all(), (cr, i, L)

@HB
def test_short_windows(emu, hb):
    """Traceback windows shorter than a deferred walk lasts (the standalone decoder accepts any depth that is a multiple of 8): a trigger then
    finds the previous window's walk still in flight and has to finish it first; every window size must still give the oracle's bytes."""
    rng = np.random.default_rng(21)
    for cr, per in ((CR_12, 2), (CR_23, 3), (CR_34, 4)):
        L = 90; ns = (8 * L + 22 + 5) * per                               # more groups than the block needs, for every rate
        ns = ns // per * per
        soft = rng.integers(0, 8, (5, ns)).astype(np.uint8)
        for depth, look in ((8, 0), (8, 24), (16, 7), (64, 24), (128, 40), (248, 32)):
            g, _ = emu(soft, cr, L, depth, look, hb=hb)
            o = oracle_py.viterbi_blocks(soft, cr, L, depth, look)
            assert (g == o).all(), (cr, depth, look, np.argwhere(g != o)[:4])
