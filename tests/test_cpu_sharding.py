"""world_size-2 gloo test of the multi-GPU host logic (sora_b200/shard.py): contiguous slot blocks per rank, gather on
root, results identical to a single-process decode.  The per-rank decoder here is the CPU oracle (tests only)."""
import os, sys, subprocess, textwrap, numpy as np, pytest
from sora_b200 import shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

def test_shard_ranges_cover_everything():
    for n in (0, 1, 7, 64, 65, 1000):
        for w in (1, 2, 3, 8):
            spans = [shard.shard_range(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(hi - lo for lo, hi in spans) <= -(-n // w) if n else True

WORKER = textwrap.dedent('''
    import os, sys, numpy as np, torch, torch.distributed as dist
    sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
    import oracle_py
    from sora_b200 import shard, synth, api
    dist.init_process_group("gloo")
    iq, ps = synth.make_frames(7, psdu_len=120, rate_kbps=36000, snr_db=26, seed0=11)
    F, slot, _ = iq.shape
    off = np.arange(F, dtype=np.uint64) * slot; ln = np.full(F, slot, np.uint32)
    def dec(iq2, o, l):
        r, b = oracle_py.rx11a_batch(iq2, o, l, out_stride=128)
        rr = np.zeros(len(r), api.RESULT_DTYPE)
        for k in api.RESULT_DTYPE.names: rr[k] = r[k]
        return rr, b
    res, out = shard.decode_sharded(dec, iq.reshape(-1, 2), off, ln, dist)
    if dist.get_rank() == 0:
        ref, refo = dec(iq.reshape(-1, 2), off, ln)
        assert (res == ref).all() and (out == refo).all() and (res["status"] == 1).all()
        assert (out[:, :120] == ps).all()
        print("SHARD_OK")
    else:
        assert res is None
    dist.destroy_process_group()
''')

def test_two_rank_gloo_gather(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % (ROOT, ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29631", str(script)], capture_output=True, text=True, timeout=300, env=env)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert "SHARD_OK" in p.stdout
