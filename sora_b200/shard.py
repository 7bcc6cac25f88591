"""Multi-GPU plumbing for the 802.11 RX path: capture slots are independent units, so a batch is split
into contiguous blocks of ceil(F/G) slots per rank, every rank decodes its block with its own engine, and only the
results travel: verdict records and PSDU bytes are gathered on the root (torch.distributed gather — NCCL over NVLink on
GPUs, gloo in the CPU tests).  No collective touches the data path itself.
"""
import numpy as np

def shard_range(nslots, world, rank):
    """Contiguous block [lo, hi) of slot indices owned by `rank` (keeps each rank's IQ slab contiguous)."""
    per = -(-nslots // world)
    lo = min(rank * per, nslots)
    return lo, min(lo + per, nslots)

def gather_results(res, out, nslots, dist=None, device=None, root=0):
    """res: structured array [n_local] (api.RESULT_DTYPE), out: uint8 [n_local, stride] for this rank's block.
    Returns (res_all, out_all) on the root, (None, None) elsewhere.  Blocks are padded to equal size for the gather."""
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return res, out
    world, rank = dist.get_world_size(), dist.get_rank()
    per = -(-nslots // world)
    stride = out.shape[1]
    rbytes = np.zeros((per, res.dtype.itemsize), np.uint8)
    rbytes[: len(res)] = res.view(np.uint8).reshape(len(res), -1)
    obuf = np.zeros((per, stride), np.uint8); obuf[: len(out)] = out
    tr = torch.from_numpy(rbytes); to = torch.from_numpy(obuf)
    if device is not None:
        tr = tr.to(device); to = to.to(device)
    gr = [torch.empty_like(tr) for _ in range(world)] if rank == root else None
    go = [torch.empty_like(to) for _ in range(world)] if rank == root else None
    dist.gather(tr, gr, dst=root); dist.gather(to, go, dst=root)
    if rank != root:
        return None, None
    res_all = np.concatenate([g.cpu().numpy() for g in gr])[:nslots].copy().view(res.dtype).reshape(-1)
    out_all = np.concatenate([g.cpu().numpy() for g in go])[:nslots]
    return res_all, out_all

def decode_sharded(decode_fn, iq, frame_off, frame_len, dist=None, device=None, root=0):
    """decode_fn(iq, off, len) -> (res, out) for a block of slots; runs it on this rank's block and gathers on root."""
    n = len(frame_off)
    world = dist.get_world_size() if (dist is not None and dist.is_initialized()) else 1
    rank = dist.get_rank() if world > 1 else 0
    lo, hi = shard_range(n, world, rank)
    res, out = decode_fn(iq, np.asarray(frame_off)[lo:hi], np.asarray(frame_len)[lo:hi])
    return gather_results(res, out, n, dist, device, root)
