"""Multi-GPU plumbing for the tile path: one process per GPU, independent shards, no hot-path collective.

The reference fans images out to GPUs from one host queue (src/main.cpp:778-828); here every rank owns a static,
balanced shard of the work list and the only communication is the one-time weight broadcast at init."""
import os
import tempfile

import numpy as np


def shard_items(n_items, world, rank):
    """Indices of the work items (images or tiles) rank `rank` processes: contiguous, sizes differ by at most one."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


def broadcast_model_files(parampath, binpath, rank, world, device="cpu"):
    """Rank 0 reads x4.param / x4.bin; every other rank receives the bytes with torch.distributed.broadcast (NCCL on
    GPUs, gloo in the CPU tests) and stores them in a private temp dir.  Returns (parampath, binpath, keepalive)."""
    if world == 1:
        return parampath, binpath, None
    import torch
    import torch.distributed as dist
    blobs = []
    for path in (parampath, binpath):
        n = torch.tensor([os.path.getsize(path) if rank == 0 else 0], dtype=torch.int64, device=device)
        dist.broadcast(n, 0)
        t = torch.empty(int(n.item()), dtype=torch.uint8, device=device)
        if rank == 0:
            t.copy_(torch.from_numpy(np.fromfile(path, dtype=np.uint8)))
        dist.broadcast(t, 0)
        blobs.append(t)
    if rank == 0:
        return parampath, binpath, None
    td = tempfile.TemporaryDirectory()
    p, b = os.path.join(td.name, "x4.param"), os.path.join(td.name, "x4.bin")
    blobs[0].cpu().numpy().tofile(p)
    blobs[1].cpu().numpy().tofile(b)
    return p, b, td
