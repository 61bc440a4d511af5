"""CPU test of the N>1 host logic with world_size 2 over gloo: weight broadcast from rank 0 and work sharding."""
import hashlib
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from common import model_paths


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, param, binf, q):
    import realsr_ncnn_vulkan_b200.dist as rd
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # rank 1 must not read the originals: give it bogus paths
    p, b, keep = rd.broadcast_model_files(param if rank == 0 else "/nonexistent.param", binf if rank == 0 else "/nonexistent.bin", rank, world)
    h = hashlib.sha256(open(p, "rb").read() + open(b, "rb").read()).hexdigest()
    mine = rd.shard_items(100, world, rank)
    t = torch.tensor([float(len(mine))])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)  # the bench's max-over-ranks reduction
    q.put((rank, h, mine, float(t[0])))
    dist.barrier()
    dist.destroy_process_group()


def test_weight_broadcast_and_sharding_world2(models_dir):
    param, binf = model_paths("models-DF2K")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, param, binf, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    ref = hashlib.sha256(open(param, "rb").read() + open(binf, "rb").read()).hexdigest()
    assert res[0][1] == ref and res[1][1] == ref
    assert sorted(res[0][2] + res[1][2]) == list(range(100)) and res[0][3] == 50.0


def test_shard_items_balanced():
    import realsr_ncnn_vulkan_b200.dist as rd
    for n, w in ((100, 8), (15, 4), (3, 8), (0, 2)):
        parts = [rd.shard_items(n, w, r) for r in range(w)]
        assert sorted(sum(parts, [])) == list(range(n))
        assert max(map(len, parts)) - min(map(len, parts)) <= 1


def test_split_tiles_partitions_the_grid():
    """tile-granular sharding of one image (sharded.split_tiles): contiguous, complete, disjoint, balanced"""
    import realsr_ncnn_vulkan_b200 as rb
    for (w, h, T, n) in [(1920, 1080, 400, 2), (1920, 1080, 400, 8), (3840, 2160, 200, 8), (3840, 2160, 200, 3), (256, 256, 200, 4),
                         (100, 100, 200, 8), (1280, 720, 200, 5), (33, 977, 32, 7)]:
        costs = rb.tile_costs(w, h, T, 10)
        assert len(costs) == -(-w // T) * -(-h // T)
        parts = rb.split_tiles(w, h, T, 10, n)
        assert len(parts) == n
        nxt = 0
        for first, count in parts:
            assert first == nxt and count >= 0
            nxt = first + count
        assert nxt == len(costs)
        loads = [sum(costs[f:f + c]) for f, c in parts]
        assert sum(loads) == sum(costs)
        # no part carries more than its fair share plus one (largest) tile
        assert max(loads) <= sum(costs) / n + max(costs)
    # BASELINE configs[3] over 8 GPUs: 220 tiles, every GPU within 4 % of 1/8 of the work
    costs = rb.tile_costs(3840, 2160, 200, 10)
    loads = [sum(costs[f:f + c]) for f, c in rb.split_tiles(3840, 2160, 200, 10, 8)]
    assert max(loads) / (sum(costs) / 8) < 1.04, loads
