"""The multi-GPU path shards independent worlds with no data-path collective; here the host logic (partition, timing
reductions) runs on 2 CPU ranks over gloo."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from moleengine_b200 import scenes, shard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world_size, port, n_worlds, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world_size), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    lo, hi = shard.world_range(rank, world_size, n_worlds)
    sc = scenes.c5_worlds(n_worlds=hi - lo, bodies_per_side=3, first_world=lo)
    # every rank built only its own worlds, seeded by the GLOBAL world index
    full = scenes.c5_worlds(n_worlds=n_worlds, bodies_per_side=3)
    per = 9
    assert sc["bodies"].tobytes() == full["bodies"][lo * per: hi * per].tobytes()
    n_bodies = shard.reduce_sum(len(sc["bodies"]))
    t = shard.reduce_max(1.0 + rank)
    if rank == 0:
        out.put((n_bodies, t, lo, hi))
    dist.barrier()
    dist.destroy_process_group()


def test_world_range_partitions_exactly():
    for n in (0, 1, 7, 8, 8192, 8193):
        for ws in (1, 2, 3, 8):
            blocks = [shard.world_range(r, ws, n) for r in range(ws)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(ws - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 5, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    n_bodies, t, lo, hi = out.get(timeout=5)
    assert n_bodies == 5 * 9 and t == 2.0 and (lo, hi) == (0, 3)
