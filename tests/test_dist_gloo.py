"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: rotation sharding + best-tuple exchange."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, table, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from smg_b200 import parallel
    n_obj, R = table.shape
    rots = parallel.shard_rotations(R, rank, world)
    q_local = torch.from_numpy(table[:, rots].copy()).reshape(-1)
    flat = torch.tensor([o * R + r for o in range(n_obj) for r in rots], dtype=torch.int64)
    best, idx = parallel.gather_best(q_local, flat)
    g = [torch.full((3,), float(rank + 1)), torch.full((2, 2), 10.0 * (rank + 1))]
    parallel.allreduce_grads(g)
    results[rank] = (best, idx, g[0][0].item(), g[1][0, 0].item())
    dist.destroy_process_group()


def test_rotation_shards_cover_everything():
    from smg_b200 import parallel
    for R in (1, 3, 16):
        for world in (1, 2, 4, 8):
            got = sum((parallel.shard_rotations(R, r, world) for r in range(world)), [])
            assert got == list(range(R))


def test_gather_best_equals_numpy_argmax_world2():
    rs = np.random.RandomState(0)
    table = rs.randn(5, 16)
    table[3, 9] = table.max()      # plant a tie across ranks: np.argmax picks the lowest flat index
    table[1, 2] = table.max()
    port = _free_port()
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(2, port, table, results), nprocs=2, join=True)
    for rank in (0, 1):
        best, idx, g0, g1 = results[rank]
        assert best == table.max() and idx == int(np.argmax(table))
        assert g0 == 3.0 and g1 == 30.0
