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
    flat_g = torch.cat([torch.full((3,), float(rank + 1)), torch.full((4,), 10.0 * (rank + 1))])
    work = parallel.allreduce_flat(flat_g, async_op=True)       # the flat gradient buffer of the fused step, summed in place
    work.wait()
    # BatchNorm running statistics of a batch split over the ranks: weighted partial sums, all-reduced
    total = 6
    mine = range(rank * 3, rank * 3 + 3)
    part = torch.tensor([sum(parallel.ema_pass_weights(j, total) * float(j + 1) for j in mine)], dtype=torch.float64)
    dist.all_reduce(part)
    results[rank] = (best, idx, flat_g[0].item(), flat_g[3].item(), part.item())
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
    serial = 0.0                       # six serial EMA updates with batch statistics 1..6, starting from 0
    for j in range(6):
        serial = 0.9 * serial + 0.1 * (j + 1)
    for rank in (0, 1):
        best, idx, g0, g1, ema = results[rank]
        assert best == table.max() and idx == int(np.argmax(table))
        assert g0 == 3.0 and g1 == 30.0
        assert abs(ema - serial) < 1e-12


def test_decision_plan_covers_every_sample_once():
    from smg_b200 import parallel
    for K in (1, 2, 4, 10):
        for world in (1, 2, 3, 8):
            plan, cap = parallel.decision_plan(K, (16, 16), world, is_ets=True)
            assert (2 in plan) == (K > 1)
            total = 0
            for style, (n_rot, n_masks, shares) in plan.items():
                assert len(shares) == world
                covered = [i for lo, hi in shares for i in range(lo, hi)]
                assert covered == list(range(n_rot + n_masks))
                total += n_rot + n_masks
            assert total == 2 * (16 + K) + ((1 + K * (K - 1) // 2) if K > 1 else 0)
            assert cap * world >= total
