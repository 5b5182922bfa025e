"""Multi-GPU sharding of the Q pass: one process per GPU, torch.distributed.

The path shards over independent units with no data-path collective: rotations of one decision, objects,
primitives or replay samples are independent given per-sample BatchNorm.  The only exchange is the argmax
(code/main.py:170-173,194-195): every rank reduces its local Q table to one (Q, flat index) tuple and the
tuples are all-gathered (16 bytes per rank; NCCL over NVLink on GPUs, gloo in the CPU tests).
Ties resolve to the lowest flat index so the result equals np.argmax over the full table.
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous block of `n` items owned by `rank` (sizes differ by at most one, lower ranks first)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_rotations(num_rotations, rank, world):
    """Rotation indices evaluated by `rank` (the masked-scene pass is recomputed by every rank: 1 of 17 passes)."""
    lo, hi = shard_range(num_rotations, rank, world)
    return list(range(lo, hi))


def local_best(q_local, flat_index):
    """(max Q, its global flat index) of this rank's slice; first-max-wins.  q_local [n], flat_index [n] int64."""
    if q_local.numel() == 0:
        return torch.tensor([float("-inf")], dtype=torch.float64, device=q_local.device), \
            torch.tensor([2 ** 62], dtype=torch.int64, device=q_local.device)
    q = q_local.double()
    m = q.max()
    idx = flat_index[q == m].min()
    return m.view(1), idx.view(1)


def gather_best(q_local, flat_index, group=None):
    """Global (max Q, flat index) over all ranks == (np.max(table), np.argmax(table)) of the full table."""
    m, idx = local_best(q_local, flat_index)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(m), int(idx)
    world = dist.get_world_size(group)
    payload = torch.cat([m, idx.double()])            # indices < 2^53 are exact in float64
    out = [torch.empty_like(payload) for _ in range(world)]
    dist.all_gather(out, payload, group=group)
    table = torch.stack(out)                          # [world, 2]
    best = table[:, 0].max()
    cand = table[table[:, 0] == best]
    return float(best), int(cand[:, 1].min())


def allreduce_grads(grads, group=None):
    """Data-parallel training: sum the touched trunk + head gradients over ranks (one flat all-reduce)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return grads
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()
    return grads
