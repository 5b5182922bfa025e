"""Multi-GPU sharding of the Q pass: one process per GPU, torch.distributed.

The path shards over independent units with no data-path collective: rotations of one decision, objects,
primitives or replay samples are independent given per-sample BatchNorm.  The only exchange is the argmax
(code/main.py:170-173,194-195): every rank reduces its local Q table to one (Q, flat index) tuple and the
tuples are all-gathered (16 bytes per rank; NCCL over NVLink on GPUs, gloo in the CPU tests).
Ties resolve to the lowest flat index so the result equals np.argmax over the full table.
"""
import numpy as np
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


def _world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def allreduce_flat(flat, group=None, async_op=False):
    """Sum one flat gradient buffer over the ranks in place (the training exchange: ncclAllReduce over NVLink on GPUs).
    The fused step already keeps the 368 touched gradients as views of ONE flat buffer, so nothing is concatenated or
    copied back.  Returns the work handle when async_op (None on a single rank)."""
    if _world(group)[1] == 1:
        return None
    return dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)


def ema_pass_weights(global_pass_index, total_passes, momentum=0.1):
    """Weight of BatchNorm pass j of `total_passes` serial passes in the final running statistic:
    running_T = (1-m)^T running_0 + sum_j m (1-m)^(T-1-j) batch_j  (SURVEY.md section 8(e))."""
    return momentum * (1.0 - momentum) ** (total_passes - 1 - global_pass_index)


# --------------------------------------------------------------------------------------------------------------------
# one decision, sharded over the ranks (code/main.py:158-233)
# --------------------------------------------------------------------------------------------------------------------
def decision_plan(K, rotations, world, is_ets=True):
    """For each style the flat list of distinct trunk samples (rotation indices first, then mask indices) and the
    contiguous share of every rank.  Returns {style: (n_rot, n_masks, [(lo, hi) per rank])} and the per-rank capacity
    (partials a rank contributes to the gather, summed over the styles, max over ranks)."""
    plan = {}
    styles = [(0, rotations[0], K), (1, rotations[1], K)]
    if is_ets and K > 1:
        styles.append((2, 1, K * (K - 1) // 2))
    cap = 0
    for style, n_rot, n_masks in styles:
        shares = [shard_range(n_rot + n_masks, r, world) for r in range(world)]
        plan[style] = (n_rot, n_masks, shares)
        cap += max(hi - lo for lo, hi in shares)
    return plan, cap


def decide_sharded(trainer, depth_heightmap, masks, is_ets=True, is_target=False, group=None):
    """`decision.decide` with the distinct trunk passes of ONE decision spread over the ranks.

    Every rank runs its contiguous share of each primitive's samples (rotated scenes and object-masked scenes alike:
    16 + K for enveloping and sucking, 1 + K(K-1)/2 for the pairs) through the trunk and the per-sample half of the head's
    first convolution (`smg_qpartials`); ONE all-gather exchanges the partial products (100 KB per sample, 9.8 MB for K = 10,
    R = 16); every rank then pairs them into the three Q tables (`smg_qcombine`, microseconds) and takes the argmax.  The
    per-rank best (Q, flat index) tuples are exchanged as well (`gather_best`) and must agree with the local argmax.
    Results equal `decide` (per-sample BatchNorm makes every pass independent of its batch).  The BatchNorm
    running-statistics side effect is not reproduced on this path."""
    from . import decision as _decision
    from . import engine as _engine
    rank, world = _world(group)
    ctx = decision_context(trainer, depth_heightmap, masks, world, is_ets, is_target)
    send = decision_local_partials(trainer, ctx, rank)
    if world > 1:
        recv = torch.empty((world * ctx["cap"], 400, 64), dtype=torch.float32, device=send.device)
        dist.all_gather_into_tensor(recv, send, group=group)
    else:
        recv = send
    return decision_from_partials(trainer, ctx, recv, rank, group)


def decision_context(trainer, depth_heightmap, masks, world, is_ets=True, is_target=False):
    """Everything the ranks agree on before the passes: scene, masks, pair list, per-style sample shares."""
    from . import decision as _decision
    masks = np.asarray(masks, np.float64)
    K = masks.shape[0]
    model = trainer.model_target if (is_target and trainer.method == "reinforcement") else trainer.model
    R = {0: model.gnum_rotations, 1: model.snum_rotations, 2: 1}
    plan, cap = decision_plan(K, (R[0], R[1]), world, is_ets)
    return {"scene": _decision.scene_from_masks(depth_heightmap, masks), "masks": masks, "K": K, "model": model, "R": R,
            "pairs": [(g, s) for g in range(K) for s in range(g + 1, K)], "plan": plan, "cap": cap, "world": world}


def decision_local_partials(trainer, ctx, rank):
    """This rank's share of every primitive's samples through the trunk + head half: the all-gather's send buffer
    [cap, 400, 64] (zero-padded where the shares are uneven)."""
    import os
    from . import engine as _engine
    model, plan, cap, scene, masks, pairs = ctx["model"], ctx["plan"], ctx["cap"], ctx["scene"], ctx["masks"], ctx["pairs"]
    max_local = max(hi - lo for _, _, shares in plan.values() for lo, hi in shares)
    eng = model._engine(max(max_local, 18))
    dev = eng.device
    # The three primitives use three different trunks, and a rank's share of one primitive is a handful of samples whose
    # kernel chain cannot fill the GPU: each primitive gets its own handle and stream, so the three chains interleave
    # (SMG_DECISION_STREAMS=1 runs them one after the other on the model's own handle).
    lanes = os.environ.get("SMG_DECISION_STREAMS", "3") != "1"
    main = torch.cuda.current_stream(dev)
    # the scene and the K object masks cross PCIe once; every masked heightmap (code/main.py:160,186) is formed on the device
    scene_t = torch.from_numpy(np.ascontiguousarray(scene)).to(dev, non_blocking=True)
    masks_t = torch.from_numpy(np.ascontiguousarray(masks)).to(dev, non_blocking=True)
    send = torch.zeros((cap, 400, 64), dtype=torch.float32, device=dev)
    off = 0
    used = []
    for style, (n_rot, n_masks, shares) in plan.items():
        lo, hi = shares[rank]
        width = max(h_ - l_ for l_, h_ in shares)
        if hi > lo:
            if lanes:
                e = _engine.get_engine(dev, max(max_local, 18), 640, model.precision, owner=("decision", int(style), id(model)))
                stream = _decision_stream(dev, int(style))
                if not getattr(model, "_smg_decision_finalizer", False):
                    import weakref
                    for s_ in (0, 1, 2):
                        weakref.finalize(model, _engine.drop_engine, ("decision", s_, id(model)))
                    object.__setattr__(model, "_smg_decision_finalizer", True)
            else:
                e, stream = eng, main
            stream.wait_stream(main)
            used.append(stream)
            with torch.cuda.stream(stream):
                e.sync_weights(model, style=style)
                rots = [i for i in range(lo, min(hi, n_rot))]
                mids = [i - n_rot for i in range(max(lo, n_rot), hi)]
                mh_t = None
                if mids and style == 2:
                    gi = torch.tensor([pairs[i][0] for i in mids], device=dev)
                    si = torch.tensor([pairs[i][1] for i in mids], device=dev)
                    mh_t = (scene_t[None] * (masks_t[gi] + masks_t[si])).contiguous()
                elif mids:
                    mh_t = (scene_t[None] * masks_t[mids[0]:mids[-1] + 1]).contiguous()
                nrot_div = model.gnum_rotations if style != 1 else model.snum_rotations
                send[off:off + (hi - lo)] = e.qpartials(style, scene_t, rots, nrot_div, mh_t, trainer.image_mean, trainer.image_std)
        off += width
    for stream in used:
        main.wait_stream(stream)
    return send


_DECISION_STREAMS = {}


def _decision_stream(dev, style):
    key = (str(dev), int(style))
    if key not in _DECISION_STREAMS:
        _DECISION_STREAMS[key] = torch.cuda.Stream(device=dev)
    return _DECISION_STREAMS[key]


def decision_from_partials(trainer, ctx, recv, rank=0, group=None):
    """Pair the gathered partials into the three Q tables and apply the reference's selection rules (every rank)."""
    model, plan, cap, K, R, pairs, world = ctx["model"], ctx["plan"], ctx["cap"], ctx["K"], ctx["R"], ctx["pairs"], ctx["world"]
    eng = model._engine(18)
    dev = eng.device
    offsets, off = {}, 0
    for style, (n_rot, n_masks, shares) in plan.items():
        offsets[style] = off
        off += max(h_ - l_ for l_, h_ in shares)
    out = {}
    tables = {}
    for style, (n_rot, n_masks, shares) in plan.items():
        o = offsets[style]
        parts = torch.cat([recv[r * cap + o: r * cap + o + (hi - lo)] for r, (lo, hi) in enumerate(shares) if hi > lo])
        eng.sync_weights(model, style=style)
        q = eng.qcombine(style, parts[:n_rot], parts[n_rot:])                 # [n_masks, n_rot, n_out]
        if trainer.method == "reactive":
            tables[style] = torch.softmax(q[:, :1, :], dim=2)[:, :, 0]
        else:
            tables[style] = q[:, :, 0]
    gra, suc = tables[0], tables[1]
    out["gra_conf"], out["suc_conf"] = gra.double().cpu().numpy(), suc.double().cpu().numpy()
    vg, ig = eng.argmax(gra)
    vs, isx = eng.argmax(suc)
    # the north-star exchange: every rank reduces ITS rows of the table to one (Q, flat index) tuple, the tuples are gathered
    if rank is not None:
        lo, hi = shard_range(K, rank, world)
        flat_idx = torch.arange(lo * R[0], hi * R[0], device=dev, dtype=torch.int64)
        bq, bi = gather_best(gra[lo:hi].reshape(-1), flat_idx, group)
        assert bi == int(ig.item()), "sharded argmax disagrees with the argmax of the gathered table"
    bestg_conf, bests_conf = float(vg.item()), float(vs.item())
    out["bestg_id"] = (int(ig.item()) // R[0], int(ig.item()) % R[0])
    out["bests_id"] = (int(isx.item()) // R[1], int(isx.item()) % R[1])
    bestgs_conf, bestgs_num, bestgs_g_id, bestgs_s_id = 0.0, (), [], []
    gs_conf = np.zeros((K, K))
    if 2 in tables:
        flat = torch.full((K * K,), -100.0, device=dev)
        idx = torch.tensor([g * K + s for g, s in pairs], device=dev)
        flat[idx] = tables[2][:, 0].float()
        v, i = eng.argmax(flat)
        bestgs_conf = float(v.item())
        bestgs_num = (int(i.item()) // K, int(i.item()) % K)
        gs_conf = flat.view(K, K).double().cpu().numpy()
        gnu_best, gro_best = out["gra_conf"].max(axis=1), out["gra_conf"].argmax(axis=1)
        sro_best = out["suc_conf"].argmax(axis=1)
        a, b = bestgs_num
        if gnu_best[a] > gnu_best[b]:                                              # code/main.py:196-201
            bestgs_g_id, bestgs_s_id = [a, int(gro_best[a])], [b, int(sro_best[b])]
        else:
            bestgs_g_id, bestgs_s_id = [b, int(gro_best[b])], [a, int(sro_best[a])]
    primitive = "grasp"
    if 2 not in tables:
        if bests_conf > bestg_conf:
            primitive = "suction"
    else:
        g2 = 2 * bestgs_conf if trainer.method == "reactive" else bestgs_conf      # code/main.py:221-233
        if bests_conf > max(bestg_conf, g2):
            primitive = "suction"
        elif g2 > max(bests_conf, bestg_conf):
            primitive = "grasp_then_suction"
    out.update({"primitive": primitive, "gs_conf": gs_conf, "bestgs_num": bestgs_num, "bestgs_g_id": bestgs_g_id,
                "bestgs_s_id": bestgs_s_id, "bestg_conf": bestg_conf, "bests_conf": bests_conf,
                "bestgs_conf": bestgs_conf, "exchange_bytes": int(cap * 400 * 64 * 4 * world)})
    return out
