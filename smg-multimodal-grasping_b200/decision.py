"""One decision step of the reference's main loop, de-duplicated.

Reference: code/main.py:137-233.  For K objects the reference calls `trainer.forward` 2K times (enveloping and
sucking, every call re-running the R rotated-scene passes) plus K(K-1)/2 times for enveloping-then-sucking pairs:
2*K*R*2 + K(K-1) trunk passes.  The distinct work is (R + K) passes per single-object primitive and
1 + K(K-1)/2 for the pairs; `decide` evaluates exactly that with three `Trainer.forward_all` calls, reduces the
three Q tables with the device argmax (first max wins, like np.argmax) and applies the reference's primitive
selection rules on the three winning scalars.
"""
import numpy as np
import torch


def scene_from_masks(depth_heightmap, masks):
    """valid_depth_heightmap_a = depth x sum(masks) (code/main.py:145-151)."""
    return np.asarray(depth_heightmap, np.float64) * np.asarray(masks, np.float64).sum(0)


def decide(trainer, depth_heightmap, masks, is_ets=True, is_target=False):
    """Returns dict(primitive, gra_conf [K,R], suc_conf [K,R], gs_conf [K,K], bestg_id, bests_id, bestgs_num,
    bestgs_g_id, bestgs_s_id, bestg_conf, bests_conf, bestgs_conf) with the reference's conventions."""
    masks = np.asarray(masks, np.float64)
    K = masks.shape[0]
    scene = scene_from_masks(depth_heightmap, masks)
    model = trainer.model_target if (is_target and trainer.method == "reinforcement") else trainer.model
    eng = model._engine(max(model.gnum_rotations, model.snum_rotations) + max(K, K * (K - 1) // 2))
    # the scene and the K object masks cross PCIe once; the masked heightmaps are formed on the device
    scene_t = torch.from_numpy(np.ascontiguousarray(scene)).to(eng.device)
    masks_t = torch.from_numpy(np.ascontiguousarray(masks)).to(eng.device)
    single = scene_t[None] * masks_t                                 # code/main.py:160

    def table(style, m):
        q = trainer.forward_all(scene_t, m, style=style, is_target=is_target)     # [n, R, n_out] on the device
        if trainer.method == "reactive":
            # the reference's reactive forward only looks at rotation 0 -> softmax P(class 0) (code/trainer.py:195-199)
            return torch.softmax(q[:, :1, :], dim=2)[:, :, 0]
        return q[:, :, 0]

    # all three passes are enqueued before the first device-to-host read, so the GPU never waits for the host in between
    gra = table(0, single)
    suc = table(1, single)
    vg, ig = eng.argmax(gra)
    vs, isx = eng.argmax(suc)
    pairs, flat, v, i = [], None, None, None
    if is_ets and K > 1:
        pairs = [(g, s) for g in range(K) for s in range(g + 1, K)]
        gi = torch.from_numpy(np.asarray([g for g, _ in pairs], np.int64)).to(eng.device)
        si = torch.from_numpy(np.asarray([s_ for _, s_ in pairs], np.int64)).to(eng.device)
        pm = scene_t[None] * (masks_t[gi] + masks_t[si])                            # code/main.py:186
        gs = table(2, pm)[:, 0]
        flat = torch.full((K * K,), -100.0, device=gs.device)
        idx = torch.from_numpy(np.asarray([g * K + s for g, s in pairs], np.int64)).to(gs.device, non_blocking=True)
        flat[idx] = gs.float()
        v, i = eng.argmax(flat)
    out = {"gra_conf": gra.double().cpu().numpy(), "suc_conf": suc.double().cpu().numpy()}
    bestg_conf, bests_conf = float(vg.item()), float(vs.item())
    Rg, Rs = gra.shape[1], suc.shape[1]
    out["bestg_id"] = (int(ig.item()) // Rg, int(ig.item()) % Rg)      # np.unravel_index(np.argmax(...)) (main.py:172-173)
    out["bests_id"] = (int(isx.item()) // Rs, int(isx.item()) % Rs)
    bestgs_conf, bestgs_num, bestgs_g_id, bestgs_s_id = 0.0, (), [], []
    gs_conf = np.zeros((K, K))
    if is_ets and K > 1:
        bestgs_conf = float(v.item())
        bestgs_num = (int(i.item()) // K, int(i.item()) % K)
        gs_conf = flat.view(K, K).double().cpu().numpy()
        gnu_best = out["gra_conf"].max(axis=1)
        gro_best = out["gra_conf"].argmax(axis=1)
        sro_best = out["suc_conf"].argmax(axis=1)
        a, b = bestgs_num
        if gnu_best[a] > gnu_best[b]:                                                # code/main.py:196-201
            bestgs_g_id, bestgs_s_id = [a, int(gro_best[a])], [b, int(sro_best[b])]
        else:
            bestgs_g_id, bestgs_s_id = [b, int(gro_best[b])], [a, int(sro_best[a])]
    primitive = "grasp"
    if not is_ets or K == 1:
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
                "bestgs_conf": bestgs_conf})
    return out
