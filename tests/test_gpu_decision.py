"""End-to-end decision (config 5 shape at reduced K, R): camera depth -> K11 heightmap -> K12 NMS -> de-duplicated
E / S / ES Q tables -> K9 argmax -> primitive choice, against the CPU oracle evaluated call by call the way the
reference's step loop does (code/main.py:158-233)."""
import numpy as np
import pytest
import torch

from conftest import MEAN, STD
from oracle import action as oaction
from oracle import heightmap as ohm
from oracle import nms as onms
from oracle import qnet

pytestmark = pytest.mark.gpu


def test_decision_matches_oracle_step_loop():
    import smg_b200.synth as synth
    from smg_b200 import NMS, decision, utils
    from smg_b200.trainer import Trainer
    K, R = 3, 4
    torch.manual_seed(0)
    tr = Trainer("reinforcement", 0.5, False, None, False, precision="fp32")
    tr.model.gnum_rotations = tr.model.snum_rotations = R
    sd = {k: v.detach().cpu().clone() for k, v in tr.model.state_dict().items()}
    # front of the pipeline: heightmap + NMS (bit / list exact)
    cam = synth.make_camera(5, num_objects=K, cluttered=False)
    _, d224, _, _, _ = utils.get_heightmap(cam["color"], cam["depth"], cam["intrinsics"], cam["pose"], synth.WORKSPACE_LIMITS, 0.002)
    r224, _, _ = ohm.get_heightmap_depth(cam["depth"], cam["intrinsics"], cam["pose"])
    assert np.array_equal(d224, r224)
    boxes, scores = synth.make_boxes(5, 40)
    assert NMS.py_cpu_nms(boxes, scores, 0.4, 224 * 224 / 60, 224 * 224 / 5) == onms.nms(boxes, scores, 0.4, 224 * 224 / 60, 224 * 224 / 5)
    # Q tables
    sc = synth.make_scene(5, num_objects=K, cluttered=False)
    depth, masks = sc["depth"], sc["masks"].astype(np.float64)
    got = decision.decide(tr, depth, masks, is_ets=True)
    scene = decision.scene_from_masks(depth, masks)
    x = qnet.preprocess(scene, MEAN, STD)
    gra, suc, gs = np.zeros((K, R)), np.zeros((K, R)), np.full((K, K), -100.0)
    for k in range(K):                                   # the reference's loop: one call per object and primitive
        m = qnet.preprocess(scene * masks[k], MEAN, STD)
        gra[k] = [float(o) for o in qnet.model_forward(sd, x, m, 0, True, -1, R, R)]
        suc[k] = [float(o) for o in qnet.model_forward(sd, x, m, 1, True, -1, R, R)]
    for g in range(K):
        for s in range(g + 1, K):
            m = qnet.preprocess(scene * (masks[g] + masks[s]), MEAN, STD)
            gs[g, s] = float(qnet.model_forward(sd, x, m, 2, True, -1, R, R)[0])
    ref = oaction.select_action(gra, suc, gs, is_ets=True)
    scale = max(np.abs(gra).max(), np.abs(suc).max(), np.abs(gs[gs > -100]).max())
    assert np.abs(got["gra_conf"] - gra).max() / scale <= 1e-4
    assert np.abs(got["suc_conf"] - suc).max() / scale <= 1e-4
    assert np.abs(got["gs_conf"] - gs).max() / scale <= 1e-4
    for key in ("primitive", "bestg_id", "bests_id", "bestgs_num", "bestgs_g_id", "bestgs_s_id"):
        assert got[key] == ref[key], (key, got[key], ref[key])
