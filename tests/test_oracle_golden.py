"""The CPU oracle against the golden fixtures produced by the unmodified reference (CPU only)."""
import hashlib
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, MEAN, STD, check_fingerprint
from oracle import action as oaction
from oracle import heightmap as ohm
from oracle import nms as onms
from oracle import qnet


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_versions_match_fixture(golden):
    import cv2, torchvision
    v = golden["versions"]
    assert v["torch"] == torch.__version__ and v["torchvision"] == torchvision.__version__
    assert v["numpy"] == np.__version__ and v["cv2"] == cv2.__version__


def test_same_seed_weights(golden, rl_state_dict):
    g = golden["weights"]
    assert len(rl_state_dict) == g["n_keys"] == 2217
    assert hashlib.sha256("\n".join(rl_state_dict.keys()).encode()).hexdigest() == g["keys_sha"]
    assert sha(rl_state_dict["grasp_depth_trunk.features.conv0.weight"].numpy()) == g["conv0_sha"]
    assert sha(rl_state_dict["graspnet_val.grasp-val-conv1.weight"].numpy()) == g["grasp_head_conv1_sha"]


def test_inputs_match_fixture(golden, scene_inputs):
    scene, mask, pair, _ = scene_inputs
    assert sha(scene) == golden["inputs"]["scene_sha"]
    assert sha(mask) == golden["inputs"]["mask_sha"]
    assert sha(pair) == golden["inputs"]["pair_sha"]


def test_preprocess_is_zoom_pad_normalise():
    from scipy import ndimage
    rs = np.random.RandomState(0)
    d = rs.uniform(0, 0.1, size=(224, 224))
    z = ndimage.zoom(d, zoom=[2, 2], order=0)          # code/trainer.py:165
    z = np.pad(z, 96, "constant", constant_values=0)   # code/trainer.py:169-173
    ref = ((z - MEAN) / STD).astype(np.float32)
    x = qnet.preprocess(d, MEAN, STD)
    assert x.shape == (1, 3, 640, 640)
    for c in range(3):
        assert np.array_equal(x[0, c].numpy(), ref)


@pytest.mark.parametrize("H,R", [(64, 16), (640, 16), (640, 1), (96, 4)])
def test_rotate_index_map_matches_torch(H, R):
    src = torch.arange(H * H, dtype=torch.float32).reshape(1, 1, H, H) + 1
    for r in range(R):
        ref = qnet.rotate_nearest(src, r, R)[0, 0].numpy().astype(np.int64) - 1
        assert np.array_equal(qnet.rotate_index_map(H, r, R).astype(np.int64), ref), (H, R, r)


def test_trunk_taps(golden, rl_state_dict, scene_inputs):
    x = qnet.preprocess(scene_inputs[0], MEAN, STD)
    taps = {}
    with torch.no_grad():
        f = qnet.densenet_features(rl_state_dict, "grasp_depth_trunk.features.", x, taps)
    g = golden["trunk_taps"]
    check_fingerprint(taps["conv0"], g["conv0"], 1e-6)
    check_fingerprint(taps["pool0"], g["pool0"], 1e-6)
    for b in (1, 2, 3, 4):
        check_fingerprint(taps["block%d" % b], g["denseblock%d" % b], 1e-5)
    for b in (1, 2, 3):
        check_fingerprint(taps["trans%d" % b], g["transition%d" % b], 1e-5)
    check_fingerprint(f, g["norm5"], 1e-5)


def test_q_values_R1(golden, rl_state_dict, scene_inputs):
    scene, mask, pair, _ = scene_inputs
    x, m, m2 = (qnet.preprocess(v, MEAN, STD) for v in (scene, mask, pair))
    for style in (0, 1, 2):
        out = qnet.model_forward(rl_state_dict, x, m2 if style == 2 else m, style, True, -1)
        assert isinstance(out, list) and len(out) == 1 and tuple(out[0].shape) == (1, 1, 1, 1)
        assert abs(float(out[0]) - golden["q"]["rl_style%d_R1" % style][0]) < 2e-5


def test_q_values_R16(golden, rl_state_dict, scene_inputs):
    scene, mask, pair, _ = scene_inputs
    x, m, m2 = (qnet.preprocess(v, MEAN, STD) for v in (scene, mask, pair))
    out = qnet.model_forward(rl_state_dict, x, m, 0, True, -1, gnum_rotations=16, snum_rotations=16)
    got = np.array([float(o) for o in out])
    ref = np.array(golden["q"]["rl_style0_R16"])
    assert got.shape == (16,) and np.abs(got - ref).max() < 2e-5
    q1 = qnet.model_forward(rl_state_dict, x, m, 1, True, 5, gnum_rotations=16, snum_rotations=16)
    assert abs(float(q1) - golden["q"]["rl_style1_R16_rot5"][0]) < 2e-5
    q2 = qnet.model_forward(rl_state_dict, x, m2, 2, True, 5, gnum_rotations=16, snum_rotations=16)
    assert abs(float(q2) - golden["q"]["rl_style2_R16_rot5"][0]) < 2e-5  # ES ignores the rotation


def test_reactive_logits(golden, scene_inputs):
    import smg_b200.models as models
    torch.manual_seed(0)
    sd = models.reactive_net(True).state_dict()
    scene, mask, pair, _ = scene_inputs
    x, m, m2 = (qnet.preprocess(v, MEAN, STD) for v in (scene, mask, pair))
    out = qnet.model_forward(sd, x, m, 0, True, -1)[0].view(-1)
    assert np.abs(out.numpy() - np.array(golden["reactive_style0_R1"])).max() < 2e-5
    out = qnet.model_forward(sd, x, m2, 2, True, -1)[0].view(-1)
    assert np.abs(out.numpy() - np.array(golden["reactive_style2_R1"])).max() < 2e-5


def test_backprop_rl(golden, rl_state_dict, scene_inputs):
    scene, mask, _, _ = scene_inputs
    x, m = qnet.preprocess(scene, MEAN, STD), qnet.preprocess(mask, MEAN, STD)
    g = golden["backprop_rl_grasp"]
    loss, grads = qnet.backprop_grads(rl_state_dict, x, m, 0, 0, g["label"], "reinforcement")
    assert abs(loss - g["loss"]) < 1e-5
    assert len(grads) == g["n_grads"] == 368
    for k, fp in g["grads"].items():
        check_fingerprint(grads[k], fp, 2e-3)
    # Adam first step: delta = -lr * g / (|g| + eps)  (code/trainer.py:99)
    k = "graspnet_val.grasp-val-conv1.weight"
    p = rl_state_dict[k]
    newp, _, _ = qnet.adam_step(p, grads[k], torch.zeros_like(p), torch.zeros_like(p), 1)
    check_fingerprint(newp - p, g["param_delta"][k], 2e-3)


def test_backprop_reactive(golden, scene_inputs):
    import smg_b200.models as models
    torch.manual_seed(0)
    sd = models.reactive_net(True).state_dict()
    scene, mask, _, _ = scene_inputs
    x, m = qnet.preprocess(scene, MEAN, STD), qnet.preprocess(mask, MEAN, STD)
    g = golden["backprop_reactive_suction"]
    loss, grads = qnet.backprop_grads(sd, x, m, 1, 0, g["label"], "reactive")
    assert abs(loss - g["loss"]) < 1e-5 and len(grads) == g["n_grads"]
    for k, fp in g["grads"].items():
        check_fingerprint(grads[k], fp, 2e-3)


def test_heightmap_bit_exact(golden):
    import smg_b200.synth as synth
    cam = synth.make_camera(golden["heightmap"]["camera_seed"])
    assert sha(cam["depth"]) == golden["heightmap"]["depth_sha"]
    d224, d448, A = ohm.get_heightmap_depth(cam["depth"], cam["intrinsics"], cam["pose"])
    z = np.load(os.path.join(GOLDEN_DIR, "heightmap_seed3.npz"))
    assert np.array_equal(d224, z["depth224"])
    assert np.array_equal(d448[::7], z["depth448_rows"])
    assert np.array_equal(A, z["A_htor"])
    assert sha(d224) == golden["heightmap"]["depth224_sha"]
    assert sha(d448) == golden["heightmap"]["depth448_sha"]
    # colour outputs: cv2's 15-bit fixed-point 8-bit remap, pinned on the reference's own output of the same camera
    c224, c448 = ohm.get_heightmap_color(cam["color"])
    assert sha(c224) == golden["heightmap"]["color224_sha"]
    assert sha(c448) == golden["heightmap"]["color448_sha"]


def test_nms_cases(golden):
    import smg_b200.synth as synth
    known = np.array([[[10, 10], [60, 60]], [[12, 12], [62, 62]], [[100, 100], [160, 150]], [[0, 0], [5, 5]],
                      [[0, 0], [200, 200]]], np.float32)
    for case in golden["nms"]:
        if case["kind"] == "known5":
            keep = onms.nms(known, np.ones(5), 0.40, 224 * 224 / 60, 224 * 224 / 5)
            assert keep == case["keep"] == [0, 2]
        else:
            n = case["n"]
            boxes, scores = synth.make_boxes(case["seed"], n) if n else (np.zeros((0, 2, 2), np.float32), np.zeros(0))
            assert onms.nms(boxes, scores, 0.40, 224 * 224 / 60, 224 * 224 / 5) == case["keep"]


def test_action_selection_rules():
    gra = np.array([[0.1, 0.5], [0.5, 0.2]])   # tie: first max wins -> (0,1)
    suc = np.array([[0.3, 0.1], [0.2, 0.6]])
    a = oaction.select_action(gra, suc)
    assert a["bestg_id"] == (0, 1) and a["bests_id"] == (1, 1) and a["primitive"] == "suction"
    gs = np.full((2, 2), -100.0)
    gs[0, 1] = 0.9
    a = oaction.select_action(gra, suc, gs, is_ets=True)
    assert a["primitive"] == "grasp_then_suction" and a["bestgs_num"] == (0, 1)
    # gnu_best equal -> else branch: second object grasps
    assert a["bestgs_g_id"][0] == 1 and a["bestgs_s_id"][0] == 0
    a = oaction.select_action(gra, suc, gs * 0 + 0.25, is_ets=True, method="reactive")
    assert a["primitive"] == "suction"  # 2*0.25 = 0.5 < 0.6


# ---------------------------------------------------------------------------------------------------------------
# round 2 fixtures (tests/golden/golden_r02.json, made by tests/golden/make_golden_r02.py from the unmodified reference)
# ---------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def gold2():
    import json
    import os
    from conftest import GOLDEN_DIR
    with open(os.path.join(GOLDEN_DIR, "golden_r02.json")) as f:
        return json.load(f)


def test_oracle_highly_cluttered_k10_entries(gold2, rl_state_dict):
    """The oracle on the K = 10 highly-cluttered scene: one object's grasp row (R = 4) and one ES pair."""
    import smg_b200.synth as synth
    g = gold2["hc"]
    sc = synth.make_scene(g["scene_seed"], num_objects=g["K"], cluttered=True)
    scene, masks = sc["scene"], sc["masks"].astype(np.float64)
    x = qnet.preprocess(scene, MEAN, STD)
    with torch.no_grad():
        row = qnet.q_forward(rl_state_dict, x, qnet.preprocess(scene * masks[1], MEAN, STD), 0, range(g["R"]), g["R"])
        pair = qnet.model_forward(rl_state_dict, x, qnet.preprocess(scene * (masks[0] + masks[9]), MEAN, STD), 2, True, -1,
                                  gnum_rotations=g["R"], snum_rotations=g["R"])
    assert np.abs(np.asarray([float(o.view(-1)[0]) for o in row]) - np.asarray(g["gra_conf"][1])).max() <= 2e-5
    assert abs(float(pair[0].view(-1)[0]) - g["gs_conf"][0][9]) <= 2e-5


def test_oracle_label_value_cases(gold2, rl_state_dict, scene_inputs):
    from oracle import labels
    g = gold2["label_value"]
    scene, _, _, sc = scene_inputs
    for c in g["cases"]:
        lv, rv = labels.label_value(c["method"], rl_state_dict, c["args"], scene, sc["masks"].astype(np.float64),
                                    num_rotations=g["num_rotations"], mean=MEAN, std=STD)
        assert float(rv) == c["reward_value"]
        assert abs(float(lv) - c["label_value"]) <= 2e-5, c


def test_preload_restores_the_ten_logs(gold2):
    """Trainer.preload (code/trainer.py:118-158) on log files written by the reference's logger: same iteration, same lists."""
    import os
    import types
    from conftest import GOLDEN_DIR
    from smg_b200.trainer import Trainer
    g = gold2["preload"]
    t = types.SimpleNamespace(_LOGS=Trainer._LOGS)
    Trainer.preload(t, os.path.join(GOLDEN_DIR, g["dir"]))
    assert t.iteration == g["iteration"]
    for attr, want in g["logs"].items():
        got = getattr(t, attr)
        assert isinstance(got, list) and got == want, attr
    t.executed_action_log.append([1, 2, 3, 4])      # main.py:369 appends to the restored list


def test_oracle_geometry_vs_reference(gold2):
    """oracle/geometry.py (the step-by-step restatement csrc/geometry.cu follows) against the reference's own
    global_position / get_best_grasp_angle / get_best_suction_angle outputs."""
    import smg_b200.synth as synth
    from oracle import geometry as ogeo
    n_oo = 0
    for e in gold2["geometry"]:
        cam = synth.make_camera(e["camera_seed"])
        A, K, P, depth = np.asarray(e["A_htor"]), cam["intrinsics"], cam["pose"], cam["depth"]
        box, cter = np.asarray(e["box_mask_cors"]), np.asarray(e["masks_cter"])
        for gp in e["global_position"]:
            assert np.allclose(ogeo.global_position(gp["pix"], A, K, P, depth), gp["xyz"], rtol=0, atol=1e-12)
        for gr in e["grasp"]:
            c, ang, dist = ogeo.grasp_angle(gr["is_pe"], box, gr["obj"], A, K, P, depth)
            assert np.allclose(c, gr["center"], atol=1e-12) and abs(ang - gr["angle"]) <= 1e-12 and abs(dist - gr["open_distance"]) <= 1e-12
        for su in e["suction"]:
            c, ang = ogeo.suction_angle(su["is_oo"], e["K"], cter, box, su["obj"], A, K, P, depth)
            assert np.allclose(c, su["center"], atol=1e-12)
            assert abs(float(ang) - su["angle"]) <= 1e-12, (e["scene_seed"], su, float(ang))
            n_oo += su["is_oo"]
    assert n_oo == 20


def test_snapshot_round_trip_with_reference_logger(tmp_path):
    """N4: a `.pth` written by the reference's own logger (code/logger.py:121-125) from the reference's own
    net loads into this package's net unchanged, and a snapshot of ours loads back into the reference net (2217 keys, same
    shapes and values).  Needs /root/reference: runs in the build container only."""
    from oracle import refshim
    if not os.path.isdir("/root/reference/code"):
        pytest.skip("the reference tree is only present in the build container")
    mods = refshim.install(cpu=True)
    try:
        import logger as ref_logger
        import smg_b200.models as mine
        torch.manual_seed(3)
        ref = mods["models"].reinforcement_net(True)

        class L:
            models_directory = str(tmp_path)
        ref_logger.Logger.save_model(L, 50, ref, "reinforcement")                     # model.cpu().state_dict()
        ref_logger.Logger.save_backup_model(L, ref, "reinforcement")
        snap = os.path.join(str(tmp_path), "snapshot-000050.reinforcement.pth")
        net = mine.reinforcement_net(True)
        missing = net.load_state_dict(torch.load(snap))                                 # code/trainer.py:85-87
        assert not missing.missing_keys and not missing.unexpected_keys
        sd_ref, sd_mine = ref.state_dict(), net.state_dict()
        assert list(sd_ref.keys()) == list(sd_mine.keys()) and len(sd_mine) == 2217
        assert all(torch.equal(sd_ref[k], sd_mine[k]) for k in sd_ref)
        torch.save(net.state_dict(), os.path.join(str(tmp_path), "mine.pth"))
        ref2 = mods["models"].reinforcement_net(True)
        ref2.load_state_dict(torch.load(os.path.join(str(tmp_path), "mine.pth")))
        assert all(torch.equal(v, ref2.state_dict()[k]) for k, v in sd_mine.items())
        assert os.path.exists(os.path.join(str(tmp_path), "snapshot-backup.reinforcement.pth"))
    finally:
        refshim.set_cpu_mode(False)
