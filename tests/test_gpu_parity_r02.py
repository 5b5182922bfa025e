"""Round-2 GPU parity: exactly what bench.py times, the highly-cluttered K = 10 configuration, `get_label_value`, and a
kink-free end-to-end gradient check (VERDICT r01 "next round" item 1).  All through the C ABI; fixtures in
tests/golden/golden_r02.json were produced by the unmodified reference (tests/golden/make_golden_r02.py)."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, MEAN, STD
from oracle import action as oaction
from oracle import qnet

pytestmark = pytest.mark.gpu

TOL = {"fp32": 1e-4, "tf32": 1e-2}     # north_star: max |dQ| / max_candidates |Q_ref|


@pytest.fixture(scope="module")
def gold2():
    with open(os.path.join(GOLDEN_DIR, "golden_r02.json")) as f:
        return json.load(f)


def _trainer(method="reinforcement", precision="tf32", rotations=1):
    from smg_b200.trainer import Trainer
    torch.manual_seed(0)
    tr = Trainer(method, 0.5, False, None, False, precision=precision)
    nets = [tr.model] + ([tr.model_target] if method == "reinforcement" else [])
    for m in nets:
        m.gnum_rotations = m.snum_rotations = rotations
    return tr


def _near_tie(table, tol):
    """True if the two largest entries of `table` are closer than tol * max|table| (a stated near-tie)."""
    v = np.sort(np.asarray(table, np.float64).ravel())[::-1]
    return len(v) > 1 and (v[0] - v[1]) <= tol * np.abs(v).max()


# --------------------------------------------------------------------------------------------------
# (a) the benched configuration: tf32, Trainer.forward_batch with G = 4, CUDA-graph replay, R = 16
# --------------------------------------------------------------------------------------------------
def test_benched_config_tf32_batch4_graph_replay_vs_reference(scene_inputs, golden, rl_state_dict):
    import smg_b200.synth as synth
    scene, mask0, _, sc = scene_inputs
    tr = _trainer(precision="tf32", rotations=16)
    tr.model.update_running_stats = False                  # as bench.py's timed region
    scenes = np.stack([scene] * 4)
    masks = np.stack([synth.masked_scene(scene, sc["masks"], [k]) for k in range(4)])
    q1 = tr.forward_batch(scenes, masks, 0)                # first sighting: eager
    eng = tr.model._engine(4 * 17, 0)
    n1 = eng.launch_count()
    q2 = tr.forward_batch(scenes, masks, 0)                # second: capture + replay
    q3 = tr.forward_batch(scenes, masks, 0)                # third: pure replay (what the timed loop runs)
    assert eng.launch_count() - n1 > 2 * 100, "the replayed passes must be counted as launches of this library"
    assert q3.shape == (4, 1, 16)
    ref = np.zeros((4, 16))
    ref[0] = golden["q"]["rl_style0_R16"]                  # unit 0: recorded from the reference itself
    x = qnet.preprocess(scene, MEAN, STD)
    with torch.no_grad():
        for k in (1, 2, 3):                                # the other units: the pinned oracle on the host
            m = qnet.preprocess(masks[k], MEAN, STD)
            ref[k] = [float(o.view(-1)[0]) for o in qnet.q_forward(rl_state_dict, x, m, 0, range(16), 16)]
    scale = np.abs(ref).max()
    for name, q in (("eager", q1), ("capture", q2), ("replay", q3)):
        err = np.abs(q[:, 0, :] - ref).max() / scale
        print("benched config (%s): max|dQ|/max|Q| = %.2e" % (name, err))
        assert err <= TOL["tf32"], (name, err)
    for k in range(4):                                     # chosen rotation identical except on stated near-ties
        if not _near_tie(ref[k], 2 * TOL["tf32"]):
            assert int(np.argmax(q3[k, 0])) == int(np.argmax(ref[k])), k
    # the same call with the running-statistics side effect switched on gives the same Q (it only exports statistics)
    q_single = tr.forward(scene, masks[1], 0, True, False)
    assert np.abs(q_single - q3[1, 0]).max() / scale <= 1e-5


def test_repeated_passes_are_stable(scene_inputs):
    """The Q pass has no data-dependent schedule: repeating it must give the same Q every time (the only run-to-run freedom
    is the order of the double-precision statistics atomics, ~1e-7).  Regression test for a barrier-parity aliasing fault of
    the first transition kernel that produced ~1e-2 deviations in about one pass of a hundred, and only inside the network
    (L2-hot inputs), never in the isolated kernel tests."""
    import smg_b200.synth as synth
    scene, mask0, _, sc = scene_inputs
    tr = _trainer(precision="tf32", rotations=16)
    tr.model.update_running_stats = False
    scenes = np.stack([scene] * 4)
    masks = np.stack([synth.masked_scene(scene, sc["masks"], [k]) for k in range(4)])
    q_single = tr.forward(scene, masks[0], 0, True, False)
    q_batch = tr.forward_batch(scenes, masks, 0)
    scale = np.abs(q_batch).max()
    worst = 0.0
    for rep in range(60):
        for k in range(2):
            worst = max(worst, np.abs(tr.forward(scene, masks[0], 0, True, False) - q_single).max() / scale)
        worst = max(worst, np.abs(tr.forward_batch(scenes, masks, 0) - q_batch).max() / scale)
    print("180 repeated passes: worst deviation %.2e of scale" % worst)
    assert worst <= 1e-5


# --------------------------------------------------------------------------------------------------
# (b) highly-cluttered K = 10 scene through forward_all / decide, all three primitives
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_highly_cluttered_k10_tables_vs_reference(gold2, precision):
    import smg_b200.synth as synth
    from smg_b200 import decision
    g = gold2["hc"]
    K, R = g["K"], g["R"]
    sc = synth.make_scene(g["scene_seed"], num_objects=K, cluttered=True)
    tr = _trainer(precision=precision, rotations=R)
    out = decision.decide(tr, sc["depth"], sc["masks"], is_ets=True)
    gra, suc, gs = np.asarray(g["gra_conf"]), np.asarray(g["suc_conf"]), np.asarray(g["gs_conf"])
    scale = max(np.abs(gra).max(), np.abs(suc).max(), np.abs(gs[gs > -99]).max())
    e_g = np.abs(out["gra_conf"] - gra).max() / scale
    e_s = np.abs(out["suc_conf"] - suc).max() / scale
    e_p = np.abs(out["gs_conf"] - gs).max() / scale
    print("hc K=10 R=%d %s: grasp %.2e suction %.2e pairs %.2e (of scale %.3f)" % (R, precision, e_g, e_s, e_p, scale))
    assert max(e_g, e_s, e_p) <= TOL[precision]
    ref = oaction.select_action(gra, suc, gs, is_ets=True)
    tol = 2 * TOL[precision]
    if not _near_tie(gra, tol):
        assert tuple(out["bestg_id"]) == ref["bestg_id"]
    if not _near_tie(suc, tol):
        assert tuple(out["bests_id"]) == ref["bests_id"]
    if not _near_tie(gs, tol):
        assert tuple(out["bestgs_num"]) == ref["bestgs_num"]
    tops = sorted([ref["bestg_conf"], ref["bests_conf"], ref["bestgs_conf"]], reverse=True)
    if tops[0] - tops[1] > tol * scale:
        assert out["primitive"] == ref["primitive"]


# --------------------------------------------------------------------------------------------------
# (c) get_label_value, every branch of code/trainer.py:212-274
# --------------------------------------------------------------------------------------------------
def test_get_label_value_all_branches_vs_reference(gold2, scene_inputs):
    g = gold2["label_value"]
    scene, _, _, sc = scene_inputs
    masks = sc["masks"].astype(np.float64)
    trainers = {"reactive": _trainer("reactive", "fp32", g["num_rotations"]),
                "reinforcement": _trainer("reinforcement", "fp32", g["num_rotations"])}
    seen = set()
    for c in g["cases"]:
        a = c["args"]
        tr = trainers[c["method"]]
        lv, rv = tr.get_label_value(a["primitive_action"], a["objects_number"], a["suction_success"], a["grasp_success"],
                                    a["gs_success"], scene, masks, masks, a["bestg_id"], a["bests_id"], a["bestgs_g_id"],
                                    a["bestgs_s_id"], a["exploit_action"], 0.0, 0.0, 0.0)
        assert float(rv) == c["reward_value"], c
        assert abs(float(lv) - c["label_value"]) <= 1e-4 * max(1.0, abs(c["label_value"])), (c, lv)
        seen.add((c["method"], a["primitive_action"], a["exploit_action"], c["label_value"] == c["reward_value"]))
    assert len(g["cases"]) == 15 and len(seen) >= 10
    with pytest.raises(ValueError):
        trainers["reinforcement"].get_label_value("grasp", 4, 0, 1, 0, scene, masks, masks, [0, 0], [0, 0], [0, 0], [0, 0],
                                                  "push", 0.0, 0.0, 0.0)


# --------------------------------------------------------------------------------------------------
# (d) kink-free end-to-end gradients
# --------------------------------------------------------------------------------------------------
def _kink_free_state(sd, beta=None):
    """Every BatchNorm bias = +12: the ReLU kinks move 12 sigma away from the batch mean, so (almost) no pre-activation sits
    within rounding of a kink (the masks are still computed and applied).  With the kinks out of the way the same PyTorch
    code in fp32 and fp64 agrees to 2.3e-4 on every tensor (conv0: 5e-4, its max-pool is a kink of its own), instead of 6e-2
    at the reference's own initialisation - measured with oracle/qnet.py.  "Almost": the heightmap features are heavy-tailed
    (a flat background and a few object pixels many sigma away), so for a given bias a single element can still land on a
    kink and move its layer's gradients by a fixed amount whichever way the last bit of the statistics falls: at +6 one
    element of block-2 channel 177 (2e-3, profiles/debug_kinkfree.py); at +10 one element of the first dense layer (7.5e-3
    on three tensors once pool0 accumulated its statistics in double; with the earlier fp32 partial sums it fell on the
    oracle's side).  +8, +9, +11 and +12 are clean (every tensor <= 3e-4); SMG_TEST_BETA overrides the value."""
    if beta is None:
        beta = float(os.environ.get("SMG_TEST_BETA", "12"))
    out = {k: v.clone() for k, v in sd.items()}
    for k in out:
        if "norm" in k and k.endswith(".bias"):
            out[k] = torch.full_like(out[k], beta)
    return out


def check_kink_free_grads(grads, ref, precision, what):
    """fp32: EVERY tensor within 1e-3 (conv0 5e-3: it sits behind the max-pool's ties).  tf32 (operands rounded to 10
    mantissa bits in 120 stacked convolutions, forward and backward): stated tolerance median <= 2e-2, 95 % of the tensors
    <= 6e-2, every tensor <= 0.5 of its scale."""
    assert set(grads) == set(ref) and len(ref) == 368
    gscale = max(float(v.abs().max()) for v in ref.values())
    worst = []
    for k, r in ref.items():
        if k.endswith("features.norm5.weight") or k.endswith("features.norm5.bias"):
            # norm5 feeds the head's BatchNorm(2048) directly: its gradient is analytically zero, both sides hold noise
            assert float(grads[k].abs().max()) <= (1e-4 if precision == "fp32" else 1e-2) * gscale, k
            continue
        scale = r.double().abs().max()
        if ".norm" in k and k.endswith(".bias") or "-norm" in k and k.endswith(".bias"):
            # d beta = sum_p dz cancels to ~0 when no ReLU clips: dz is the data gradient of a convolution whose own
            # output gradient is a BatchNorm backward (zero sum per channel), so sum_p dz = W^T sum_p dy = 0 analytically and
            # both sides hold rounding noise of the SUMMANDS.  Their natural scale is the sibling d gamma = sum_p dz * xhat.
            scale = torch.maximum(scale, ref[k[:-len("bias")] + "weight"].double().abs().max())
        worst.append((float((grads[k].double() - r.double()).abs().max() / scale.clamp_min(1e-30)), k))
    worst.sort(reverse=True)
    med = worst[len(worst) // 2][0]
    print("kink-free gradients, %s (%s): worst %s, median %.2e" % (what, precision, [("%.2e" % e, k) for e, k in worst[:3]], med))
    if precision == "fp32":
        for e, k in worst:
            assert e <= (5e-3 if k.endswith("features.conv0.weight") else 1e-3), (k, e)
    else:
        assert med <= 2e-2 and worst[0][0] <= 0.5, (med, worst[0])
        assert sum(1 for e, _ in worst if e > 6e-2) <= 0.05 * len(worst), worst[:20]


@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_kink_free_gradients_every_tensor(scene_inputs, rl_state_dict, precision):
    import smg_b200.models as models
    scene, mask, _, _ = scene_inputs
    x, m = qnet.preprocess(scene, MEAN, STD), qnet.preprocess(mask, MEAN, STD)
    sd = _kink_free_state(rl_state_dict)
    torch.manual_seed(0)
    net = models.reinforcement_net(True)
    net.load_state_dict(sd)
    net = net.cuda()
    net.precision = precision
    net.gnum_rotations = net.snum_rotations = 16
    out = net.forward(x, m, 0, False, 3)
    label = 1.0
    d = out[0, 0, 0, 0] - label
    loss = 0.5 * (d ** 2) if abs(float(d)) < 1 else abs(d) - 0.5
    loss.sum().backward()
    ref_loss, ref = qnet.backprop_grads(sd, x, m, 0, 3, label, "reinforcement", gnum_rotations=16)
    assert abs(float(loss) - ref_loss) <= (1e-4 if precision == "fp32" else 2e-2) * max(1.0, abs(ref_loss))
    grads = {n: p.grad.detach().cpu() for n, p in net.named_parameters() if p.grad is not None}
    check_kink_free_grads(grads, ref, precision, "autograd path")


def test_backward_after_another_forward_fails_loudly(scene_inputs):
    """A volatile forward between the grad-enabled forward and backward() overwrites the saved activations: the
    backward must raise instead of silently differentiating the other pass."""
    import smg_b200.models as models
    scene, mask, _, _ = scene_inputs
    x, m = qnet.preprocess(scene, MEAN, STD), qnet.preprocess(mask, MEAN, STD)
    torch.manual_seed(0)
    net = models.reinforcement_net(True).cuda()
    out = net.forward(x, m, 0, False, 0)
    net.forward(x, m, 0, True, -1)
    with pytest.raises(RuntimeError):
        out.sum().backward()


def test_second_device_handle_sets_its_own_kernel_attributes():
    """cudaFuncSetAttribute is per device: a handle on another GPU of the same process must opt in again."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs in one process")
    from smg_b200 import engine
    x = torch.randn((1, 40, 40, 256))
    w = torch.randn((128, 96, 1, 1)) / 10
    outs = []
    for dev in (0, 1):
        eng = engine.Engine(dev, 2, 640, "tf32")
        with torch.cuda.device(dev):
            d = torch.device("cuda", dev)
            o, _ = eng.debug_conv("tf32", x.to(d), 96, torch.ones((1, 96), device=d), torch.zeros((1, 96), device=d), True, 0,
                                  w.to(d))
        outs.append(o.cpu())
    assert torch.allclose(outs[0], outs[1], atol=1e-5)
