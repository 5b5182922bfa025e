"""GPU parity of the trunk / head / Q pass through the C ABI against the CPU oracle and the golden
fixtures (values produced by the unmodified reference).

Tolerances (BASELINE.json north_star): fp32 mode <= 1e-4, reduced-precision modes <= 1e-2, both
measured as |dQ| / max_candidates |Q_ref| and, for activations, max|d| / max|ref| per tensor
(point-wise relative error is meaningless for a 25 600-term dot product).
"""
import numpy as np
import pytest
import torch

from conftest import MEAN, STD, check_fingerprint
from oracle import qnet

pytestmark = pytest.mark.gpu

TOL = {"fp32": 1e-4, "tf32": 1e-2, "bf16": 1e-1}  # bf16: measured, SURVEY.md section 7 predicts 3e-2..7e-2
FEAT_TOL = {"fp32": 1e-4, "tf32": 2e-2, "bf16": 2e-1}


def relmax(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.fixture(scope="module")
def net():
    import smg_b200.models as models
    torch.manual_seed(0)
    n = models.reinforcement_net(True).cuda()
    n.train()
    return n


@pytest.fixture(scope="module")
def inputs(scene_inputs):
    scene, mask, pair, _ = scene_inputs
    return tuple(qnet.preprocess(v, MEAN, STD) for v in (scene, mask, pair))


@pytest.fixture(scope="module")
def oracle_taps(rl_state_dict, inputs):
    taps = {}
    with torch.no_grad():
        f = qnet.densenet_features(rl_state_dict, "grasp_depth_trunk.features.", inputs[0], taps)
    taps["feat"] = f
    return taps


@pytest.mark.parametrize("precision", ["fp32", "tf32", "bf16"])
def test_trunk_activations(net, inputs, oracle_taps, golden, precision):
    net.precision = precision
    eng = net._engine(2)
    x = torch.cat([inputs[0], inputs[1]]).cuda()
    feat = eng.trunk_forward(1, x)  # grasp trunk, 2 samples: per-sample BN must keep them independent
    H = 640
    shapes = {"conv0": (64, H // 2, H // 2), "pool0": (64, H // 4, H // 4), "block1": (256, H // 4, H // 4),
              "trans1": (128, H // 8, H // 8), "block2": (512, H // 8, H // 8), "trans2": (256, H // 16, H // 16),
              "block3": (1024, H // 16, H // 16), "trans3": (512, H // 32, H // 32), "block4": (1024, H // 32, H // 32)}
    report = {}
    for name, shp in shapes.items():
        got = eng.debug_read(name, 0, shp)
        report[name] = relmax(got, oracle_taps[name][0])
    report["feat"] = relmax(feat[0], oracle_taps["feat"][0])
    print("%s activation rel-max errors: %s" % (precision, {k: "%.2e" % v for k, v in report.items()}))
    assert report["conv0"] <= 1e-5 and report["pool0"] <= 1e-5  # stem is fp32 in every mode
    for name, err in report.items():
        assert err <= FEAT_TOL[precision], "%s: %s rel-max error %.3g" % (precision, name, err)
    if precision == "fp32":
        check_fingerprint(feat[0:1], golden["trunk_taps"]["norm5"], 1e-4)
    # the second sample (masked scene) must match its own single-sample oracle pass
    with torch.no_grad():
        f1 = qnet.densenet_features({k: v.cpu() for k, v in net.state_dict().items()},
                                    "grasp_depth_trunk.features.", inputs[1])
    assert relmax(feat[1], f1[0]) <= FEAT_TOL[precision]


@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_q_R1_all_styles_vs_reference(net, inputs, golden, precision):
    net.precision = precision
    x, m, m2 = inputs
    ref = [golden["q"]["rl_style%d_R1" % s][0] for s in (0, 1, 2)]
    scale = max(abs(v) for v in ref)
    for style in (0, 1, 2):
        out = net.forward(x, m2 if style == 2 else m, style, True, -1)
        assert isinstance(out, list) and len(out) == 1 and tuple(out[0].shape) == (1, 1, 1, 1) and out[0].is_cuda
        err = abs(float(out[0]) - ref[style]) / scale
        print("%s style %d: Q=%.6f ref=%.6f err/scale=%.2e" % (precision, style, float(out[0]), ref[style], err))
        assert err <= TOL[precision]


@pytest.mark.parametrize("precision", ["fp32", "tf32", "bf16"])
def test_q_R16_vs_reference(net, inputs, golden, precision):
    net.precision = precision
    net.gnum_rotations = net.snum_rotations = 16
    try:
        x, m, m2 = inputs
        ref = np.array(golden["q"]["rl_style0_R16"])
        out = net.forward(x, m, 0, True, -1)
        got = np.array([float(o) for o in out])
        err = np.abs(got - ref).max() / np.abs(ref).max()
        print("%s R=16: max|dQ|/max|Q| = %.2e, argmax got %d ref %d" % (precision, err, got.argmax(), ref.argmax()))
        assert len(out) == 16 and err <= TOL[precision]
        srt = np.sort(ref)
        if srt[-1] - srt[-2] > 2 * TOL[precision] * np.abs(ref).max():  # not a stated near-tie
            assert got.argmax() == ref.argmax()
        if precision != "bf16":
            q1 = net.forward(x, m, 1, True, 5)
            assert tuple(q1.shape) == (1, 1, 1, 1)
            assert abs(float(q1) - golden["q"]["rl_style1_R16_rot5"][0]) / np.abs(ref).max() <= TOL[precision]
            q2 = net.forward(x, m2, 2, True, 5)
            assert abs(float(q2) - golden["q"]["rl_style2_R16_rot5"][0]) / np.abs(ref).max() <= TOL[precision]
    finally:
        net.gnum_rotations = net.snum_rotations = 1


def test_dedup_table_equals_single_calls(net, scene_inputs):
    """forward_all-style evaluation (many masks x rotations in one pass) == independent single calls."""
    import smg_b200.synth as synth
    net.precision = "fp32"
    scene, _, _, sc = scene_inputs
    masks = np.stack([synth.masked_scene(scene, sc["masks"], [k]) for k in range(3)])
    eng = net._engine(4 + 3)
    hm_s = torch.from_numpy(scene).cuda()
    hm_m = torch.from_numpy(masks).cuda()
    table = eng.qforward_maps(0, hm_s, hm_m, MEAN, STD, [0, 1, 2, 3], 16).cpu().numpy()[:, :, 0]
    for k in (0, 2):
        for r in (1, 3):
            single = eng.qforward_maps(0, hm_s, hm_m[k:k + 1], MEAN, STD, [r], 16).cpu().numpy()[0, 0, 0]
            assert abs(single - table[k, r]) <= 1e-5 * max(1.0, np.abs(table).max())


def test_maps_path_tensor_core_stem(net, scene_inputs):
    """Heightmap entry point (one input channel): conv0 runs on the tensor cores in tf32 AND fp32 mode (stem_umma.cu: TMA
    patch -> im2col operand -> tcgen05.mma with hi/lo split operands); its raw output must agree with a float64 convolution
    of the same input to fp32 accuracy, and the two modes must agree on the pooled block-1 input and on Q."""
    import smg_b200.synth as synth
    scene, _, _, sc = scene_inputs
    masks = np.stack([synth.masked_scene(scene, sc["masks"], [k]) for k in range(2)])
    hm_s = torch.from_numpy(scene).cuda()
    hm_m = torch.from_numpy(masks).cuda()
    H = 640
    got = {}
    for prec in ("fp32", "tf32"):
        net.precision = prec
        eng = net._engine(4 + 3)
        q = eng.qforward_maps(0, hm_s, hm_m, MEAN, STD, [0, 1, 2, 3], 16).cpu().numpy()
        got[prec] = (q, eng.debug_read("conv0", 1, (64, H // 2, H // 2)).cpu(), eng.debug_read("pool0", 1, (64, H // 4, H // 4)).cpu())
    x1 = qnet.rotate_nearest(qnet.preprocess(scene, MEAN, STD), 1, 16)   # sample 1 of the pass = the scene at rotation 1
    w0 = net.state_dict()["grasp_depth_trunk.features.conv0.weight"].cpu().double()
    ref0 = torch.nn.functional.conv2d(x1.double(), w0, stride=2, padding=3)[0]
    c_err = max(float((got[p][1].double() - ref0).abs().max() / ref0.abs().max()) for p in ("fp32", "tf32"))
    p_err = float((got["tf32"][2] - got["fp32"][2]).abs().max() / got["fp32"][2].abs().max())
    q_err = float(np.abs(got["tf32"][0] - got["fp32"][0]).max() / np.abs(got["fp32"][0]).max())
    print("tensor-core stem vs fp32: conv0 %.2e pool0 %.2e Q %.2e" % (c_err, p_err, q_err))
    assert c_err <= 1e-5 and p_err <= 1e-5          # 3xTF32 split: fp32 accuracy
    assert q_err <= TOL["tf32"]
    net.precision = "fp32"


def test_batched_units_equal_single_calls(scene_inputs):
    """Trainer.forward_batch (G units in one pass) == G independent Trainer.forward calls."""
    import smg_b200.synth as synth
    from smg_b200.trainer import Trainer
    torch.manual_seed(0)
    tr = Trainer("reinforcement", 0.5, False, None, False, precision="fp32")
    tr.model.gnum_rotations = tr.model.snum_rotations = 4
    tr.model.update_running_stats = False
    scenes, masks = [], []
    for seed in (1, 2, 3):
        sc = synth.make_scene(seed, num_objects=4, cluttered=False)
        scenes.append(sc["scene"])
        masks.append(synth.masked_scene(sc["scene"], sc["masks"], [seed % 4]))
    batch = tr.forward_batch(np.stack(scenes), np.stack(masks), style=0)
    assert batch.shape == (3, 1, 4)
    for g in range(3):
        single = tr.forward(scenes[g], masks[g], 0, True, False)
        assert np.abs(batch[g, 0] - single).max() <= 1e-5 * max(1.0, np.abs(single).max())
    batch2 = tr.forward_batch(np.stack(scenes), np.stack(masks), style=0)   # second call: CUDA-graph replay
    assert np.abs(batch2 - batch).max() <= 1e-5


def test_batched_units_running_stats_equal_serial_calls(scene_inputs):
    """The BatchNorm running-statistics side effect of Trainer.forward_batch == the same units evaluated one after the other
    by Trainer.forward (what the reference's step loop does, one train-mode forward per call)."""
    import smg_b200.synth as synth
    from smg_b200.trainer import Trainer
    scenes, masks = [], []
    for seed in (1, 2):
        sc = synth.make_scene(seed, num_objects=4, cluttered=False)
        scenes.append(sc["scene"])
        masks.append(synth.masked_scene(sc["scene"], sc["masks"], [seed % 4]))
    states = []
    for batched in (True, False):
        torch.manual_seed(0)
        tr = Trainer("reinforcement", 0.5, False, None, False, precision="tf32")
        tr.model.gnum_rotations = tr.model.snum_rotations = 4
        if batched:
            tr.forward_batch(np.stack(scenes), np.stack(masks), style=0)
        else:
            for g in range(2):
                tr.forward(scenes[g], masks[g], 0, True, False)
        states.append({k: v.clone() for k, v in tr.model.state_dict().items() if "running" in k or "num_batches" in k})
    a, b = states
    for k in ("grasp_depth_trunk.features.norm0.running_mean", "grasp_depth_trunk.features.denseblock4.denselayer16.norm2.running_var",
              "grasp_depth_trunk.features.norm5.running_var", "graspnet_val.grasp-val-norm0.running_var",
              "graspnet_val.grasp-val-norm1.running_mean", "grasp_depth_trunk.features.norm0.num_batches_tracked",
              "graspnet_val.grasp-val-norm1.num_batches_tracked"):
        assert torch.allclose(a[k].double(), b[k].double(), rtol=1e-4, atol=1e-6), k
    assert int(a["grasp_depth_trunk.features.norm0.num_batches_tracked"]) == 16      # 2 units x 4 rotations x (scene, mask)


def test_reactive_logits_vs_reference(inputs, golden):
    import smg_b200.models as models
    torch.manual_seed(0)
    n = models.reactive_net(True).cuda()
    x, m, m2 = inputs
    out = n.forward(x, m, 0, True, -1)
    ref = np.array(golden["reactive_style0_R1"])
    assert tuple(out[0].shape) == (1, 3, 1, 1)
    assert np.abs(out[0].view(-1).cpu().numpy() - ref).max() / np.abs(ref).max() <= 1e-4
    out = n.forward(x, m2, 2, True, -1)
    ref = np.array(golden["reactive_style2_R1"])
    assert np.abs(out[0].view(-1).cpu().numpy() - ref).max() / np.abs(ref).max() <= 1e-4


def test_running_stats_side_effect(inputs, golden):
    """BatchNorm running_mean/var/num_batches_tracked after one forward == torch's own EMA."""
    import smg_b200.models as models
    torch.manual_seed(0)
    n = models.reinforcement_net(True).cuda()
    x, m, _ = inputs
    n.forward(x, m, 0, True, -1)
    sd = {k: v.cpu() for k, v in n.state_dict().items()}
    # reference semantics: trunk(scene) then trunk(mask), both in train mode (code/models.py:384-385)
    import torchvision
    ref = torchvision.models.densenet121(weights=None).features
    ref.load_state_dict({k[len("grasp_depth_trunk.features."):]: v for k, v in golden_like_state(n).items()})
    ref.train()
    with torch.no_grad():
        ref(x)
        ref(m)
    for name in ("norm0", "denseblock2.denselayer3.norm1", "denseblock4.denselayer16.norm2", "norm5"):
        for stat in ("running_mean", "running_var"):
            a = sd["grasp_depth_trunk.features.%s.%s" % (name, stat)]
            b = ref.state_dict()["%s.%s" % (name, stat)]
            assert relmax(a, b) <= 1e-4, (name, stat)
        assert int(sd["grasp_depth_trunk.features.%s.num_batches_tracked" % name]) == 2
    # the head's BatchNorm2d(2048) and BatchNorm2d(64) see cat(features(scene), features(mask)) once
    import smg_b200.models as models2
    torch.manual_seed(0)
    fresh = models2.reinforcement_net(True)
    fresh.train()
    with torch.no_grad():
        feat = torch.cat((ref(x), ref(m)), dim=1)   # third and fourth pass: statistics of `ref` are not used below
        fresh.graspnet_val(feat)
    hsd = fresh.graspnet_val.state_dict()
    for key in ("grasp-val-norm0.running_mean", "grasp-val-norm0.running_var", "grasp-val-norm1.running_mean",
                "grasp-val-norm1.running_var"):
        a, b = sd["graspnet_val." + key].double(), hsd[key].double()
        # norm0's batch mean is norm5.bias (0 at init): torch accumulates ~1e-9 of rounding noise there, so absolute
        assert float((a - b).abs().max()) <= 1e-4 * max(float(b.abs().max()), 1e-2), key
    assert int(sd["graspnet_val.grasp-val-norm1.num_batches_tracked"]) == 1


def golden_like_state(n):
    """Initial (pre-forward) grasp-trunk state: weights of seed 0 with fresh running statistics."""
    import smg_b200.models as models
    torch.manual_seed(0)
    fresh = models.reinforcement_net(True).state_dict()
    return {k: v for k, v in fresh.items() if k.startswith("grasp_depth_trunk.features.")}


def test_trainer_forward_dropin(scene_inputs, golden):
    from smg_b200.trainer import Trainer
    torch.manual_seed(0)
    tr = Trainer("reinforcement", 0.5, False, None, False)
    scene, mask, _, _ = scene_inputs
    pred = tr.forward(scene, mask, style=0, is_volatile=True, is_target=False)
    assert isinstance(pred, np.ndarray) and pred.shape == (1,) and pred.dtype == np.float64
    assert abs(pred[0] - golden["trainer_forward_rl_style0"][0]) <= 1e-4 * abs(golden["trainer_forward_rl_style0"][0])
    tgt = tr.forward(scene, mask, style=0, is_volatile=True, is_target=True)
    assert abs(tgt[0] - pred[0]) <= 1e-5  # model_target starts as a copy (code/trainer.py:74-75)
