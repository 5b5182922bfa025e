"""GPU parity of the training step (Trainer.backprop, code/trainer.py:278-384) through the C ABI:
loss, all 368 parameter gradients of the touched trunk + head, the Adam update and the BatchNorm
running statistics, against the CPU oracle (autograd over the restated network) and the golden
fixtures recorded from the unmodified reference.

Tolerance.  This network (120 ReLUs behind train-mode BatchNorm, bias 0 at init so every ReLU kink sits at the
batch mean) does not have gradients that are stable to fp32 rounding: the SAME PyTorch code evaluated in fp32 and
fp64 disagrees by up to 6e-2 per tensor (max|d|/max|ref|; median 6e-4; 100 of 368 tensors above 1e-3 - measured
with oracle/qnet.py) because a handful of pixels sit within rounding of a kink and flip
their mask.  The kernels themselves are exact to 1e-7 where no kink is involved (tests/test_gpu_bn_bwd.py, and
every tensor up to the first flipped pixel here agrees to 5e-6).  Which pixels flip also changes from run to run
(the forward statistics are accumulated with atomics), so the end-to-end bar is statistical: >= 95 % of the tensors
within 3e-2 (max|d|/max|ref|), every tensor within 0.5, median over tensors <= 1e-2, cosine similarity >= 0.98
(observed over several runs: worst 4e-2..2e-1 on one or two tensors, median 3e-3)."""
import numpy as np
import pytest
import torch

from conftest import MEAN, STD, check_fingerprint
from oracle import qnet

pytestmark = pytest.mark.gpu

GRAD_TOL = 3e-2      # bar for >= 95 % of the tensors
GRAD_TOL_ALL = 0.5   # bar for every tensor
MEDIAN_TOL = 1e-2


def relmax(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def make_net(kind="reinforcement"):
    import smg_b200.models as models
    torch.manual_seed(0)
    net = (models.reinforcement_net if kind == "reinforcement" else models.reactive_net)(True)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    net = net.cuda()
    net.train()
    return net, sd


@pytest.fixture(scope="module")
def inputs(scene_inputs):
    scene, mask, pair, _ = scene_inputs
    return tuple(qnet.preprocess(v, MEAN, STD) for v in (scene, mask, pair))


def collect_grads(net):
    return {n: p.grad.detach().cpu() for n, p in net.named_parameters() if p.grad is not None}


def compare_all(grads, ref, tol=GRAD_TOL):
    assert set(grads) == set(ref), "gradient key sets differ: %s" % sorted(set(grads) ^ set(ref))[:5]
    worst = []
    scale = max(float(v.abs().max()) for v in ref.values())
    for k in ref:
        if float(ref[k].abs().max()) < 1e-5 * scale:  # analytically-zero gradients (norm5 feeds another BatchNorm)
            assert float(grads[k].abs().max()) < 1e-4 * scale, k
            continue
        cos = float(torch.nn.functional.cosine_similarity(grads[k].double().flatten(), ref[k].double().flatten(), dim=0))
        assert cos >= (0.98 if ref[k].numel() >= 4096 else 0.9), (k, cos)   # one flipped channel dominates a small vector
        worst.append((relmax(grads[k], ref[k]), k))
    worst.sort(reverse=True)
    med = worst[len(worst) // 2][0]
    print("gradient errors: worst %s, median %.2e" % ([("%.2e" % e, k) for e, k in worst[:4]], med))
    bad = [(e, k) for e, k in worst if e > tol]
    assert len(bad) <= 0.05 * len(worst), "gradient mismatch: %s" % bad[:8]
    assert worst[0][0] <= GRAD_TOL_ALL, worst[0]
    assert med <= MEDIAN_TOL
    return worst[0][0]


def test_rl_grads_style0_vs_oracle_and_golden(inputs, golden):
    net, sd = make_net()
    x, m, _ = inputs
    g = golden["backprop_rl_grasp"]
    out = net.forward(x, m, 0, False, 0)
    assert out.requires_grad and tuple(out.shape) == (1, 1, 1, 1) and net.gra_prob is out
    d = net.gra_prob[0, 0, 0, 0] - g["label"]
    loss = 0.5 * (d ** 2) if abs(float(d)) < 1 else abs(d) - 0.5   # code/trainer.py:345-348
    loss.sum().backward()
    assert abs(float(loss) - g["loss"]) <= 1e-4 * max(1.0, g["loss"])
    grads = collect_grads(net)
    assert len(grads) == g["n_grads"] == 368
    for k, fp in g["grads"].items():               # the reference's own gradients
        check_fingerprint(grads[k], fp, 1e-1)
    _, ref = qnet.backprop_grads(sd, x, m, 0, 0, g["label"], "reinforcement")
    compare_all(grads, ref)


def test_rl_grads_style2_rotated_vs_oracle(inputs):
    net, sd = make_net()
    net.gnum_rotations = net.snum_rotations = 16
    x, _, m2 = inputs
    out = net.forward(x, m2, 2, False, 5)          # ES: gs trunk + suction head, rotation pinned to 0
    d = net.gs_prob[0, 0, 0, 0] - 2.5
    loss = 0.5 * (d ** 2) if abs(float(d)) < 1 else abs(d) - 0.5
    loss.sum().backward()
    grads = collect_grads(net)
    ref_loss, ref = qnet.backprop_grads(sd, x, m2, 2, 5, 2.5, "reinforcement", gnum_rotations=16)
    assert abs(float(loss) - ref_loss) <= 1e-4 * max(1.0, abs(ref_loss))
    assert all(k.startswith(("gs_depth_trunk.", "suctionnet_val.")) for k in grads)
    compare_all(grads, ref)


def test_rl_grads_style1_rotation3_vs_oracle(inputs):
    net, sd = make_net()
    net.gnum_rotations = net.snum_rotations = 16
    x, m, _ = inputs
    net.forward(x, m, 1, False, 3)
    d = net.suc_prob[0, 0, 0, 0] - 0.0
    loss = 0.5 * (d ** 2) if abs(float(d)) < 1 else abs(d) - 0.5
    loss.sum().backward()
    _, ref = qnet.backprop_grads(sd, x, m, 1, 3, 0.0, "reinforcement", gnum_rotations=16)
    compare_all(collect_grads(net), ref)


def test_trainer_backprop_rl_dropin(scene_inputs, golden):
    from smg_b200.trainer import Trainer
    scene, mask, _, sc = scene_inputs
    torch.manual_seed(0)
    tr = Trainer("reinforcement", 0.5, False, None, False)
    g = golden["backprop_rl_grasp"]
    tr.forward(scene, mask, style=0, is_volatile=True, is_target=False)   # same call sequence as the golden run
    before = {k: v.detach().clone() for k, v in tr.model.state_dict().items()}
    masks = sc["masks"].astype(np.float64).copy()
    loss = tr.backprop(scene, "grasp", [0, 0], [0, 0], [], [], g["label"], masks, [0] * 4, [0] * 4, [])
    assert abs(float(loss) - g["loss"]) <= 1e-4 * g["loss"]
    after = tr.model.state_dict()
    # Adam's first step is -lr * g / (|g| + eps) ~ -lr * sign(g): entries whose gradient is within the fp32 noise of
    # zero may flip; everything else must match what the reference recorded
    for k, fp in g["param_delta"].items():
        if ".norm5." in k:      # norm5 feeds another BatchNorm: its gradient is analytically zero, Adam steps on noise
            continue
        d = (after[k] - before[k]).cpu().double().numpy().ravel()[np.asarray(fp["pos"])]
        ok = np.abs(d - np.asarray(fp["val"])) <= 0.05 * 1e-4
        assert ok.mean() >= 0.85, (k, ok.mean())
    check_fingerprint(after["grasp_depth_trunk.features.denseblock2.denselayer3.norm1.running_mean"].cpu(),
                      g["bn_running_mean_after"], 1e-4)
    check_fingerprint(after["grasp_depth_trunk.features.denseblock2.denselayer3.norm1.running_var"].cpu(),
                      g["bn_running_var_after"], 1e-4)
    untouched = "suction_depth_trunk.features.conv0.weight"
    assert torch.equal(after[untouched], before[untouched])   # Adam skips parameters without gradient


def test_trainer_backprop_reactive_dropin(scene_inputs, golden):
    from smg_b200.trainer import Trainer
    scene, mask, _, sc = scene_inputs
    torch.manual_seed(0)
    tr = Trainer("reactive", 0.5, False, None, False)
    g = golden["backprop_reactive_suction"]
    pred = tr.forward(scene, mask, style=1, is_volatile=True, is_target=False)
    assert abs(float(pred[0]) - golden["trainer_forward_reactive_style1"]) <= 1e-4
    masks = sc["masks"].astype(np.float64).copy()
    loss = tr.backprop(scene, "suction", [0, 0], [0, 0], [], [], g["label"], masks, [0] * 4, [0] * 4, [])
    assert abs(float(loss) - g["loss"]) <= 1e-4 * g["loss"]


def test_reactive_grads_vs_golden(inputs, golden):
    net, sd = make_net("reactive")
    x, m, _ = inputs
    g = golden["backprop_reactive_suction"]
    out = net.forward(x, m, 1, False, 0)
    w = torch.tensor([1.0, 1.0, 0.0], device=out.device)
    target = torch.full((1, 1, 1), int(g["label"]), dtype=torch.long, device=out.device)
    loss = torch.nn.functional.nll_loss(torch.log_softmax(out.view(1, 3, 1, 1), dim=1), target, weight=w)
    loss.sum().backward()
    assert abs(float(loss) - g["loss"]) <= 1e-4 * g["loss"]
    grads = collect_grads(net)
    assert len(grads) == g["n_grads"]
    for k, fp in g["grads"].items():
        check_fingerprint(grads[k], fp, 1e-1)


def test_fused_adam_matches_torch():
    from smg_b200 import engine
    eng = engine.get_engine(0, 2, 640, "fp32", owner="adam")
    torch.manual_seed(3)
    ps = [torch.randn(n, device="cuda") for n in (7, 4096, 100001)]
    gs = [torch.randn_like(p) * 10 ** (-i) for i, p in enumerate(ps)]
    ref = [p.clone().requires_grad_(True) for p in ps]
    opt = torch.optim.Adam(ref, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0)
    m = [torch.zeros_like(p) for p in ps]
    v = [torch.zeros_like(p) for p in ps]
    for step in (1, 2, 3):
        for r, g in zip(ref, gs):
            r.grad = g.clone()
        opt.step()
        eng.adam_step(ps, gs, m, v, step)
        for a, b in zip(ps, ref):
            assert float((a - b.detach()).abs().max()) <= 2e-7
