"""GPU tests of the multi-GPU split of the path on ONE device: the sharded decision with the ranks
emulated in sequence (the all-gather replaced by concatenating the ranks' send buffers) and the data-parallel replay step
against the serial sequence.  The collectives themselves are covered by tests/test_dist_gloo.py (gloo, world 2) and by
bench.py's --verify at N > 1 (NCCL)."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR

pytestmark = pytest.mark.gpu


def _trainer(precision, rotations):
    from smg_b200.trainer import Trainer
    torch.manual_seed(0)
    tr = Trainer("reinforcement", 0.5, False, None, False, precision=precision)
    tr.model.gnum_rotations = tr.model.snum_rotations = rotations
    return tr


@pytest.mark.parametrize("world", [1, 3, 8])
def test_sharded_decision_equals_single_device_decision(world):
    import smg_b200.synth as synth
    from smg_b200 import decision, parallel
    with open(os.path.join(GOLDEN_DIR, "golden_r02.json")) as f:
        g = json.load(f)["hc"]
    sc = synth.make_scene(g["scene_seed"], num_objects=g["K"], cluttered=True)
    tr = _trainer("tf32", g["R"])
    tr.model.update_running_stats = False
    ref = decision.decide(tr, sc["depth"], sc["masks"], is_ets=True)
    ctx = parallel.decision_context(tr, sc["depth"], sc["masks"], world, True)
    sends = [parallel.decision_local_partials(tr, ctx, r) for r in range(world)]     # what the all-gather would deliver
    out = parallel.decision_from_partials(tr, ctx, torch.cat(sends), rank=None)
    scale = np.abs(ref["gra_conf"]).max()
    for k in ("gra_conf", "suc_conf", "gs_conf"):
        assert np.abs(out[k] - ref[k]).max() <= 1e-5 * scale, k          # same kernels, per-sample BatchNorm: same numbers
    for k in ("primitive", "bestg_id", "bests_id", "bestgs_num", "bestgs_g_id", "bestgs_s_id"):
        assert out[k] == ref[k], k
    gold = np.asarray(g["gra_conf"])
    assert np.abs(out["gra_conf"] - gold).max() <= 1e-2 * np.abs(gold).max()    # and the reference's own table (tf32 bar)
    if world == 1:
        full = parallel.decide_sharded(tr, sc["depth"], sc["masks"], is_ets=True)   # the real entry point, no process group
        assert full["primitive"] == ref["primitive"] and full["bestg_id"] == ref["bestg_id"]
        assert full["exchange_bytes"] == 98 * 400 * 64 * 4 or g["R"] != 16


def test_replay_batch_step_equals_serial_accumulation(scene_inputs):
    """backprop_batch (gradients-only fused steps, summed, averaged, ONE multi-tensor Adam launch) against the serial
    sequence on a second trainer: per-sample autograd passes accumulated into .grad, averaged, torch.optim.Adam.step()."""
    import smg_b200.synth as synth
    scene, _, _, sc = scene_inputs
    a, b = _trainer("fp32", 16), _trainer("fp32", 16)
    samples = []
    for i, (obj, rot, label) in enumerate([(0, 3, 1.0), (1, 7, 0.0), (2, 12, 2.5)]):
        samples.append({"depth_heightmap": scene, "m_depth_heightmap": synth.masked_scene(scene, sc["masks"], [obj]), "style": 0,
                        "rotation": rot, "label_value": label})
    loss_a, _ = a.backprop_batch(samples)
    eng = b.model._engine(2, 0)
    b.optimizer.zero_grad()
    losses = []
    for s in samples:
        x = eng.prep(torch.from_numpy(np.stack([s["depth_heightmap"], s["m_depth_heightmap"]])).cuda(), b.image_mean, b.image_std)
        out = b.model.forward(x[0:1], x[1:2], 0, False, s["rotation"])
        d = out[0, 0, 0, 0] - s["label_value"]
        loss = 0.5 * (d ** 2) if abs(float(d)) < 1 else abs(d) - 0.5
        loss.backward()                                   # accumulates into .grad
        losses.append(float(loss))
    for p in b.model.parameters():
        if p.grad is not None:
            p.grad /= len(samples)
    b.optimizer.step()
    assert abs(loss_a - np.mean(losses)) <= 2e-4 * max(1.0, abs(np.mean(losses)))
    pa, pb = dict(a.model.named_parameters()), dict(b.model.named_parameters())
    touched = [k for k, p in pb.items() if p.grad is not None]
    assert len(touched) == 368
    gscale = max(float(pb[k].grad.abs().max()) for k in touched)
    errs = sorted(float((pa[k].grad - pb[k].grad).abs().max()) / max(float(pb[k].grad.abs().max()), 1e-3 * gscale) for k in touched)
    print("replay batch vs serial: gradient error median %.2e, 95th %.2e, max %.2e" % (errs[len(errs) // 2], errs[int(0.95 * len(errs))], errs[-1]))
    assert errs[len(errs) // 2] <= 1e-2 and errs[int(0.95 * len(errs))] <= 5e-2 and errs[-1] <= 0.5
    k = "grasp_depth_trunk.features.denseblock2.denselayer5.conv1.weight"
    assert float(((pa[k].detach() - pb[k].detach()).abs() <= 0.1e-4).float().mean()) >= 0.85
    # BatchNorm running statistics: weighted-sum rule == the serial EMA over the six passes
    ba, bb = dict(a.model.named_buffers()), dict(b.model.named_buffers())
    for k in ("grasp_depth_trunk.features.denseblock3.denselayer9.norm2.running_var", "grasp_depth_trunk.features.norm0.running_mean",
              "graspnet_val.grasp-val-norm1.running_mean", "graspnet_val.grasp-val-norm0.running_var"):
        assert torch.allclose(ba[k], bb[k], rtol=2e-4, atol=1e-6), k


def test_replay_two_pipelines_equal_one_pipeline(scene_inputs, monkeypatch):
    """backprop_batch alternates consecutive samples between two handles (own workspace, streams and gradient buffer); the
    result must equal the single-pipeline run of the same batch: same mean loss, same averaged gradient (up to the order of
    the float sums), same weights after the Adam step, same BatchNorm running statistics."""
    import smg_b200.synth as synth
    scene, _, _, sc = scene_inputs
    samples = []
    for obj, rot, label in [(0, 3, 1.0), (1, 7, 0.0), (2, 12, 2.5), (3, 0, 0.3), (0, 9, 1.7), (1, 15, 0.8), (2, 5, 0.1)]:
        samples.append({"depth_heightmap": scene, "m_depth_heightmap": synth.masked_scene(scene, sc["masks"], [obj]), "style": 0,
                        "rotation": rot, "label_value": label})
    a, b = _trainer("tf32", 16), _trainer("tf32", 16)
    monkeypatch.setenv("SMG_REPLAY_STREAMS", "1")
    loss_b, _ = b.backprop_batch(samples)
    monkeypatch.setenv("SMG_REPLAY_STREAMS", "2")
    loss_a, _ = a.backprop_batch(samples)
    assert "replay2" in a._fused[0] and "replay2" not in b._fused[0]
    assert abs(loss_a - loss_b) <= 1e-5 * max(1.0, abs(loss_b))
    pa, pb = dict(a.model.named_parameters()), dict(b.model.named_parameters())
    worst = 0.0
    for k, p in pb.items():
        if p.grad is not None:
            worst = max(worst, float((pa[k].grad - p.grad).abs().max()) / max(float(p.grad.abs().max()), 1e-12))
    print("two pipelines vs one: worst per-tensor deviation of the averaged gradient %.2e" % worst)
    assert worst <= 2e-4      # same per-sample gradients (atomics noise ~4e-5), summed in a different order
    k = "grasp_depth_trunk.features.denseblock2.denselayer5.conv1.weight"
    assert float((pa[k].detach() - pb[k].detach()).abs().max()) <= 2.1e-4          # one Adam step of lr = 1e-4 each way at most
    ba, bb = dict(a.model.named_buffers()), dict(b.model.named_buffers())
    for k in ("grasp_depth_trunk.features.denseblock3.denselayer9.norm2.running_var", "grasp_depth_trunk.features.norm0.running_mean",
              "graspnet_val.grasp-val-norm1.running_mean"):
        assert torch.allclose(ba[k], bb[k], rtol=2e-4, atol=1e-6), k
    # second batch: the second handle must have picked up the updated weights (a stale copy would reproduce the first loss)
    loss_a2, _ = a.backprop_batch(samples)
    monkeypatch.setenv("SMG_REPLAY_STREAMS", "1")
    loss_b2, _ = b.backprop_batch(samples)
    print("losses: first batch %.6f / %.6f, second batch %.6f / %.6f" % (loss_a, loss_b, loss_a2, loss_b2))
    assert abs(loss_a2 - loss_b2) <= 2e-2 * max(1.0, abs(loss_b2)) and abs(loss_a2 - loss_a) > 1e-6


def test_secondary_handles_follow_in_place_training_steps(scene_inputs):
    """The fused training step updates the parameters in place from a library kernel (torch's version counters do not move);
    the per-primitive handles of the sharded decision and the second replay pipeline must still re-pack: after a training
    step the sharded decision (secondary handles) and `decision.decide` (the model's own handle) give the same tables."""
    import smg_b200.synth as synth
    from smg_b200 import decision, parallel
    sc = synth.make_scene(5, num_objects=3, cluttered=False)
    masks = sc["masks"].astype(np.float64)
    tr = _trainer("tf32", 4)
    tr.model.update_running_stats = False
    before = parallel.decide_sharded(tr, sc["depth"], masks, is_ets=True)            # secondary handles pack the initial weights
    for i in range(2):                                                                # two in-place updates of the grasp trunk + head
        tr.backprop(sc["depth"] * masks.sum(0), "grasp", [i, 1], [0, 0], [], [], 1.0, masks.copy(), [0] * 3, [0] * 3, [])
    ref = decision.decide(tr, sc["depth"], masks, is_ets=True)
    out = parallel.decide_sharded(tr, sc["depth"], masks, is_ets=True)
    scale = np.abs(ref["gra_conf"]).max()
    assert np.abs(ref["gra_conf"] - before["gra_conf"]).max() > 1e-3 * scale, "the training steps must have changed the table"
    for k in ("gra_conf", "suc_conf", "gs_conf"):
        assert np.abs(out[k] - ref[k]).max() <= 1e-5 * scale, k
