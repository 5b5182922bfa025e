"""Generate the golden fixtures in tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py

The reference ships no tests or golden vectors, so parity is
pinned on outputs of its own modules, imported through oracle/refshim.py (matplotlib /
apex stubs, densenet121(weights=None), .cuda() -> identity for CPU execution).
`Trainer.forward` as published divides by std = 0; the Trainer
fixtures are produced from an IN-MEMORY copy of trainer.py whose two literals are
replaced by mean 0.01 / std 0.03 (a harness choice, recorded in every fixture).

Recorded library versions matter: the arithmetic lives in torch / torchvision / numpy /
cv2, which the reference does not pin.
"""
import hashlib
import json
import os
import sys
import time
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = os.path.dirname(os.path.abspath(__file__))

from oracle import refshim  # noqa: E402

mods = refshim.install()
import cv2  # noqa: E402
import scipy  # noqa: E402
import torch  # noqa: E402
import torchvision  # noqa: E402

import smg_b200.synth as synth  # noqa: E402
from oracle import qnet  # noqa: E402

MEAN, STD = 0.01, 0.03
VERSIONS = {"torch": torch.__version__, "torchvision": torchvision.__version__, "numpy": np.__version__,
            "cv2": cv2.__version__, "scipy": scipy.__version__, "image_mean": MEAN, "image_std": STD}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def fingerprint(t, npos=32, seed=0):
    """Small, position-addressed summary of a tensor."""
    a = t.detach().cpu().double().numpy().ravel()
    rs = np.random.RandomState(seed)
    pos = rs.randint(0, a.size, size=npos)
    return {"shape": list(t.shape), "mean": float(a.mean()), "absmean": float(np.abs(a).mean()),
            "max": float(a.max()), "min": float(a.min()), "pos": pos.tolist(), "val": a[pos].tolist()}


def patched_trainer_module():
    """trainer.py with the NaN-producing literals replaced, compiled in memory (nothing written to disk)."""
    src = open(os.path.join(refshim.REFERENCE_CODE, "trainer.py")).read()
    assert "image_mean = [0.0, 0.0, 0.0]" in src and "image_std = [0.0, 0.0, 0.0]" in src
    src = src.replace("image_mean = [0.0, 0.0, 0.0]", "image_mean = [%r, %r, %r]" % (MEAN, MEAN, MEAN))
    src = src.replace("image_std = [0.0, 0.0, 0.0]", "image_std = [%r, %r, %r]" % (STD, STD, STD))
    m = types.ModuleType("trainer_patched")
    exec(compile(src, "trainer_patched.py", "exec"), m.__dict__)
    return m


def main():
    torch.set_num_threads(os.cpu_count())
    gold = {"versions": VERSIONS}
    models = mods["models"]

    # ---------------------------------------------------------------- weights
    torch.manual_seed(0)
    ref_rl = models.reinforcement_net(True)
    ref_rl.train()
    sd = ref_rl.state_dict()
    import smg_b200.models as mymodels
    torch.manual_seed(0)
    my_rl = mymodels.reinforcement_net(True)
    my_sd = my_rl.state_dict()
    assert list(sd.keys()) == list(my_sd.keys()), "state_dict keys differ"
    assert all(torch.equal(sd[k], my_sd[k]) for k in sd), "same-seed weights differ"
    gold["weights"] = {
        "seed": 0, "n_keys": len(sd),
        "keys_sha": hashlib.sha256("\n".join(sd.keys()).encode()).hexdigest(),
        "conv0_sha": sha(sd["grasp_depth_trunk.features.conv0.weight"].numpy()),
        "grasp_head_conv1_sha": sha(sd["graspnet_val.grasp-val-conv1.weight"].numpy()),
        "sum_abs": float(sum(v.double().abs().sum() for k, v in sd.items() if v.is_floating_point())),
    }
    print("weights ok: %d keys" % len(sd))

    # ---------------------------------------------------------------- inputs
    sc = synth.make_scene(1, num_objects=4, cluttered=False)
    scene_hm = sc["scene"]
    mask_hm = synth.masked_scene(sc["scene"], sc["masks"], [0])
    pair_hm = synth.masked_scene(sc["scene"], sc["masks"], [0, 1])
    gold["inputs"] = {"scene_seed": 1, "scene_sha": sha(scene_hm), "mask_sha": sha(mask_hm), "pair_sha": sha(pair_hm)}
    x = qnet.preprocess(scene_hm, MEAN, STD)
    m = qnet.preprocess(mask_hm, MEAN, STD)
    m2 = qnet.preprocess(pair_hm, MEAN, STD)

    # ---------------------------------------------------------------- trunk activations (hooks on the reference module)
    feats = {}
    trunk = ref_rl.grasp_depth_trunk.features
    hooks = []
    for name in ("conv0", "pool0", "denseblock1", "transition1", "denseblock2", "transition2", "denseblock3",
                 "transition3", "denseblock4", "norm5"):
        hooks.append(getattr(trunk, name).register_forward_hook(
            lambda mod, inp, out, name=name: feats.__setitem__(name, out.detach().clone())))
    with torch.no_grad():
        f = trunk(x)
    for hk in hooks:
        hk.remove()
    gold["trunk_taps"] = {k: fingerprint(v) for k, v in feats.items()}
    # the oracle must reproduce them
    taps = {}
    with torch.no_grad():
        fo = qnet.densenet_features(sd, "grasp_depth_trunk.features.", x, taps)
    err = float((fo - f).abs().max() / f.abs().max())
    print("oracle trunk vs reference: rel-max err %.3g" % err)
    assert err < 1e-5
    gold["trunk_feat_relerr_oracle"] = err

    # ---------------------------------------------------------------- RL forward, R = 1
    q = {}
    t0 = time.time()
    with torch.no_grad():
        for style in (0, 1, 2):
            out = ref_rl.forward(x, m2 if style == 2 else m, style, True, -1)
            q["rl_style%d_R1" % style] = [float(o.view(-1)[0]) for o in out]
    print("reference R=1 x3: %.1fs" % (time.time() - t0), q)
    # ---------------------------------------------------------------- RL forward, R = 16
    ref_rl.gnum_rotations = 16
    ref_rl.snum_rotations = 16
    t0 = time.time()
    with torch.no_grad():
        out = ref_rl.forward(x, m, 0, True, -1)
    q["rl_style0_R16"] = [float(o.view(-1)[0]) for o in out]
    gold["cpu_seconds_R16"] = time.time() - t0
    print("reference R=16: %.1fs" % gold["cpu_seconds_R16"])
    with torch.no_grad():
        q["rl_style1_R16_rot5"] = [float(ref_rl.forward(x, m, 1, True, 5).view(-1)[0])]
        q["rl_style2_R16_rot5"] = [float(ref_rl.forward(x, m2, 2, True, 5).view(-1)[0])]
    ref_rl.gnum_rotations = 1
    ref_rl.snum_rotations = 1
    gold["q"] = q

    # oracle check
    with torch.no_grad():
        for style in (0, 1, 2):
            o = qnet.model_forward(sd, x, m2 if style == 2 else m, style, True, -1)
            assert abs(float(o[0].view(-1)[0]) - q["rl_style%d_R1" % style][0]) < 2e-5, (style, o, q)
        o = qnet.model_forward(sd, x, m, 1, True, 5, gnum_rotations=16, snum_rotations=16)
        assert abs(float(o.view(-1)[0]) - q["rl_style1_R16_rot5"][0]) < 2e-5
    print("oracle Q ok")

    # ---------------------------------------------------------------- reactive forward
    torch.manual_seed(0)
    ref_re = models.reactive_net(True)
    ref_re.train()
    with torch.no_grad():
        out = ref_re.forward(x, m, 0, True, -1)
        gold["reactive_style0_R1"] = [float(v) for v in out[0].view(-1)]
        out = ref_re.forward(x, m2, 2, True, -1)
        gold["reactive_style2_R1"] = [float(v) for v in out[0].view(-1)]
    print("reactive", gold["reactive_style0_R1"])

    # ---------------------------------------------------------------- Trainer.forward / backprop (patched literals)
    tp = patched_trainer_module()
    torch.manual_seed(0)
    tr = tp.Trainer("reinforcement", 0.5, False, None, False)
    tr.use_cuda = True
    tr.model.use_cuda = True
    tr.model_target.use_cuda = True
    pred = tr.forward(scene_hm, mask_hm, style=0, is_volatile=True, is_target=False)
    gold["trainer_forward_rl_style0"] = [float(v) for v in pred]
    assert abs(pred[0] - q["rl_style0_R1"][0]) < 1e-6
    masks = sc["masks"].astype(np.float64).copy()
    before = {k: v.detach().clone() for k, v in tr.model.state_dict().items()}
    t0 = time.time()
    loss = tr.backprop(scene_hm, "grasp", [0, 0], [0, 0], [], [], 1.0, masks.copy(), [0] * 4, [0] * 4, [])
    gold["backprop_rl_grasp"] = {"label": 1.0, "loss": float(loss), "seconds": time.time() - t0}
    grads = {n: p.grad for n, p in tr.model.named_parameters() if p.grad is not None}
    gold["backprop_rl_grasp"]["n_grads"] = len(grads)
    sel = ["grasp_depth_trunk.features.conv0.weight", "grasp_depth_trunk.features.norm0.weight",
           "grasp_depth_trunk.features.denseblock1.denselayer1.conv1.weight",
           "grasp_depth_trunk.features.denseblock1.denselayer6.conv2.weight",
           "grasp_depth_trunk.features.transition1.conv.weight",
           "grasp_depth_trunk.features.denseblock3.denselayer24.norm1.bias",
           "grasp_depth_trunk.features.denseblock4.denselayer16.conv2.weight",
           "grasp_depth_trunk.features.norm5.weight",
           "graspnet_val.grasp-val-norm0.weight", "graspnet_val.grasp-val-conv0.weight",
           "graspnet_val.grasp-val-norm1.bias", "graspnet_val.grasp-val-conv1.weight"]
    gold["backprop_rl_grasp"]["grads"] = {k: fingerprint(grads[k]) for k in sel}
    after = tr.model.state_dict()
    gold["backprop_rl_grasp"]["param_delta"] = {
        k: fingerprint(after[k] - before[k]) for k in sel}
    gold["backprop_rl_grasp"]["bn_running_mean_after"] = fingerprint(
        after["grasp_depth_trunk.features.denseblock2.denselayer3.norm1.running_mean"])
    gold["backprop_rl_grasp"]["bn_running_var_after"] = fingerprint(
        after["grasp_depth_trunk.features.denseblock2.denselayer3.norm1.running_var"])
    print("backprop loss", loss, "grads", len(grads))

    torch.manual_seed(0)
    tre = tp.Trainer("reactive", 0.5, False, None, False)
    tre.use_cuda = True
    tre.model.use_cuda = True
    pre = tre.forward(scene_hm, mask_hm, style=1, is_volatile=True, is_target=False)
    gold["trainer_forward_reactive_style1"] = float(pre)
    loss = tre.backprop(scene_hm, "suction", [0, 0], [0, 0], [], [], 1, masks.copy(), [0] * 4, [0] * 4, [])
    gre = {n: p.grad for n, p in tre.model.named_parameters() if p.grad is not None}
    gold["backprop_reactive_suction"] = {
        "label": 1, "loss": float(loss), "n_grads": len(gre),
        "grads": {k: fingerprint(gre[k]) for k in ("suction_depth_trunk.features.conv0.weight",
                                                   "suctionnet_val.suction-val-conv1.weight")}}
    print("reactive backprop loss", loss)

    # ---------------------------------------------------------------- heightmap
    utils = mods["utils"]
    cam = synth.make_camera(3)
    out = utils.get_heightmap(cam["color"].copy(), cam["depth"].copy(), cam["intrinsics"], cam["pose"],
                              synth.WORKSPACE_LIMITS, 0.002)
    np.savez_compressed(os.path.join(OUT, "heightmap_seed3.npz"), depth224=out[1], A_htor=out[4],
                        depth448_rows=out[3][::7])
    gold["heightmap"] = {"camera_seed": 3, "depth_sha": sha(cam["depth"]), "depth224_sha": sha(out[1]),
                         "depth448_sha": sha(out[3]), "color224_sha": sha(out[0]), "color448_sha": sha(out[2])}

    # ---------------------------------------------------------------- NMS
    nms = mods["NMS"]
    cases = []
    known = np.array([[[10, 10], [60, 60]], [[12, 12], [62, 62]], [[100, 100], [160, 150]], [[0, 0], [5, 5]],
                      [[0, 0], [200, 200]]], np.float32)
    cases.append({"kind": "known5", "keep": [int(v) for v in nms.py_cpu_nms(known, np.ones(5), 0.40, 224 * 224 / 60, 224 * 224 / 5)]})
    for seed, n in ((0, 100), (1, 100), (2, 37), (3, 1), (4, 0)):
        boxes, scores = synth.make_boxes(seed, n) if n else (np.zeros((0, 2, 2), np.float32), np.zeros(0, np.float32))
        keep = nms.py_cpu_nms(boxes, scores, 0.40, 224 * 224 / 60, 224 * 224 / 5)
        cases.append({"kind": "random", "seed": seed, "n": n, "keep": [int(v) for v in keep]})
    gold["nms"] = cases

    with open(os.path.join(OUT, "golden.json"), "w") as fjs:
        json.dump(gold, fjs, indent=1)
    print("wrote", os.path.join(OUT, "golden.json"))


if __name__ == "__main__":
    main()
