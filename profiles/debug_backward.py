"""Debug helper: per-tensor gradient error of the CUDA backward vs the CPU oracle, in backward order."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from oracle import qnet
import smg_b200.models as models, smg_b200.synth as synth

MEAN, STD = 0.01, 0.03
sc = synth.make_scene(1, num_objects=4, cluttered=False)
scene = sc["scene"]; mask = synth.masked_scene(scene, sc["masks"], [0])
x, m = qnet.preprocess(scene, MEAN, STD), qnet.preprocess(mask, MEAN, STD)
torch.manual_seed(0)
net = models.reinforcement_net(True)
sd = {k: v.clone() for k, v in net.state_dict().items()}
net = net.cuda(); net.train()
out = net.forward(x, m, 0, False, 0)
d = net.gra_prob[0, 0, 0, 0] - 1.0
loss = 0.5 * (d ** 2) if abs(float(d)) < 1 else abs(d) - 0.5
loss.sum().backward()
ref_loss, ref = qnet.backprop_grads(sd, x, m, 0, 0, 1.0, "reinforcement")
print("loss", float(loss), ref_loss)
names = [n for n, p in net.named_parameters() if p.grad is not None]
rows = []
for n in names:
    g = dict(net.named_parameters())[n].grad.detach().cpu().double(); r = ref[n].double()
    rows.append((n, float((g - r).abs().max() / r.abs().max().clamp_min(1e-30)), float(r.abs().max()), float(g.abs().max())))
for n, e, rm, gm in reversed(rows):
    flag = "  <<<<" if e > 1e-3 else ""
    print("%-75s err %.2e  ref %.2e got %.2e%s" % (n, e, rm, gm, flag))

print("---- per-channel detail")
params = dict(net.named_parameters())
for n in ("grasp_depth_trunk.features.transition3.norm.bias", "grasp_depth_trunk.features.denseblock3.denselayer24.norm1.bias",
          "grasp_depth_trunk.features.denseblock3.denselayer24.norm1.weight"):
    g = params[n].grad.detach().cpu().double(); r = ref[n].double()
    d = (g - r).abs()
    idx = torch.argsort(d, descending=True)[:8]
    print(n, "n_ch", g.numel(), "mean abs ref %.3e" % r.abs().mean(), "mean abs err %.3e" % d.mean())
    for i in idx.tolist():
        print("   ch %4d got % .6e ref % .6e diff % .3e" % (i, g[i], r[i], g[i] - r[i]))
