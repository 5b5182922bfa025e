"""Debug helper: per-tensor / per-channel gradient error of the CUDA backward vs the CPU oracle on the kink-free network
(all BatchNorm biases = +6, see tests/test_gpu_parity_r02.py), in backward order."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from oracle import qnet
import smg_b200.models as models, smg_b200.synth as synth

precision = sys.argv[1] if len(sys.argv) > 1 else "fp32"
MEAN, STD = 0.01, 0.03
sc = synth.make_scene(1, num_objects=4, cluttered=False)
scene = sc["scene"]; mask = synth.masked_scene(scene, sc["masks"], [0])
x, m = qnet.preprocess(scene, MEAN, STD), qnet.preprocess(mask, MEAN, STD)
torch.manual_seed(0)
net = models.reinforcement_net(True)
sd = {k: v.clone() for k, v in net.state_dict().items()}
for k in sd:
    if "norm" in k and k.endswith(".bias"):
        sd[k] = torch.full_like(sd[k], 6.0)
net.load_state_dict(sd)
net = net.cuda(); net.train(); net.precision = precision
net.gnum_rotations = net.snum_rotations = 16
out = net.forward(x, m, 0, False, 3)
d = net.gra_prob[0, 0, 0, 0] - 1.0
loss = 0.5 * (d ** 2) if abs(float(d)) < 1 else abs(d) - 0.5
loss.sum().backward()
ref_loss, ref = qnet.backprop_grads(sd, x, m, 0, 3, 1.0, "reinforcement", gnum_rotations=16)
print("loss", float(loss), ref_loss)
params = dict(net.named_parameters())
names = [n for n, p in params.items() if p.grad is not None]
for n in reversed(names):
    g = params[n].grad.detach().cpu().double(); r = ref[n].double()
    e = float((g - r).abs().max() / r.abs().max().clamp_min(1e-30))
    if e > 3e-4:
        dd = (g - r).abs().flatten()
        idx = torch.argsort(dd, descending=True)[:4].tolist()
        print("%-72s err %.2e ref %.2e | worst idx %s diffs %s" % (n.replace("grasp_depth_trunk.features.", ""), e, float(r.abs().max()),
              idx, ["%.2e" % float(dd[i]) for i in idx]))
