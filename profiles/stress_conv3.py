import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.nn.functional as F
from smg_b200 import engine
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
eng = engine.Engine(0, 70, 640, "fp32")
for (n, hin) in [(17, 160), (40, 80), (68, 40), (68, 20)]:
    for rep in range(3):
        g = torch.Generator(device="cuda").manual_seed(n * 1000 + hin + rep)
        x = torch.randn((n, hin, hin, 128), generator=g, device="cuda")
        scale = torch.rand((n, 128), generator=g, device="cuda") + 0.5
        shift = torch.randn((n, 128), generator=g, device="cuda") * 0.3
        w = torch.randn((32, 128, 3, 3), generator=g, device="cuda") / (128 * 9) ** 0.5
        out, stats = eng.debug_conv("tf32", x, 128, scale, shift, True, 0, w, 96, 32)
        a = torch.relu(x * scale[:, None, None, :] + shift[:, None, None, :]).permute(0, 3, 1, 2)
        ref = F.conv2d(a, w, padding=1).permute(0, 2, 3, 1)
        got = out[..., 32:64]
        d = (got - ref).abs()
        err = float(d.max() / ref.abs().max())
        bad = (d > 5e-3 * ref.abs().max()).nonzero()
        print("n=%d hin=%d rep=%d err=%.2e bad=%d" % (n, hin, rep, err, bad.shape[0]), bad[:6].tolist() if bad.shape[0] else "")
        torch.cuda.synchronize()
