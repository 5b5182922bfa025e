"""Repeat the three densenet transitions through trans_t.cu and compare every run with the first (outputs are deterministic:
any difference is a race).  Debugging aid."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from smg_b200 import engine  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
eng = engine.Engine(0, 70, 640, "fp32")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
for (n, hin, cin) in [(17, 160, 256), (17, 80, 512), (17, 40, 1024), (68, 40, 1024), (68, 80, 512), (2, 40, 1024), (2, 160, 256)]:
    g = torch.Generator(device="cuda").manual_seed(n * 1000 + hin)
    x = torch.randn((n, hin, hin, cin), generator=g, device="cuda")
    scale = torch.rand((n, cin), generator=g, device="cuda") + 0.5
    shift = torch.randn((n, cin), generator=g, device="cuda") * 0.3
    w = torch.randn((cin // 2, cin, 1, 1), generator=g, device="cuda") / cin ** 0.5
    first = None
    nbad = 0
    for rep in range(reps):
        out, stats = eng.debug_conv("tf32", x, cin, scale, shift, True, 1, w)
        torch.cuda.synchronize()
        if first is None:
            first = out.clone()
            a = torch.relu(x * scale[:, None, None, :] + shift[:, None, None, :]).permute(0, 3, 1, 2)
            ref = F.conv2d(F.avg_pool2d(a, 2, 2), w).permute(0, 2, 3, 1)
            print("n=%d hin=%d cin=%d err vs torch %.2e" % (n, hin, cin, float((out - ref).abs().max() / ref.abs().max())), flush=True)
        else:
            d = (out != first)
            if bool(d.any()):
                nbad += 1
                idx = d.nonzero()
                print("  rep %d: %d elements differ, first %s, max diff %.3e" % (rep, idx.shape[0], idx[0].tolist(),
                                                                               float((out - first).abs().max())), flush=True)
    print("  %d of %d repeats differ" % (nbad, reps - 1), flush=True)
