"""Is conv3_wt bound by DRAM or by the SM-side pipeline?  Time the block-1 3x3 layer for growing sample counts: the first
few fit the 126 MB L2 (13 MB of input per sample, launch repeated on the same data), the large ones stream from HBM.
Prints us per launch and the marginal us per sample."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from smg_b200 import engine  # noqa: E402

eng = engine.Engine(0, 70, 640, "fp32")
hin = int(sys.argv[1]) if len(sys.argv) > 1 else 160
prev = None
for n in ([int(v) for v in sys.argv[2].split(',')] if len(sys.argv) > 2 else (1, 2, 3, 4, 6, 8, 17, 34, 68)):
    g = torch.Generator(device="cuda").manual_seed(n)
    x = torch.randn((n, hin, hin, 128), generator=g, device="cuda")
    scale = torch.rand((n, 128), generator=g, device="cuda") + 0.5
    shift = torch.randn((n, 128), generator=g, device="cuda") * 0.3
    w = torch.randn((32, 128, 3, 3), generator=g, device="cuda") / (128 * 9) ** 0.5
    for _ in range(3):
        eng.debug_conv("tf32", x, 128, scale, shift, True, 0, w, 256, 64)
    torch.cuda.synchronize()
    eng.profile_enable(True)
    reps = 10
    for _ in range(reps):
        eng.debug_conv("tf32", x, 128, scale, shift, True, 0, w, 256, 64)
    prof = eng.profile_read()
    eng.profile_enable(False)
    us = prof["conv3x3"]["ms"] * 1e3 / reps               # class 2 = 3x3 convolutions, CUDA events around the launch itself
    note = "" if prev is None else "  marginal %.2f us/sample" % ((us - prev[1]) / (n - prev[0]))
    print("hin=%d n=%2d: %.1f us per launch%s" % (hin, n, us, note), flush=True)
    prev = (n, us)
