"""How much of Trainer.backprop's time per step is host work?  Compare the public call (host heightmaps in, loss out) with
back-to-back smg_train_step launches on device-resident inputs (no host read between steps)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import smg_b200.synth as synth  # noqa: E402
from smg_b200.trainer import Trainer  # noqa: E402


def main():
    torch.manual_seed(0)
    tr = Trainer("reinforcement", 0.5, False, None, False, precision="tf32")
    sc = synth.make_scene(100, num_objects=4, cluttered=False)
    masks = sc["masks"].astype(np.float64)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def bp(i):
        return tr.backprop(sc["scene"], "grasp", [i % 4, 3], [0, 0], [], [], 1.0, masks.copy(), [0] * 4, [0] * 4, [])

    for i in range(6):
        bp(i)
    n = 40
    for stats in (True, False):
        tr.model.update_running_stats = stats
        for i in range(3):
            bp(i)
        torch.cuda.synchronize()
        e0.record()
        for i in range(n):
            bp(i)
        e1.record()
        torch.cuda.synchronize()
        print("Trainer.backprop, running stats %s: %.3f ms per step" % (stats, e0.elapsed_time(e1) / n))
    eng = tr.model._engine(2, 0)
    st = tr._fused_state(0)
    hm = torch.from_numpy(np.stack([sc["scene"], synth.masked_scene(sc["scene"], sc["masks"], [1])])).cuda()
    for full in (False, True):
        for i in range(3):
            eng.train_step(0, hm[0], hm[1], 3, 16, 0, 0.7, [1.0, 1.0, 1.0], st["ptrs"], len(st["params"]), 7 + i, grads_only=not full,
                           want_bn_stats=False)
        torch.cuda.synchronize()
        e0.record()
        for i in range(n):
            eng.train_step(0, hm[0], hm[1], 3, 16, 0, 0.7, [1.0, 1.0, 1.0], st["ptrs"], len(st["params"]), 10 + i, grads_only=not full,
                           want_bn_stats=False)
        e1.record()
        torch.cuda.synchronize()
        print("smg_train_step back to back (%s): %.3f ms per step" % ("with Adam + re-pack" if full else "gradients only", e0.elapsed_time(e1) / n))


if __name__ == "__main__":
    main()
