"""Soak test of the captured training step (debugging aid): the same sample through smg_train_step with SMG_STEP_GRADS_ONLY
(weights untouched) many times - every repeat must reproduce the first flat gradient up to the order of the float atomics
(split-K weight gradients: ~1e-6 of a tensor's scale).  Catches intermittent faults of the backward kernels and of the
two-stream schedule that single-shot parity tests cannot see."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import smg_b200.synth as synth  # noqa: E402
from smg_b200.trainer import Trainer  # noqa: E402


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    torch.manual_seed(0)
    tr = Trainer("reinforcement", 0.5, False, None, False, precision="tf32")
    sc = synth.make_scene(100, num_objects=4, cluttered=False)
    worst_all = 0.0
    for style, rot in ((0, 3), (1, 7), (2, 0)):
        eng = tr.model._engine(2, style)
        st = tr._fused_state(style)
        eng._mean_std = (tr.image_mean, tr.image_std)
        mk = synth.masked_scene(sc["scene"], sc["masks"], [style % 4])
        hm = torch.from_numpy(np.stack([sc["scene"], mk])).cuda()
        flat = st["flat"]["grad"] if "flat" in st else None
        views = st["views"]["grad"]

        def grads():
            eng.train_step(style, hm[0], hm[1], rot, tr.model.gnum_rotations, 0, 0.7, [1.0, 1.0, 1.0], st["ptrs"], len(st["params"]),
                           1, grads_only=True, want_bn_stats=False)
            torch.cuda.synchronize()
            return [v.clone() for v in views]

        for _ in range(3):
            first = grads()                                   # eager, capture, replay
        scales = [float(g.abs().max()) + 1e-30 for g in first]
        worst = 0.0
        for r in range(reps):
            g = grads()
            dev = max(float((a - b).abs().max()) / s for a, b, s in zip(g, first, scales) if s > 1e-20)
            worst = max(worst, dev)
            if dev > 1e-3:
                print("style %d repeat %d: gradient deviates by %.3e of a tensor's scale" % (style, r, dev), flush=True)
        print("style %d: %d repeats, worst per-tensor deviation %.2e" % (style, reps, worst), flush=True)
        worst_all = max(worst_all, worst)
    print("worst %.2e" % worst_all)


if __name__ == "__main__":
    main()
