"""Does splitting the benched 4-unit inference step over TWO handles / streams (2 units each, chains interleaving on the GPU)
beat one handle with 4 units?  Device-timed, graph replay, same inputs."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from smg_b200 import engine as _engine  # noqa: E402
from smg_b200.trainer import Trainer  # noqa: E402


def main():
    torch.manual_seed(0)
    tr = Trainer("reinforcement", 0.5, False, None, False, precision="tf32")
    tr.model.gnum_rotations = tr.model.snum_rotations = bench.R
    m = tr.model
    G = 4
    scenes, masks = bench.make_units(G, 100)
    sd, md = torch.from_numpy(scenes).cuda(), torch.from_numpy(masks).cuda()
    rots = list(range(bench.R))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main_s = torch.cuda.current_stream()

    def run(nlanes, steps=30):
        per = G // nlanes
        engs = [_engine.get_engine(0, per * (bench.R + 1), 640, "tf32", owner=("probe", nlanes, k)) for k in range(nlanes)]
        streams = [main_s] + [torch.cuda.Stream() for _ in range(nlanes - 1)]
        for e in engs:
            e.sync_weights(m, style=0)

        def step():
            outs = []
            for k, (e, s) in enumerate(zip(engs, streams)):
                s.wait_stream(main_s)
                with torch.cuda.stream(s):
                    outs.append(e.qforward_maps_batch(0, sd[k * per:(k + 1) * per], md[k * per:(k + 1) * per, None], bench.MEAN, bench.STD, rots, bench.R))
            for s in streams[1:]:
                main_s.wait_stream(s)
            return torch.cat(outs)

        for _ in range(4):
            q = step()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            q = step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        return ms, q

    ms1, q1 = run(1)
    ms2, q2 = run(2)
    ms4, q4 = run(4)
    ms1b, _ = run(1)
    print("1 handle x 4 units: %.3f ms (%.1f U/s); 2 x 2: %.3f ms (%.1f U/s); 4 x 1: %.3f ms (%.1f U/s); 1 x 4 again: %.3f ms" %
          (ms1, G / ms1 * 1e3, ms2, G / ms2 * 1e3, ms4, G / ms4 * 1e3, ms1b))
    print("max |dq| 2 lanes vs 1: %.2e, 4 lanes vs 1: %.2e" % (float((q2 - q1).abs().max()), float((q4 - q1).abs().max())))


if __name__ == "__main__":
    main()
