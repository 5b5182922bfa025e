"""One warmed-up unit of work (R=16 Q pass) inside a cudaProfilerStart/Stop range, for ncu:

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python profiles/profile_step.py --precision tf32
  ncu --profile-from-start off --set full --clock-control none --import-source on \
      -k regex:conv_umma -c 6 -o gpurun_out/prof_conv python profiles/profile_step.py --precision tf32
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
from smg_b200.trainer import Trainer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="tf32")
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--mode", default="infer", choices=["infer", "train"])
    ap.add_argument("--units", type=int, default=1, help="units per step (bench.py's default step has 4)")
    args = ap.parse_args()
    if args.mode == "train":
        return train(args)
    torch.manual_seed(0)
    tr = Trainer("reinforcement", 0.5, False, None, False, precision=args.precision)
    tr.model.gnum_rotations = tr.model.snum_rotations = bench.R
    tr.model.update_running_stats = False
    if args.units > 1:
        return infer_batch(args, tr)
    eng = tr.model._engine(bench.R + 1)
    scenes, masks = bench.make_units(2, 100)
    sd, md = torch.from_numpy(scenes).cuda(), torch.from_numpy(masks).cuda()
    rots = list(range(bench.R))
    for i in range(3):
        eng.qforward_maps(0, sd[i % 2], md[i % 2:i % 2 + 1], bench.MEAN, bench.STD, rots, bench.R)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    for i in range(args.steps):
        q = eng.qforward_maps(0, sd[i % 2], md[i % 2:i % 2 + 1], bench.MEAN, bench.STD, rots, bench.R)
        eng.argmax(q)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


def infer_batch(args, tr):
    """bench.py's step: G units in one smg_qforward_maps_batch call (set SMG_NO_GRAPHS=1 to list every launch)."""
    G = args.units
    eng = tr.model._engine(G * (bench.R + 1))
    scenes, masks = bench.make_units(G, 100)
    sd, md = torch.from_numpy(scenes).cuda(), torch.from_numpy(masks).cuda()
    rots = list(range(bench.R))
    for i in range(3):
        eng.qforward_maps_batch(0, sd, md[:, None], bench.MEAN, bench.STD, rots, bench.R)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    for i in range(args.steps):
        q = eng.qforward_maps_batch(0, sd, md[:, None], bench.MEAN, bench.STD, rots, bench.R)
        eng.argmax(q)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


def train(args):
    """One warmed-up Trainer.backprop step inside the profiler range."""
    import numpy as np
    import smg_b200.synth as synth
    torch.manual_seed(0)
    tr = Trainer("reinforcement", 0.5, False, None, False, precision=args.precision)
    sc = synth.make_scene(100, num_objects=4, cluttered=False)
    masks = sc["masks"].astype(np.float64)
    for i in range(3):
        tr.backprop(sc["scene"], "grasp", [i % 4, 0], [0, 0], [], [], 1.0, masks.copy(), [0] * 4, [0] * 4, [])
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    for i in range(args.steps):
        tr.backprop(sc["scene"], "grasp", [i % 4, 0], [0, 0], [], [], 1.0, masks.copy(), [0] * 4, [0] * 4, [])
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


if __name__ == "__main__":
    main()
