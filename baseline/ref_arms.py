"""The reference arms of bench.py: the UNMODIFIED reference modules timed on the box.

  * CPU arm (`bench.py --impl reference`, `cpu_baseline`): code/models.py + code/trainer.py on the host cores through
    oracle/refshim.py (matplotlib/apex stubs, densenet121(weights=None), `.cuda()` -> identity).
  * GPU arm (`bench.py --impl reference-gpu`, `gpu_reference` inside the normal line): the same modules on `torch.cuda`
    exactly as written (stock PyTorch + cuDNN) - BASELINE.md section 4 calls this "the real bar to beat".

The modules are imported from /root/reference/code in the build container and from the build-time copy
baseline/_ref/code (git-ignored, shipped with the snapshot) on the GPU box.  `trainer.py` is compiled in memory with its
two NaN-producing literals replaced by mean 0.01 / std 0.03 - a harness choice, stated in every line.
Nothing here is imported by the product package.
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

R = 16
MEAN, STD = 0.01, 0.03


def available():
    from oracle import refshim
    return refshim.available()


def _inputs(seed=100):
    import numpy as np
    import smg_b200.synth as synth
    sc = synth.make_scene(seed, num_objects=4, cluttered=False)
    return sc["scene"], synth.masked_scene(sc["scene"], sc["masks"], [0]), sc["masks"].astype(np.float64)


def _trainer(cpu):
    import torch
    from oracle import refshim
    refshim.install(cpu=cpu)
    tp = refshim.patched_trainer_module(MEAN, STD)
    torch.manual_seed(0)
    tr = tp.Trainer("reinforcement", 0.5, False, None, False)
    if cpu:
        # the nets' forward is CUDA-only as written (code/models.py:377-385); with `.cuda()` shimmed to identity it runs on
        # the host when use_cuda is forced on
        tr.use_cuda = tr.model.use_cuda = tr.model_target.use_cuda = True
    tr.model.gnum_rotations = tr.model.snum_rotations = R
    return tr


# ------------------------------------------------------------------------------------------------
# CPU arm
# ------------------------------------------------------------------------------------------------
_CPU = {}


def cpu_rotations_seconds(n_rot, repeats=1):
    """Seconds for `n_rot` of the 16 rotations of one unit, executed by the reference itself: per rotation
    `model.forward(x, m, 0, True, r)` = rotate + trunk(scene) + trunk(mask) + head (code/models.py:444-465)."""
    import torch
    from oracle import qnet
    torch.set_num_threads(os.cpu_count())
    if "tr" not in _CPU:
        _CPU["tr"] = _trainer(cpu=True)
        scene, mask, _ = _inputs()
        _CPU["x"], _CPU["m"] = qnet.preprocess(scene, MEAN, STD), qnet.preprocess(mask, MEAN, STD)
    tr, x, m = _CPU["tr"], _CPU["x"], _CPU["m"]
    from oracle import refshim
    refshim.set_cpu_mode(True)
    best = None
    try:
        for _ in range(repeats):
            t0 = time.perf_counter()
            for r in range(n_rot):
                tr.model.forward(x, m, 0, True, r)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    finally:
        refshim.set_cpu_mode(False)   # leave torch as we found it: the GPU arms share the process
    return best


# ------------------------------------------------------------------------------------------------
# GPU arm: stock PyTorch + cuDNN on the same B200
# ------------------------------------------------------------------------------------------------
def _time_gpu(fn, steps, warmup):
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 1e3 / steps


def gpu_reference(steps=3, warmup=1, backprop_steps=5):
    """units/s and steps/s of the unmodified reference on torch.cuda, TF32 off (the fp32 parity oracle) and on (torch's
    default for cuDNN convolutions).  Per precision:
      forward_e2e       Trainer.forward(scene, mask, 0, is_volatile=True) with R = 16: host pre-processing, H2D, 32 trunk
                        passes, one D2H per rotation - the call the reference's step loop makes (code/main.py:165)
      forward_dedup17   the same arithmetic scheduled the way this repo schedules it (16 rotated scenes + ONE masked pass),
                        still stock PyTorch modules of the reference's net
      backprop          Trainer.backprop (code/trainer.py:278-384): grad-enabled forward, backward, Adam
    """
    import numpy as np
    import torch
    import torch.nn.functional as F
    from oracle import qnet
    out = {"torch": torch.__version__, "cudnn": torch.backends.cudnn.version(), "rotations": R,
           "image_mean": MEAN, "image_std": STD}
    tr = _trainer(cpu=False)
    scene, mask, obj_masks = _inputs()
    net = tr.model
    x = qnet.preprocess(scene, MEAN, STD)
    m = qnet.preprocess(mask, MEAN, STD)

    def fwd_e2e():
        return tr.forward(scene, mask, 0, True, False)

    def fwd_dedup():
        with torch.no_grad():
            xs, ms = x.cuda(), m.cuda()
            f_m = net.grasp_depth_trunk.features(ms)
            q = []
            for r in range(R):
                th = np.radians(r * (360 / R))
                aff = np.asarray([[np.cos(-th), np.sin(-th), 0], [-np.sin(-th), np.cos(-th), 0]])
                aff = torch.from_numpy(aff).float().view(1, 2, 3).cuda()
                grid = F.affine_grid(aff, xs.size(), align_corners=True)
                f_s = net.grasp_depth_trunk.features(F.grid_sample(xs, grid, mode="nearest", align_corners=True))
                q.append(net.graspnet_val(torch.cat((f_s, f_m), dim=1)))
            return torch.cat(q).cpu()

    i = [0]

    def bwd():
        i[0] += 1
        return tr.backprop(scene, "grasp", [i[0] % 4, i[0] % R], [0, 0], [], [], 1.0, obj_masks.copy(), [0] * 4, [0] * 4, [])

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    out["q_unit0"] = [float(v) for v in fwd_e2e()]       # the fp32 parity oracle's own answer on bench.py's first unit
    for name, tf32 in (("fp32", False), ("tf32", True)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        t_e2e = _time_gpu(fwd_e2e, steps, warmup)
        t_dd = _time_gpu(fwd_dedup, steps, warmup)
        t_bp = _time_gpu(bwd, backprop_steps, 2)
        out[name] = {"forward_e2e_units_per_s": 1.0 / t_e2e, "forward_dedup17_units_per_s": 1.0 / t_dd,
                     "backprop_steps_per_s": 1.0 / t_bp, "ms_per_unit_e2e": 1e3 * t_e2e, "ms_per_unit_dedup17": 1e3 * t_dd,
                     "ms_per_backprop_step": 1e3 * t_bp}
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return out
