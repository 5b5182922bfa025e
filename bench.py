#!/usr/bin/env python
"""bench.py - 16-rotation Q inferences/s of the SMG grasp-affordance hot path on B200.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--precision tf32|bf16|fp32]

Unit of work U: one (scene, object mask, primitive) evaluated at R = 16
rotations -> 16 Q scalars == one `Trainer.forward(..., is_volatile=True)` call of the reference with
gnum_rotations = 16: 17 distinct 640x640 DenseNet-121 trunk passes (16 rotated scenes + 1 masked
scene) + 16 heads = 788.0 GFLOP.  One "step" = `--units` (default 4) independent U per GPU evaluated as one batch
(every unit gets exactly its single-call result because BatchNorm statistics are per sample; `--units 1` is the
latency configuration) on seeded synthetic 224x224 heightmaps with random-init weights (no datasets offline).

  value   U/s with the heightmaps already resident in HBM (CUDA events around K steps)
  e2e     U/s through the reference-facing call Trainer.forward: float64 heightmaps in pinned host
          memory, H2D copy and D2H read of the 16 Q values inside the timed region
  roofline      dominant kernel class timed with CUDA events inside the library (smg_profile_*)
  cpu_baseline  the unmodified reference (baseline/_ref/code; the oracle port if that copy is absent) timed on this box's
                host cores on a bounded sample

  gpu_reference the unmodified reference on torch.cuda (stock PyTorch + cuDNN, TF32 off and on) on the same GPU, same
                unit and same Trainer.backprop call: the bar BASELINE.md section 4 names (N = 1, rank 0)
  fp32_mode     the <= 1e-4 mode of this library on the same input: units/s and its error against the tf32 result's reference

`--impl reference` times the CPU path alone: the UNMODIFIED reference modules (a build-time copy under baseline/_ref/,
imported through oracle/refshim.py; kind "reference") or, if that copy is absent, the oracle port (kind "port").
`--impl reference-gpu` prints the gpu_reference object alone.  N > 1 (torchrun): every rank
evaluates its own units (weak scaling over independent scenes - the rotations / replay samples shard
without a data-path collective); the per-rank best (Q, rotation) tuples are exchanged with one tiny
NCCL all_gather per step, which is the path's only real exchange (code/main.py:170-173).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

R = 16
MEAN, STD = 0.01, 0.03
GFLOP_PER_PASS = 46.255
GFLOP_PER_HEAD = 0.105
GFLOP_PER_UNIT = 17 * GFLOP_PER_PASS + 16 * GFLOP_PER_HEAD  # 788.0
METRIC = "16-rot Q-map inferences/s"
UNIT = "inferences/s"
WORKLOAD = ("reinforcement_net (E+S+ES heads) forward, primitive E (style 0), R=16 rotations, one synthetic "
            "224x224 heightmap + one object mask -> 16 Q; 17 distinct 640x640 DenseNet-121 passes per unit")


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def ncu_traffic(cls):
    """DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of one representative launch of a kernel class, read from
    the newest committed `ncu --set full` summary of that kernel under profiles/ (never a literal in this file), next to
    the algorithmic bytes of the same launch (block-1 layer of profiles/profile_step.py; the summary's `samples` row says how
    many samples that launch processed: 68 = the benched 4-unit step)."""
    import csv
    import glob
    stem = {"conv1x1": "conv1_t", "conv3x3": "conv3_wt"}.get(cls)
    if stem is None:
        return None
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r[0-9][0-9]_%s_ncu_summary.csv" % stem)))
    if not files:
        return None
    vals = {}
    with open(files[-1]) as f:
        for row in csv.reader(f):
            if len(row) >= 2 and row[0] == "samples":
                vals["samples"] = float(row[1])
            if len(row) >= 3 and row[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"):
                scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0, "us": 1.0, "ms": 1e3}.get(row[2], 1.0)
                vals[row[0]] = float(row[1]) * scale
    if "dram__bytes_read.sum" not in vals or "dram__bytes_write.sum" not in vals:
        return None
    samples = int(vals.get("samples", 17))      # the r01 / first r02 captures are one-unit steps (17 samples)
    px = samples * 160 * 160
    algo = {"conv1x1": px * (192 + 128) * 4.0, "conv3x3": px * (128 + 32) * 4.0}[cls]
    return {"source": os.path.relpath(files[-1], ROOT), "dram_bytes": vals["dram__bytes_read.sum"] + vals["dram__bytes_write.sum"],
            "algorithmic_bytes": algo, "launch_us_under_ncu": vals.get("gpu__time_duration.sum"),
            "launch": "block-1 %s layer, %d samples" % ("1x1 K=192" if cls == "conv1x1" else "3x3", samples)}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, smax, reasons = [], [], set()
        for line in self.f.read().strip().splitlines():
            c = [v.strip() for v in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                smax.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(smax), "reasons": sorted(reasons), "samples": len(sm)}


def make_units(n_units, seed0):
    import numpy as np
    import smg_b200.synth as synth
    scenes, masks = [], []
    for i in range(n_units):
        sc = synth.make_scene(seed0 + i, num_objects=4, cluttered=False)
        scenes.append(sc["scene"])
        masks.append(synth.masked_scene(sc["scene"], sc["masks"], [0]))
    return np.stack(scenes), np.stack(masks)


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path, on the host cores
# ------------------------------------------------------------------------------------------------
_CPU_CACHE = {}


def cpu_kind():
    from baseline import ref_arms
    return "reference" if ref_arms.available() else "port"


def cpu_reference_time(rotations_per_sample, repeats=1):
    """Seconds for `rotations_per_sample` rotations of one unit executed the way the reference does:
    per rotation: rotate, trunk(scene), trunk(mask), cat, head (code/models.py:371-389).  Runs the unmodified reference
    when its build-time copy is present, the oracle port otherwise."""
    from baseline import ref_arms
    if ref_arms.available():
        return ref_arms.cpu_rotations_seconds(rotations_per_sample, repeats)
    import torch
    from oracle import qnet
    import smg_b200.models as models
    torch.set_num_threads(os.cpu_count())
    if "sd" not in _CPU_CACHE:
        torch.manual_seed(0)
        _CPU_CACHE["sd"] = models.reinforcement_net(True).state_dict()
        scenes, masks = make_units(1, 100)
        _CPU_CACHE["x"] = qnet.preprocess(scenes[0], MEAN, STD)
        _CPU_CACHE["m"] = qnet.preprocess(masks[0], MEAN, STD)
    sd, x, m = _CPU_CACHE["sd"], _CPU_CACHE["x"], _CPU_CACHE["m"]
    best = None
    with torch.no_grad():
        for _ in range(repeats):
            t0 = time.perf_counter()
            for r in range(rotations_per_sample):
                qnet.q_forward(sd, x, m, 0, [r], R)  # one rotation per call -> mask pass recomputed each time
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    return best


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    cores = os.cpu_count()
    t_rot = cpu_reference_time(1)  # also the warm-up of the thread pool
    budget = 150.0
    rps = int(max(1, min(R, budget / max(1e-9, (args.steps + args.warmup) * t_rot))))
    for _ in range(args.warmup):
        cpu_reference_time(rps)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_time(rps)
    dt = time.perf_counter() - t0
    units = args.steps * rps / float(R)
    value = units / dt
    kind = cpu_kind()
    sample = ("%d of %d rotations per step (each: rotate + trunk(scene) + trunk(mask) + head, fp32, torch %s CPU, %s), "
              "scaled by 16/%d" % (rps, R, torch.__version__,
                                   "unmodified reference modules" if kind == "reference" else "oracle port", rps))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def run_reference_gpu_arm(args):
    """The unmodified reference on torch.cuda, alone (rank 0)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    from baseline import ref_arms
    if not ref_arms.available():
        print(json.dumps({"impl": "reference-gpu", "unavailable": "baseline/_ref/code is absent (built only where /root/reference exists)"}))
        return 0
    g = ref_arms.gpu_reference(steps=max(2, min(args.steps, 5)), warmup=1)
    v = g["fp32"]["forward_e2e_units_per_s"]
    print(json.dumps({"impl": "reference-gpu", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
                      "warmup": args.warmup, "higher_is_better": True, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": WORKLOAD}, "gpu_reference": g}))
    return 0


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_gpu_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))

    import __graft_entry__ as entry
    if rank == 0:
        entry.build()
    if world > 1:
        dist.barrier()
    from smg_b200.trainer import Trainer

    torch.manual_seed(0)
    tr = Trainer("reinforcement", 0.5, False, None, False, precision=args.precision)
    tr.model.gnum_rotations = tr.model.snum_rotations = R
    tr.model.update_running_stats = False  # snapshot-only side effect; not part of the Q result
    eng = tr.model._engine(args.units * (R + 1))

    # precision guard: the timed mode must agree with the fp32 mode on the bench input (tolerance 1e-2)
    scenes, masks = make_units(max(2 * args.units, 8), 100 + 1000 * rank)
    q_fast = tr.forward(scenes[0], masks[0], 0, True, False)
    precision = args.precision
    note = None
    if precision != "fp32":
        tr.model.precision = "fp32"
        q_ref = tr.forward(scenes[0], masks[0], 0, True, False)
        err = float(np.abs(q_fast - q_ref).max() / np.abs(q_ref).max())
        if not np.isfinite(err) or err > (1e-2 if precision == "tf32" else 1e-1):
            note = "%s mode disagreed with fp32 mode (err %.3g); benchmarking fp32 mode instead" % (precision, err)
            precision = "fp32"
        else:
            note = "%s vs fp32 mode on the bench input: max|dQ|/max|Q| = %.2e" % (precision, err)
        tr.model.precision = precision
        eng = tr.model._engine(args.units * (R + 1))

    rots = list(range(R))
    scenes_d = torch.from_numpy(scenes).to(dev)
    masks_d = torch.from_numpy(masks).to(dev)
    nu = scenes.shape[0]
    best = torch.zeros(2, device=dev)
    gathered = [torch.zeros(2, device=dev) for _ in range(world)] if world > 1 else None

    U = args.units

    def unit_slice(i):
        j = (i * U) % nu
        return slice(j, j + U) if j + U <= nu else slice(0, U)

    def step_resident(i, exchange=True):
        sl = unit_slice(i)
        if U == 1:
            q = eng.qforward_maps(0, scenes_d[sl][0], masks_d[sl], MEAN, STD, rots, R)
        else:
            q = eng.qforward_maps_batch(0, scenes_d[sl], masks_d[sl][:, None], MEAN, STD, rots, R)
        val, idx = eng.argmax(q)
        if world > 1 and exchange:  # per-GPU best (Q, rotation) tuples: the path's only exchange
            best[0] = val[0]
            best[1] = idx[0].float()
            dist.all_gather(gathered, best)
        return q

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: inputs resident in HBM
    for i in range(args.warmup):
        step_resident(i)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        step_resident(args.warmup + i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = eng.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * args.steps * U / (ms_max / 1e3)

    # ---- e2e: Trainer.forward with host heightmaps (pinned), H2D + D2H inside the timed region
    pin_s = torch.from_numpy(scenes).pin_memory()
    pin_m = torch.from_numpy(masks).pin_memory()

    def step_e2e(i):
        sl = unit_slice(i)
        if U == 1:
            return tr.forward(pin_s[sl][0].numpy(), pin_m[sl][0].numpy(), 0, True, False)
        return tr.forward_batch(pin_s[sl].numpy(), pin_m[sl].numpy(), 0)

    for i in range(args.warmup):
        step_e2e(i)
    barrier()
    e0.record()
    for i in range(args.steps):
        qh = step_e2e(args.warmup + i)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * args.steps * U / (float(t.item()) / 1e3)
    h2d = int(U * (scenes[0].nbytes + masks[0].nbytes))
    d2h = int(qh.size * 4)

    # ---- the same calls with the BatchNorm running-statistics side effect reproduced (the reference updates
    # running_mean / running_var on every forward; result-neutral, snapshot-only - reported so that the work is on record)
    e2e_stats = None
    if not args.no_extras:
        tr.model.update_running_stats = True
        ns = max(3, args.steps // 4)
        for i in range(2):
            step_e2e(i)
        barrier()
        e0.record()
        for i in range(ns):
            step_e2e(args.warmup + i)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_stats = world * ns * U / (float(t.item()) / 1e3)
        tr.model.update_running_stats = False

    line_extra = {}
    # ---- roofline of the dominant kernel class (rank 0, separate short pass with event pairs per launch)
    eng = tr.model._engine(R + 1)
    if rank == 0:
        peaks = measured_peaks()
        eng.profile_enable(True)
        nprof = min(3, args.steps)
        for i in range(nprof):
            step_resident(i, exchange=False)   # rank 0 only: no collective inside the profiled pass
        prof = eng.profile_read()
        eng.profile_enable(False)
        total_ms = sum(v["ms"] for v in prof.values())
        dom = max(("conv1x1", "conv3x3", "stem"), key=lambda k: prof[k]["ms"])
        d = prof[dom]
        per_launch_ms = d["ms"] / max(1, d["launches"])
        # bound by arithmetic intensity against the measured ridge: fp32-stored activations make the DenseNet convolutions
        # memory-bound (block-1 1x1, K=224: 41 FLOP/B; 3x3: 115 FLOP/B; ridge 209 FLOP/B at the bf16 peak, 105 at tf32 rate)
        tensor_peak = peaks["bf16_tflops_sustained"] * (0.5 if precision == "tf32" else 1.0)
        ridge = tensor_peak * 1e12 / (peaks["hbm_gbs"] * 1e9)
        ai = d["flops"] / max(1.0, d["bytes"])
        traffic = ncu_traffic(dom)
        if dom == "stem" or ai < ridge:
            achieved = d["bytes"] / 1e9 / (d["ms"] / 1e3)
            roof = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": achieved / peaks["hbm_gbs"],
                    "traffic": traffic["dram_bytes"] if traffic else None, "traffic_detail": traffic,
                    "arithmetic_intensity_flop_per_byte": ai, "ridge_flop_per_byte": ridge,
                    "tensor_tflops": d["flops"] / 1e12 / (d["ms"] / 1e3)}
        else:
            achieved = d["flops"] / 1e12 / (d["ms"] / 1e3)
            roof = {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                    "frac": achieved / peaks["bf16_tflops_sustained"],
                    "traffic": traffic["dram_bytes"] if traffic else None, "traffic_detail": traffic}
        roof.update({"kernel": dom, "avg_launch_ms": per_launch_ms, "launches_per_step": d["launches"] // nprof,
                     "share_of_step": d["ms"] / total_ms if total_ms else None, "peak_source": peaks["source"],
                     "note": "peak = dense bf16 cuBLAS sustained; tf32 operands run at half the bf16 tensor rate" if precision == "tf32" else None,
                     "classes": {k: {"ms_per_step": v["ms"] / nprof, "launches_per_step": v["launches"] // nprof,
                                     "tflops": (v["flops"] / 1e12 / (v["ms"] / 1e3)) if v["ms"] else None,
                                     "algo_gbs": (v["bytes"] / 1e9 / (v["ms"] / 1e3)) if v["ms"] else None}
                                 for k, v in prof.items()}})
        line_extra["roofline"] = roof
        line_extra["whole_step_tflops"] = GFLOP_PER_UNIT * value / 1e3 / world

        # ---- CPU baseline on a bounded sample (rank 0, N = 1 only)
        if world == 1 and not args.no_cpu_baseline:
            rps = 2
            dt = cpu_reference_time(rps)
            line_extra["cpu_baseline"] = {
                "value": (rps / float(R)) / dt, "unit": UNIT, "cores": os.cpu_count(), "kind": cpu_kind(),
                "sample": "%d of %d rotations of one unit (each: rotate + trunk(scene) + trunk(mask) + head, fp32 torch CPU, "
                          "all host threads), scaled by 16/%d; %.1f s of CPU work" % (rps, R, rps, dt)}

        # ---- the unmodified reference on torch.cuda, same GPU (rank 0, N = 1 only): the real bar
        ref_q = None
        if world == 1 and not args.no_gpu_reference:
            try:
                from baseline import ref_arms
                if ref_arms.available():
                    g = ref_arms.gpu_reference(steps=3, warmup=1)
                    ref_q = g.pop("q_unit0", None)
                    line_extra["gpu_reference"] = g
                else:
                    line_extra["gpu_reference"] = {"unavailable": "baseline/_ref/code is absent"}
            except Exception as exc:
                line_extra["gpu_reference"] = {"error": repr(exc)[:300]}

        # ---- fp32 mode (north_star: <= 1e-4 of scale) on the same input, and both modes against the reference's own Q
        if world == 1 and not args.no_extras:
            try:
                tr.model.precision = "fp32"
                eng32 = tr.model._engine(U * (R + 1))
                def step32(i):
                    sl = unit_slice(i)
                    return eng32.qforward_maps_batch(0, scenes_d[sl], masks_d[sl][:, None], MEAN, STD, rots, R)
                for i in range(2):
                    step32(i)
                torch.cuda.synchronize()
                n32 = 3
                e0.record()
                for i in range(n32):
                    step32(i)
                e1.record()
                torch.cuda.synchronize()
                q32 = tr.forward(scenes[0], masks[0], 0, True, False)
                fp32 = {"value": n32 * U / (e0.elapsed_time(e1) / 1e3), "unit": UNIT, "steps": n32}
                if ref_q is not None:
                    refq = np.asarray(ref_q, np.float64)
                    fp32["err"] = float(np.abs(q32 - refq).max() / np.abs(refq).max())
                    fp32["err_what"] = "max|dQ|/max|Q_ref| over the 16 rotations of unit 0 vs the unmodified reference on torch.cuda, TF32 off"
                    line_extra["precision_err_vs_reference"] = float(np.abs(q_fast - refq).max() / np.abs(refq).max())
                line_extra["fp32_mode"] = fp32
            except Exception as exc:
                line_extra["fp32_mode"] = {"error": repr(exc)[:300]}
            tr.model.precision = precision

    # the modes below are ONE job spread over the ranks: every rank must hold the same weights (they do - same seed, nothing
    # trained yet - but a broadcast from rank 0 makes it a property of the bench, not of the order of its sections)
    if world > 1 and not args.no_extras:
        with torch.no_grad():
            for t_ in tr.model.state_dict().values():
                dist.broadcast(t_, 0)
        eng.sync_weights(tr.model, force=True)

    # ---- one highly-cluttered decision (K = 10 objects, R = 16, E + S + ES: 98 distinct trunk passes) STRONG-scaled over the
    # N GPUs (BASELINE config 5): per-rank share of the passes, all-gather of the head partials, argmax
    if not args.no_extras:
        try:
            import smg_b200.synth as synth
            from smg_b200 import decision as _decision, parallel as _parallel
            tr.model.precision = precision
            hc = synth.make_scene(7, num_objects=10, cluttered=True)
            nd = 5
            for i in range(2):
                dec = _parallel.decide_sharded(tr, hc["depth"], hc["masks"], is_ets=True)
            barrier()
            e0.record()
            for i in range(nd):
                dec = _parallel.decide_sharded(tr, hc["depth"], hc["masks"], is_ets=True)
            e1.record()
            barrier()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            line_extra["decision"] = {
                "latency_ms": float(t.item()) / nd, "decisions_per_s": nd / (float(t.item()) / 1e3), "n_gpus": world,
                "scaling": "strong", "distinct_trunk_passes": 98, "objects": 10, "rotations": R, "primitive": dec["primitive"],
                "exchange": "all_gather of the per-sample head partials (%d bytes received per rank) + all_gather of the per-rank "
                            "best (Q, index) tuples" % dec["exchange_bytes"],
                "what": "host heightmap + 10 masks in -> gra/suc/gs Q tables, argmax and primitive choice out (parallel.decide_sharded)"}
            if world == 1:
                for i in range(2):
                    _decision.decide(tr, hc["depth"], hc["masks"], is_ets=True)
                torch.cuda.synchronize()
                e0.record()
                for i in range(nd):
                    _decision.decide(tr, hc["depth"], hc["masks"], is_ets=True)
                e1.record()
                torch.cuda.synchronize()
                line_extra["decision"]["single_call_path_ms"] = e0.elapsed_time(e1) / nd
        except Exception as exc:
            line_extra["decision"] = {"error": repr(exc)[:300]}
            if world > 1:
                raise

    # ---- data-parallel replay step (BASELINE config 4): 64 highly-cluttered samples, 64 / N per GPU, gradients all-reduced
    # over NCCL, one Adam step; checked against the serial single-GPU step with --verify
    if not args.no_extras and not args.no_backprop:
        try:
            import smg_b200.synth as synth
            B = 64
            lo, hi = B * rank // world, B * (rank + 1) // world
            batch = []
            for i in range(lo, hi):
                sci = synth.make_scene(500 + i, num_objects=10, cluttered=True)
                batch.append({"depth_heightmap": sci["scene"], "m_depth_heightmap": synth.masked_scene(sci["scene"], sci["masks"], [i % 10]),
                              "style": 0, "rotation": i % R, "label_value": float(i % 3)})
            tr.model.precision = precision
            tr.backprop_batch(batch, total=B, first_index=lo)        # warm-up: first sight of every rotation runs eagerly
            tr.backprop_batch(batch, total=B, first_index=lo)        # captures
            barrier()
            nrep = 2
            waits = []
            e0.record()
            for _ in range(nrep):
                _, w = tr.backprop_batch(batch, total=B, first_index=lo)
                waits.append(w)
            e1.record()
            barrier()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            grad_bytes = int(tr._fused[0]["flat"]["grad"].numel() * 4)
            rep = {"batch": B, "samples_per_gpu": hi - lo, "n_gpus": world, "scaling": "strong",
                   "ms_per_batch_step": float(t.item()) / nrep, "samples_per_s": B * nrep / (float(t.item()) / 1e3),
                   "allreduce_bytes": grad_bytes, "precision": precision,
                   "what": "Trainer.backprop_batch: per-sample grad-enabled pass + backward (graph replay, two interleaved pipelines per GPU), local sum, NCCL all-reduce of the "
                           "flat 368-tensor gradient, mean, one multi-tensor Adam, re-pack"}
            if world > 1:
                # the all-reduce alone on the same buffer (device time)
                Gf = tr._fused[0]["flat"]["grad"]
                for _ in range(3):
                    dist.all_reduce(Gf)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(10):
                    dist.all_reduce(Gf)
                e1.record()
                torch.cuda.synchronize()
                rep["allreduce_alone_ms"] = e0.elapsed_time(e1) / 10
                rep["allreduce_busbw_gbs"] = 2 * (world - 1) / world * grad_bytes / 1e9 / (rep["allreduce_alone_ms"] / 1e3)
            if world > 1 and not args.no_verify:
                # the all-reduced, averaged gradient of one more data-parallel step against the same batch accumulated
                # serially on this GPU alone, at the same weights
                st_f = tr._fused[0]
                w0 = [p.detach().clone() for p in st_f["params"]]
                tr.backprop_batch(batch, total=B, first_index=lo)
                g_dp = st_f["flat"]["grad"].clone()
                with torch.no_grad():
                    for p, w in zip(st_f["params"], w0):
                        p.copy_(w)
                eng.sync_weights(tr.model, force=True, style=0)
                full = []
                for i in range(B):
                    sci = synth.make_scene(500 + i, num_objects=10, cluttered=True)
                    full.append({"depth_heightmap": sci["scene"], "m_depth_heightmap": synth.masked_scene(sci["scene"], sci["masks"], [i % 10]),
                                 "style": 0, "rotation": i % R, "label_value": float(i % 3)})
                tr.backprop_batch(full, total=B, first_index=0, local_only=True)
                g_serial = st_f["flat"]["grad"]
                errs, off = [], 0
                for p in st_f["params"]:
                    a_, b_ = g_dp[off:off + p.numel()], g_serial[off:off + p.numel()]
                    off += p.numel()
                    errs.append(float((a_ - b_).abs().max() / b_.abs().max().clamp_min(1e-3 * float(g_serial.abs().max()))))
                errs.sort()
                e_t = torch.tensor([errs[len(errs) // 2], errs[-1]], device=dev)
                dist.all_reduce(e_t, op=dist.ReduceOp.MAX)
                rep["verify_vs_serial"] = {"median_rel_err": float(e_t[0]), "max_rel_err": float(e_t[1]),
                                           "what": "per-tensor max|g_dp - g_serial| / max|g_serial| of the averaged gradient, max over ranks "
                                                   "(differences: summation order + the ReLU-kink flips it causes)"}
            line_extra["replay"] = rep
        except Exception as exc:
            line_extra["replay"] = {"error": repr(exc)[:300]}
            if world > 1:
                raise

    # ---- backprop steps/s (second half of BASELINE.json's metric): Trainer.backprop through the public API, in the
    # benchmarked precision (tf32 forward + tensor-core data gradients) and in fp32 mode
    if not args.no_backprop:
        import smg_b200.synth as synth
        sc = synth.make_scene(100 + 1000 * rank, num_objects=4, cluttered=False)
        obj_masks = sc["masks"].astype(np.float64)
        nb = max(5, min(args.steps, 20))

        def bp(i):
            return tr.backprop(sc["scene"], "grasp", [i % 4, i % R], [0, 0], [], [], 1.0, obj_masks.copy(), [0] * 4, [0] * 4, [])

        for key, prec in (("backprop", precision), ("backprop_fp32", "fp32")):
            if key == "backprop_fp32" and (precision == "fp32" or args.no_extras):
                continue
            try:
                tr.model.precision = prec
                eng_t = tr.model._engine(2, 0)
                for i in range(2 * R):                  # first sight of each rotation runs eagerly, the second captures
                    bp(i)
                barrier()
                l0 = eng_t.launch_count()
                e0.record()
                for i in range(nb):
                    bp(i)
                e1.record()
                barrier()
                t = torch.tensor([e0.elapsed_time(e1)], device=dev)
                if world > 1:
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                line_extra[key] = {
                    "value": world * nb / (float(t.item()) / 1e3), "unit": "steps/s", "steps": nb,
                    "ms_per_step": float(t.item()) / nb, "gflop_per_step": 6 * GFLOP_PER_PASS, "precision": prec,
                    "launches_per_step": (eng_t.launch_count() - l0) // nb,
                    "tflops": 6 * GFLOP_PER_PASS / 1e3 / (float(t.item()) / nb / 1e3),
                    "what": "Trainer.backprop (fused smg_train_step, CUDA-graph replay): host heightmaps in, grad-enabled forward "
                            "(2 trunk passes + head), loss, backward, Adam, weight re-pack, BN running statistics, loss out"}
            except Exception as exc:  # the training path must never take the inference numbers down with it
                line_extra[key] = {"error": repr(exc)[:300]}
        tr.model.precision = precision

    # ---- BASELINE config 3: the reactive E+S policy (classification heads, weighted cross-entropy), forward + backward
    # through the same public calls (rank 0 only; no collective inside)
    if rank == 0 and not args.no_extras and not args.no_backprop:
        try:
            import smg_b200.synth as synth
            from smg_b200.trainer import Trainer
            sc = synth.make_scene(100, num_objects=4, cluttered=False)
            obj_masks = sc["masks"].astype(np.float64)
            trr = Trainer("reactive", 0.5, False, None, False, precision=precision)
            trr.image_mean, trr.image_std = MEAN, STD
            mk = synth.masked_scene(sc["scene"], sc["masks"], [0])

            def rfwd(i):
                return trr.forward(sc["scene"], mk, i % 2, True, False)

            def rbp(i):
                return trr.backprop(sc["scene"], "grasp" if i % 2 == 0 else "suction", [i % 4, 0], [i % 4, 0], [], [],
                                    float(i % 2), obj_masks.copy(), [0] * 4, [0] * 4, [])

            out = {}
            for key, fn, n_it in (("forward_calls_per_s", rfwd, 20), ("backprop_steps_per_s", rbp, 20)):
                for i in range(6):
                    fn(i)
                torch.cuda.synchronize()
                e0.record()
                for i in range(n_it):
                    fn(i)
                e1.record()
                torch.cuda.synchronize()
                out[key] = n_it / (e0.elapsed_time(e1) / 1e3)
            out["precision"] = precision
            out["what"] = ("reactive_net (3-class heads): Trainer.forward (rotation 0, softmax P(class 0), host heightmaps in, "
                           "probability out) and Trainer.backprop (weighted cross-entropy, fused captured step) alternating the "
                           "enveloping and sucking primitives")
            line_extra["reactive"] = out
            del trr
        except Exception as exc:
            line_extra["reactive"] = {"error": repr(exc)[:300]}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": {"fp32": "f32", "tf32": "tf32", "bf16": "bf16"}[precision],
                "data": "synthetic",
                "config": {"workload": WORKLOAD, "units_per_step_per_gpu": U, "rotations": R, "precision": precision,
                           "precision_note": note, "image_mean": MEAN, "image_std": STD,
                           "l2": "working set per step (%d samples x 87 MB of fp32 activations) exceeds the 126 MB L2" % (U * (R + 1)),
                           "parallelism": "dp%d over independent units; all_gather of per-GPU best (Q, rot)" % world},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "with_bn_running_stats": e2e_stats},
                "gpu_launches": int(launches), "clocks": clocks, "gflop_per_unit": GFLOP_PER_UNIT}
        line.update(line_extra)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-gpu"])
    ap.add_argument("--precision", default="tf32", choices=["fp32", "tf32", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-backprop", action="store_true")
    ap.add_argument("--no-gpu-reference", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip fp32_mode / running-stats / decision / replay extras")
    ap.add_argument("--no-verify", action="store_true", help="N > 1: skip checking the data-parallel replay gradient against the serial one")
    ap.add_argument("--units", type=int, default=4,
                    help="independent (scene, mask) units evaluated per step and GPU as one batch (1 = latency mode)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.impl == "reference-gpu":
        return run_reference_gpu_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
