"""Drop-in `Trainer` (reference: code/trainer.py:17-384) over the CUDA library.

Kept from the reference: constructor arguments, `forward(depth_heightmap, m_depth_heightmap, style,
is_volatile, is_target, specific_rotation)` -> numpy array of Q / P(class 0), `get_label_value`,
`backprop` (Huber / weighted CE + Adam lr 1e-4), `model` / `model_target` / `optimizer` attributes.

Differences, all deliberate:
  * `image_mean` / `image_std` are instance attributes defaulting to 0.01 / 0.03.  The reference's
    published literals are 0 / 0, which makes every prediction NaN.
  * pre-processing (zoom x2, pad, normalise, 3 channels) runs on the GPU fused with the rotation
    stage; only the 224x224 float64 heightmaps cross PCIe.
  * `forward_all` evaluates all objects x rotations of one primitive in ONE de-duplicated pass
    (the reference's step loop recomputes identical trunk passes, code/main.py:158-192).
  * a CUDA device is required; `force_cpu=True` raises instead of silently running elsewhere.
"""
import copy

import numpy as np
import torch

from . import engine as _engine
from .models import reactive_net, reinforcement_net
from .utils import CrossEntropyLoss2d


def group_is_plain_adam(opt):
    """The fused step implements exactly the reference's optimizer: one param group of plain Adam (code/trainer.py:99)."""
    if type(opt) is not torch.optim.Adam or len(opt.param_groups) != 1:
        return False
    g = opt.param_groups[0]
    return not g.get("amsgrad") and not g.get("maximize") and g.get("weight_decay", 0) == 0 and not g.get("capturable") \
        and not torch.is_tensor(g["lr"])


class Trainer(object):
    def __init__(self, method, future_reward_discount, load_snapshot, snapshot_file, force_cpu,
                 precision="fp32", device=None):
        self.method = method
        if force_cpu or not torch.cuda.is_available():
            raise RuntimeError("smg_b200.Trainer needs a CUDA device (B200); there is no CPU path")
        self.use_cuda = True
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self.image_mean = 0.01
        self.image_std = 0.03

        if self.method == 'reactive':
            self.model = reactive_net(self.use_cuda)
            w = torch.ones(3)
            w[2] = 0  # class 2 = "no loss" (code/trainer.py:38-45)
            self.suction_criterion = CrossEntropyLoss2d(w.to(self.device))
            self.grasp_criterion = CrossEntropyLoss2d(w.to(self.device))
            self.gs_criterion = CrossEntropyLoss2d(w.to(self.device))
            if load_snapshot:
                self.model.load_state_dict(torch.load(snapshot_file))
            self.model = self.model.to(self.device)
        elif self.method == 'reinforcement':
            self.model = reinforcement_net(self.use_cuda)
            self.model_target = copy.deepcopy(self.model)
            self.model_target.load_state_dict(self.model.state_dict())
            self.future_reward_discount = future_reward_discount
            if load_snapshot:
                self.model.load_state_dict(torch.load(snapshot_file))
            self.model = self.model.to(self.device)
            self.model_target = self.model_target.to(self.device)
        else:
            raise ValueError("method must be 'reactive' or 'reinforcement'")
        self.model.precision = precision
        if self.method == 'reinforcement':
            self.model_target.precision = precision
        self.model.train()
        self.optimizer = torch.optim.Adam(self.model.parameters(), lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0)
        # backprop() runs the whole step as ONE library call (smg_train_step: forward, loss, backward, Adam, re-pack inside a
        # CUDA graph).  False = the reference's own sequence: autograd node + loss.backward() + optimizer.step().
        self.fused_step = True
        self._fused = {}
        self._fused_last_style = None
        self.iteration = 0
        self.executed_action_log = []
        self.label_value_log = []
        self.reward_value_log = []
        self.predicted_value_log = []
        self.use_heuristic_log = []
        self.is_exploit_log = []
        self.clearance_log = []
        self.grasping_type_log = []
        self.episode_success_log = []
        self.training_loss_log = []

    # (log file stem, attribute, layout): "rows" = 2-D log cut to the first `iteration` rows, "col" = 1-D log cut to
    # `iteration` entries and stored as a column, "col_all" = 1-D log kept whole (the reference does not cut clearance)
    _LOGS = (("executed-action", "executed_action_log", "rows"), ("label-value", "label_value_log", "col"),
             ("predicted-value", "predicted_value_log", "col"), ("reward-value", "reward_value_log", "col"),
             ("use-heuristic", "use_heuristic_log", "col"), ("is-exploit", "is_exploit_log", "col"),
             ("clearance", "clearance_log", "col_all"), ("grasping_type", "grasping_type_log", "col"),
             ("episode_success", "episode_success_log", "rows"), ("training_loss", "training_loss_log", "rows"))

    def preload(self, transitions_directory):
        """Resume the ten `*.log.txt` files written by the reference's logger (code/trainer.py:118-158,
        code/logger.py:118-119): `iteration` = rows of the executed-action log minus two, every log becomes a Python
        list again (main.py appends to them and `logger.write_to_log` rewrites the whole file each step)."""
        import os

        def load(stem):
            return np.loadtxt(os.path.join(transitions_directory, stem + '.log.txt'), delimiter=' ')

        self.iteration = load(self._LOGS[0][0]).shape[0] - 2
        n = self.iteration
        for stem, attr, layout in self._LOGS:
            a = load(stem)
            if layout == "rows":
                a = a[0:n, :]
            elif layout == "col":
                a = a[0:n].reshape(n, 1)
            else:
                a = a.reshape(a.shape[0], 1)
            setattr(self, attr, a.tolist())

    # ------------------------------------------------------------------ forward
    def _rotations(self, model, style, specific_rotation):
        """Rotation indices / divisor exactly as the nets' forward picks them (code/models.py:361-510)."""
        if specific_rotation == -1:
            if style == 0:
                return list(range(model.gnum_rotations)), model.gnum_rotations
            if style == 1:
                return list(range(model.snum_rotations)), model.snum_rotations
            return [0], model.gnum_rotations
        return [0 if style == 2 else int(specific_rotation)], model.gnum_rotations

    def forward_all(self, depth_heightmap, m_depth_heightmaps, style=0, is_target=False, specific_rotation=-1):
        """Q table [n_masks, n_rot(, 3)] for one scene and MANY masked heightmaps in one de-duplicated pass."""
        model = self.model_target if (is_target and self.method == 'reinforcement') else self.model
        rots, nrot = self._rotations(model, style, specific_rotation)
        if torch.is_tensor(m_depth_heightmaps):
            # heightmaps already on the device (decision.decide forms the masked scenes there): no host round trip
            masks = m_depth_heightmaps if m_depth_heightmaps.dim() == 3 else m_depth_heightmaps[None]
            eng = model._engine(len(rots) + masks.shape[0], style)
            masks_t = masks.to(eng.device, torch.float64).contiguous()
            scene = (depth_heightmap if torch.is_tensor(depth_heightmap) else
                     torch.from_numpy(np.ascontiguousarray(depth_heightmap, dtype=np.float64))).to(eng.device, torch.float64).contiguous()
        else:
            masks = np.ascontiguousarray(np.asarray(m_depth_heightmaps, dtype=np.float64))
            if masks.ndim == 2:
                masks = masks[None]
            eng = model._engine(len(rots) + masks.shape[0], style)
            scene = torch.from_numpy(np.ascontiguousarray(depth_heightmap, dtype=np.float64)).to(eng.device, non_blocking=True)
            masks_t = torch.from_numpy(masks).to(eng.device, non_blocking=True)
        if model.update_running_stats:
            q, mean, var = eng.qforward_maps(style, scene, masks_t, self.image_mean, self.image_std, rots, nrot,
                                             want_bn_stats=True)
            trunk = getattr(model, _engine.TRUNK_ATTRS[_engine.STYLE_ROUTE[style][0]])
            order = []
            for k in range(masks.shape[0]):          # the reference loops objects, then rotations
                for i in range(len(rots)):
                    order += [i, len(rots) + k]
            model._apply_running_stats(trunk, mean, var, order)
            head = getattr(model, _engine.HEAD_ATTRS[_engine.STYLE_ROUTE[style][1]])
            pairs = [(i, len(rots) + k) for k in range(masks.shape[0]) for i in range(len(rots))]
            model._apply_head_running_stats(head, trunk, var, pairs, eng.head_bn_stats(len(pairs)))
        else:
            q = eng.qforward_maps(style, scene, masks_t, self.image_mean, self.image_std, rots, nrot)
        return q  # cuda tensor [n_masks, n_rot, n_out]

    def forward_batch(self, depth_heightmaps, m_depth_heightmaps, style=0, is_target=False, specific_rotation=-1):
        """G independent (scene, masked scene[s]) units in ONE pass: depth_heightmaps [G,224,224], m_depth_heightmaps
        [G,224,224] or [G,M,224,224] -> numpy Q [G, M, n_rot].  Each unit gets exactly the result `forward` would give
        it (BatchNorm statistics are per sample); batching only makes the late layers of the trunk more efficient.
        With `model.update_running_stats` the BatchNorm running statistics advance as if the G x M units had been evaluated one
        after the other by `forward` (unit by unit, rotation by rotation: trunk(scene_r) then trunk(mask))."""
        model = self.model_target if (is_target and self.method == 'reinforcement') else self.model
        rots, nrot = self._rotations(model, style, specific_rotation)
        scenes = np.ascontiguousarray(np.asarray(depth_heightmaps, dtype=np.float64))
        masks = np.ascontiguousarray(np.asarray(m_depth_heightmaps, dtype=np.float64))
        if masks.ndim == 3:
            masks = masks[:, None]
        G, M = masks.shape[0], masks.shape[1]
        eng = model._engine(G * (len(rots) + M), style)
        s_t = torch.from_numpy(scenes).to(eng.device, non_blocking=True)
        m_t = torch.from_numpy(masks).to(eng.device, non_blocking=True)
        if model.update_running_stats:
            q, mean, var = eng.qforward_maps_batch(style, s_t, m_t, self.image_mean, self.image_std, rots, nrot, want_bn_stats=True)
            nR = len(rots)
            order, pairs = [], []
            for g in range(G):
                for k in range(M):
                    for i in range(nR):
                        order += [g * nR + i, G * nR + g * M + k]
                        pairs.append((g * nR + i, G * nR + g * M + k))
            trunk = getattr(model, _engine.TRUNK_ATTRS[_engine.STYLE_ROUTE[style][0]])
            head = getattr(model, _engine.HEAD_ATTRS[_engine.STYLE_ROUTE[style][1]])
            model._apply_running_stats(trunk, mean, var, order)
            model._apply_head_running_stats(head, trunk, var, pairs, eng.head_bn_stats(len(pairs)))
        else:
            q = eng.qforward_maps_batch(style, s_t, m_t, self.image_mean, self.image_std, rots, nrot)
        if self.method == 'reactive':
            return torch.softmax(q[:, :, :1, :], dim=3)[..., 0].cpu().numpy()
        return q[..., 0].double().cpu().numpy()

    def forward(self, depth_heightmap, m_depth_heightmap, style=0, is_volatile=False, is_target=False, specific_rotation=-1):
        if not is_volatile:
            return self._forward_grad(depth_heightmap, m_depth_heightmap, style, specific_rotation)
        q = self.forward_all(depth_heightmap, m_depth_heightmap, style, is_target, specific_rotation)[0]
        if self.method == 'reactive':
            # the reference only looks at rotation 0 and returns softmax P(class 0) (code/trainer.py:195-199)
            return torch.softmax(q[0], dim=0)[0:1].cpu().numpy()
        return q[:, 0].double().cpu().numpy()

    def _forward_grad(self, depth_heightmap, m_depth_heightmap, style, specific_rotation):
        eng = self.model._engine(2)
        hm = torch.from_numpy(np.stack([np.asarray(depth_heightmap, np.float64), np.asarray(m_depth_heightmap, np.float64)]))
        x = eng.prep(hm.to(eng.device), self.image_mean, self.image_std)
        out = self.model.forward(x[0:1], x[1:2], style, False, specific_rotation)
        if self.method == 'reactive':
            return torch.softmax(out.detach().view(-1), dim=0)[0:1].cpu().numpy()
        return out.detach().view(-1).double().cpu().numpy()

    # ------------------------------------------------------------------ labels (code/trainer.py:212-274)
    def get_label_value(self, primitive_action, objects_number, suction_success, grasp_success, gs_success,
                        depth_heightmap, mask_depth, objects_mask, bestg_id, bests_id, bestgs_g_id, bestgs_s_id,
                        exploit_action, bestg_conf, bests_conf, bestgs_conf):
        if self.method == 'reactive':
            label_value = 0
            if primitive_action == 'suction':
                success_value = suction_success
                label_value = 0 if suction_success else 1
            elif primitive_action == 'grasp':
                success_value = grasp_success
                label_value = 0 if grasp_success else 1
            elif primitive_action == 'grasp_then_suction':
                success_value = gs_success
                label_value = 0 if gs_success == 2.5 else 1
            return label_value, success_value
        current_reward = 0
        if primitive_action == 'suction':
            current_reward = suction_success
        elif primitive_action == 'grasp':
            current_reward = grasp_success
        elif primitive_action == 'grasp_then_suction':
            current_reward = gs_success
        if suction_success == 0 and grasp_success == 0 and gs_success == 0:
            future_reward = 0
        elif (objects_number == 1 and suction_success == 1) or (objects_number == 1 and grasp_success == 1) or \
                (objects_number == 2 and gs_success == 2.5):
            future_reward = 0
        else:
            # Q_target(s', argmax_a Q(s', a)) at the best rotation (code/trainer.py:259-270)
            if exploit_action == 'grasp':
                m = depth_heightmap * mask_depth[bestg_id[0]]
                future_reward = self.forward(depth_heightmap, m, 0, True, True, bestg_id[1])[0]
            elif exploit_action == 'suction':
                m = depth_heightmap * mask_depth[bests_id[0]]
                future_reward = self.forward(depth_heightmap, m, 1, True, True, bests_id[1])[0]
            elif exploit_action == 'grasp_then_suction':
                m = depth_heightmap * (mask_depth[bestgs_g_id[0]] + mask_depth[bestgs_s_id[0]])
                future_reward = self.forward(depth_heightmap, m, 2, True, True, bestgs_g_id[1])[0]
            else:   # the reference leaves future_reward unbound here (UnboundLocalError)
                raise ValueError("unknown exploit_action %r" % (exploit_action,))
        expected_reward = current_reward + self.future_reward_discount * future_reward
        return expected_reward, current_reward

    # ------------------------------------------------------------------ fused step plumbing
    def _fused_state(self, style):
        """Per primitive: the 368 parameters the sample touches (trunk tensors in smg_set_trunk_weights order, then the
        head's) with flat gradient / Adam-moment buffers.  `p.grad` and `optimizer.state[p]` are views of those buffers in
        torch.optim.Adam's own format, so `optimizer.step()` / `state_dict()` keep working on the same state."""
        import ctypes
        model = self.model
        tid, hid = _engine.STYLE_ROUTE[int(style)]
        params = _engine.trunk_param_list(getattr(model, _engine.TRUNK_ATTRS[tid])) + \
            _engine.head_param_list(getattr(model, _engine.HEAD_ATTRS[hid]))
        st = self._fused.get(int(style))
        if st is not None and len(st["params"]) == len(params):
            # cheap validation on the per-step path (368 data_ptr() calls cost ~60 us): parameters move together (.cuda(),
            # .to()), so three probes decide; engine.sync_weights compares every pointer before the weights are used anyway
            probe = (0, len(params) // 2, len(params) - 1)
            if all(st["params"][i] is params[i] and st["key"][1][i] == params[i].data_ptr() for i in probe):
                return st
        key = (int(style), tuple(p.data_ptr() for p in params))
        for p in params:
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("smg_b200: training needs contiguous float32 parameters on the GPU")
        total = sum(p.numel() for p in params)
        dev = params[0].device
        flat = {k: torch.zeros(total, dtype=torch.float32, device=dev) for k in ("grad", "exp_avg", "exp_avg_sq")}
        views = {k: [] for k in flat}
        off = 0
        for p in params:
            for k in flat:
                views[k].append(flat[k][off:off + p.numel()].view_as(p))
            off += p.numel()
        # torch.optim.Adam keeps one scalar `step` tensor per parameter; here they are views of ONE flat CPU tensor, so the
        # per-step increment is a single add (torch._foreach_add_ over 368 CPU scalars costs ~1.5 ms of host time per step)
        steps_flat = torch.zeros(len(params), dtype=torch.float32)
        steps = []
        for i, p in enumerate(params):
            s = self.optimizer.state[p]
            if len(s) != 0:                       # already stepped by torch: adopt its moments
                views["exp_avg"][i].copy_(s["exp_avg"])
                views["exp_avg_sq"][i].copy_(s["exp_avg_sq"])
                steps_flat[i] = float(s["step"])
            step = steps_flat[i]                  # 0-dim view
            s["step"], s["exp_avg"], s["exp_avg_sq"] = step, views["exp_avg"][i], views["exp_avg_sq"][i]
            steps.append(step)

        def arr(ts):
            a = (ctypes.c_void_p * len(ts))()
            for i, t in enumerate(ts):
                a[i] = t.data_ptr()
            return a

        st = {"key": key, "params": params, "flat": flat, "views": views, "steps": steps, "steps_flat": steps_flat,
              "ptrs": (arr(params), arr(views["grad"]), arr(views["exp_avg"]), arr(views["exp_avg_sq"]))}
        self._fused[int(style)] = st
        return st

    def _backprop_fused(self, depth_heightmap, m_depth_heightmap, style, rot, label_value, attr):
        model = self.model
        eng = model._engine(2, style)
        st = self._fused_state(style)
        if self._fused_last_style != style:
            # optimizer.zero_grad() + a backward that only reaches this primitive's trunk and head (code/trainer.py:338-351):
            # every other parameter ends the step without a gradient, so Adam skips it
            self.optimizer.zero_grad()
            for p, g in zip(st["params"], st["views"]["grad"]):
                p.grad = g
            self._fused_last_style = style
        # the two heightmaps go through a pinned staging buffer (a pageable source makes the copy synchronous); the step ends
        # with a host read of the loss, so the buffer is free again when the next call fills it
        pin = st.get("hm_pin")
        d_np, m_np = np.asarray(depth_heightmap, np.float64), np.asarray(m_depth_heightmap, np.float64)
        if pin is None or pin.shape[1:] != d_np.shape:
            pin = st["hm_pin"] = torch.empty((2,) + d_np.shape, dtype=torch.float64).pin_memory()
            st["hm_pin_np"] = pin.numpy()
        np.copyto(st["hm_pin_np"][0], d_np)
        np.copyto(st["hm_pin_np"][1], m_np)
        hm = pin.to(eng.device, non_blocking=True)
        rot = 0 if style == 2 else int(rot)                 # ES is pinned to rotation 0 (code/models.py:567)
        group = self.optimizer.param_groups[0]
        step = int(st["steps"][0]) + 1
        eng._mean_std = (self.image_mean, self.image_std)
        if self.method == 'reactive':
            w = {0: self.grasp_criterion, 1: self.suction_criterion, 2: self.gs_criterion}[style].weight
            kind, cw = 1, [float(v) for v in w.cpu()]
        else:
            kind, cw = 0, [1.0, 1.0, 1.0]
        loss, q, mean, var = eng.train_step(style, hm[0], hm[1], rot, model.gnum_rotations, kind, float(label_value), cw,
                                            st["ptrs"], len(st["params"]), step, group["lr"], group["betas"][0],
                                            group["betas"][1], group["eps"], want_bn_stats=model.update_running_stats)
        st["steps_flat"] += 1
        _engine.bump_weight_epoch(model, style)                      # parameters changed in place: other handles must re-pack
        eng.mark_synced(model, style)                                # this one re-packed inside the step
        if model.update_running_stats:
            tid, hid = _engine.STYLE_ROUTE[int(style)]
            trunk, head = getattr(model, _engine.TRUNK_ATTRS[tid]), getattr(model, _engine.HEAD_ATTRS[hid])
            model._apply_running_stats(trunk, mean, var, [0, 1])
            model._apply_head_running_stats(head, trunk, var, [(0, 1)], eng.head_bn_stats(1))
        model.gra_prob, model.suc_prob, model.gs_prob = [], [], []
        setattr(model, attr, q.view(1, model.N_OUT, 1, 1))
        return loss.view(()).cpu().numpy()

    def backprop_batch(self, samples, group=None, total=None, first_index=0, reduction="mean", local_only=False):
        """Data-parallel replay step (BASELINE config 4): `samples` is THIS rank's share of a batch
        of `total` independent transitions, each a dict(depth_heightmap, m_depth_heightmap, style, rotation, label_value)
        with global position first_index + i.  Every sample runs the grad-enabled pass and the backward at the SAME weights
        (`smg_train_step` with SMG_STEP_GRADS_ONLY), consecutive samples alternating between two handles whose kernel chains
        interleave on the GPU; the per-sample gradients are summed locally, all-reduced over the ranks (one flat 28.5 MB
        buffer), averaged (`reduction="mean"`) and applied by ONE multi-tensor Adam launch on every rank.  The result equals the serial
        single-GPU step over the same batch; BatchNorm running statistics follow the weighted-sum rule of the serial order.
        All samples of a batch must use the same primitive (one trunk + head).  Returns (mean loss, seconds spent waiting
        for the all-reduce)."""
        import ctypes
        import os
        import time
        from . import parallel as _parallel
        rank, world = (0, 1) if local_only else _parallel._world(group)
        total = int(total if total is not None else len(samples))
        style = int(samples[0]["style"])
        model = self.model
        eng = model._engine(2, style)
        st = self._fused_state(style)
        if self._fused_last_style != style:
            self.optimizer.zero_grad()
            for p, g in zip(st["params"], st["views"]["grad"]):
                p.grad = g
            self._fused_last_style = style
        kind = 1 if self.method == 'reactive' else 0
        cw = [1.0, 1.0, 0.0] if kind else [1.0, 1.0, 1.0]
        group_ = self.optimizer.param_groups[0]
        G = st["flat"]["grad"]
        n = len(samples)
        # Two pipelines: the captured per-sample step is a chain of ~500 kernels, most of them far too small to fill 148 SMs
        # with two samples (blocks 3-4 run on 7-25 CTAs), so consecutive samples alternate between two handles (own
        # workspace, own streams, own gradient buffer, same weights) and their chains interleave on the GPU.
        # SMG_REPLAY_STREAMS=1 keeps the single pipeline.
        lanes = 2 if (n >= 4 and os.environ.get("SMG_REPLAY_STREAMS", "2") != "1") else 1
        pipes = [{"eng": eng, "G": G, "ptrs": st["ptrs"], "stream": torch.cuda.current_stream(eng.device)}]
        if lanes == 2:
            rp = st.get("replay2")
            if rp is None:
                G2 = torch.zeros_like(G)
                base = G.storage_offset()
                views2 = [G2[v.storage_offset() - base:v.storage_offset() - base + v.numel()].view_as(v) for v in st["views"]["grad"]]
                arr2 = (ctypes.c_void_p * len(views2))(*[v.data_ptr() for v in views2])
                import weakref
                weakref.finalize(model, _engine.drop_engine, ("replay2", id(model)))
                rp = st["replay2"] = {"G": G2, "views": views2, "ptrs": (st["ptrs"][0], arr2, st["ptrs"][2], st["ptrs"][3]),
                                      "stream": torch.cuda.Stream(device=eng.device), "acc": torch.zeros_like(G)}
            eng2 = _engine.get_engine(model._device_index(), 18, 640, model.precision, owner=("replay2", id(model)))
            eng2.sync_weights(model, force=True, style=style)   # the fused steps update the parameters in place: always re-pack
            pipes.append({"eng": eng2, "G": rp["G"], "ptrs": rp["ptrs"], "stream": rp["stream"], "acc": rp["acc"]})
        pipes[0]["acc"] = st.setdefault("acc", torch.zeros_like(G))
        tid, hid = _engine.STYLE_ROUTE[style]
        trunk, head = getattr(model, _engine.TRUNK_ATTRS[tid]), getattr(model, _engine.HEAD_ATTRS[hid])
        main = pipes[0]["stream"]
        for pp in pipes:
            pp["eng"]._mean_std = (self.image_mean, self.image_std)
            pp["stream"].wait_stream(main)
            with torch.cuda.stream(pp["stream"]):
                pp["acc"].zero_()
                pp["loss"] = torch.zeros(1, dtype=torch.float32, device=eng.device)
                pp["bn"] = None
        for i, smp in enumerate(samples):
            pp = pipes[i % lanes]
            e = pp["eng"]
            with torch.cuda.stream(pp["stream"]):
                hm = torch.from_numpy(np.stack([np.asarray(smp["depth_heightmap"], np.float64),
                                                np.asarray(smp["m_depth_heightmap"], np.float64)])).to(eng.device, non_blocking=True)
                rot = 0 if style == 2 else int(smp["rotation"])
                loss, q, mean, var = e.train_step(style, hm[0], hm[1], rot, model.gnum_rotations, kind, float(smp["label_value"]),
                                                  cw, pp["ptrs"], len(st["params"]), 1, want_bn_stats=model.update_running_stats,
                                                  grads_only=True)
                pp["loss"] += loss
                pp["acc"] += pp["G"]
                if model.update_running_stats:
                    j = 2 * (first_index + i)
                    w0 = _parallel.ema_pass_weights(j, 2 * total)
                    w1 = _parallel.ema_pass_weights(j + 1, 2 * total)
                    contrib = torch.cat([w0 * mean[0].double() + w1 * mean[1].double(), w0 * var[0].double() + w1 * var[1].double()])
                    # the head's two BatchNorms see one call per sample (see models._apply_head_running_stats)
                    wh = _parallel.ema_pass_weights(first_index + i, total)
                    norm5 = trunk.features.norm5
                    v5 = var[:, -1024:].double()
                    var_z = norm5.weight.double() ** 2 * v5 / (v5 + 1e-5)
                    h1 = e.head_bn_stats(1).double()[0]
                    unb = 400.0 / 399.0
                    contrib = torch.cat([contrib, wh * norm5.bias.double(), wh * norm5.bias.double(), wh * unb * var_z[0],
                                         wh * unb * var_z[1], wh * h1[0], wh * unb * h1[1]])
                    pp["bn"] = contrib if pp["bn"] is None else pp["bn"] + contrib
        for pp in pipes[1:]:
            main.wait_stream(pp["stream"])
        torch.add(pipes[0]["acc"], pipes[1]["acc"], out=G) if lanes == 2 else G.copy_(pipes[0]["acc"])
        loss_sum = pipes[0]["loss"] + pipes[1]["loss"] if lanes == 2 else pipes[0]["loss"]
        bn_sum = pipes[0]["bn"]
        if lanes == 2 and pipes[1]["bn"] is not None:
            bn_sum = pipes[1]["bn"] if bn_sum is None else bn_sum + pipes[1]["bn"]
        t0 = time.perf_counter()
        if world > 1:
            _parallel.allreduce_flat(G, group, async_op=False)      # one flat 28.5 MB all-reduce (0.1 ms over NVLink)
            torch.cuda.current_stream(eng.device).synchronize()
        wait_s = time.perf_counter() - t0
        if reduction == "mean":
            G /= float(total)
        if world > 1:
            dist_loss = loss_sum.clone()
            torch.distributed.all_reduce(dist_loss, group=group)
            loss_sum = dist_loss
        step = int(st["steps"][0]) + 1
        numel = st.get("numel")
        if numel is None:
            import ctypes
            numel = st["numel"] = (ctypes.c_int64 * len(st["params"]))(*[p.numel() for p in st["params"]])
        eng.adam_step_ptrs(st["ptrs"], numel, len(st["params"]), step, group_["lr"], group_["betas"][0], group_["betas"][1],
                           group_["eps"])
        st["steps_flat"] += 1
        _engine.bump_weight_epoch(model, style)
        eng.sync_weights(model, force=True, style=style)        # re-pack the updated weights
        if model.update_running_stats and bn_sum is not None:
            if world > 1:
                torch.distributed.all_reduce(bn_sum, group=group)
            C = _engine.TRUNK_BN_CHANNELS
            model._apply_running_sums(trunk, bn_sum[:C], bn_sum[C:2 * C], 2 * total)
            bns = [m for m in head.children() if isinstance(m, torch.nn.BatchNorm2d)]
            hsum = bn_sum[2 * C:]
            decay = 0.9 ** total
            with torch.no_grad():
                bns[0].running_mean.mul_(decay).add_(hsum[0:2048].to(bns[0].running_mean))
                bns[0].running_var.mul_(decay).add_(hsum[2048:4096].to(bns[0].running_var))
                bns[1].running_mean.mul_(decay).add_(hsum[4096:4160].to(bns[1].running_mean))
                bns[1].running_var.mul_(decay).add_(hsum[4160:4224].to(bns[1].running_var))
                bns[0].num_batches_tracked += total
                bns[1].num_batches_tracked += total
        return float(loss_sum.item()) / total, wait_s

    # ------------------------------------------------------------------ backprop (code/trainer.py:278-384)
    def backprop(self, depth_heightmap, primitive_action, bestg_id, bests_id, bestgs_g_id, bestgs_s_id,
                 label_value, objects_mask, sro_best, gro_best, bestgs_num):
        mask_depth = np.asarray(objects_mask).reshape(objects_mask.shape[0], objects_mask.shape[1], objects_mask.shape[2])
        if primitive_action == 'grasp':
            style, m, rot, attr = 0, depth_heightmap * mask_depth[bestg_id[0]], bestg_id[1], 'gra_prob'
        elif primitive_action == 'suction':
            style, m, rot, attr = 1, depth_heightmap * mask_depth[bests_id[0]], bests_id[1], 'suc_prob'
        elif primitive_action == 'grasp_then_suction':
            style = 2
            m = depth_heightmap * (mask_depth[bestgs_g_id[0]] + mask_depth[bestgs_s_id[0]])
            rot, attr = bestgs_g_id[1], 'gs_prob'
        else:
            raise ValueError(primitive_action)
        if self.fused_step and group_is_plain_adam(self.optimizer):
            return self._backprop_fused(depth_heightmap, m, style, rot, label_value, attr)
        self._fused_last_style = None
        self.optimizer.zero_grad()
        self.forward(depth_heightmap, m, style=style, is_volatile=False, is_target=False, specific_rotation=rot)
        out = getattr(self.model, attr)
        if self.method == 'reactive':
            label = torch.full((1, 1, 1), int(label_value), dtype=torch.long, device=out.device)
            crit = {0: self.grasp_criterion, 1: self.suction_criterion, 2: self.gs_criterion}[style]
            loss = crit(out[0].view([1, 3, 1, 1]), label)
        else:
            d = out[0, 0, 0, 0] - label_value
            loss = 0.5 * (d ** 2) if abs(float(d)) < 1 else abs(d) - 0.5  # hand-written Huber (code/trainer.py:345-348)
        loss = loss.sum()
        loss.backward()
        loss_value = loss.detach().cpu().numpy()
        self.optimizer.step()
        return loss_value
