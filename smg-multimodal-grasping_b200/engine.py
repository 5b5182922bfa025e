"""Host-side owner of one `smg_handle` (one per GPU): weight re-packing and raw calls.

PyTorch is used for device memory, streams and tensors only; every kernel that runs
comes from csrc/libsmg_b200.so through the C ABI in include/smg_b200.h.
"""
import ctypes

import numpy as np
import torch

from . import _lib

TRUNK_ATTRS = ("suction_depth_trunk", "grasp_depth_trunk", "gs_depth_trunk")  # == SMG_TRUNK_* ids
HEAD_ATTRS = ("suctionnet_val", "graspnet_val", "gsnet_val")                  # == SMG_HEAD_* ids
# style -> (trunk id, head id); style 2 feeds the gs trunk to the SUCTION head (code/models.py:434,507,582)
STYLE_ROUTE = {0: (1, 1), 1: (0, 0), 2: (2, 0)}
PRECISIONS = {"fp32": 0, "tf32": 1, "bf16": 2}
TRUNK_BN_CHANNELS = 41824

_engines = {}


def trunk_param_list(trunk):
    """The 362 conv / BN-affine tensors of `densenet121().features` in state_dict order."""
    out = []
    for name, p in trunk.features.named_parameters():
        out.append(p)
    return out


def head_param_list(head):
    """norm0.w, norm0.b, conv0.w, norm1.w, norm1.b, conv1.w of a head Sequential (code/models.py:316-323)."""
    mods = list(head.children())
    norm0, conv0, norm1, conv1 = mods[0], mods[2], mods[3], mods[5]
    return [norm0.weight, norm0.bias, conv0.weight, norm1.weight, norm1.bias, conv1.weight]


def _param_signature(model, epoch, params):
    """What sync_weights compares per call: every parameter's torch version counter (load_state_dict, optimizer.step and any
    in-place op bump it), the library's own in-place update count, and the storage address of three probes - parameters
    change address together (.cuda(), .to()), and reading all 362 addresses on every forward call costs ~80 us."""
    n = len(params)
    return (id(model), epoch, n, params[0].data_ptr(), params[n // 2].data_ptr(), params[-1].data_ptr(),
            tuple([p._version for p in params]))


def _ptr_array(tensors):
    arr = (ctypes.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


class Engine:
    """One handle bound to one CUDA device."""

    def __init__(self, device=0, max_samples=32, H=640, precision="fp32"):
        if not torch.cuda.is_available():
            raise _lib.SmgError("smg_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device("cuda", device if isinstance(device, int) else torch.device(device).index or 0)
        self.max_samples = int(max_samples)   # 0: stateless kernels only (heightmap, NMS, argmax)
        self.H = int(H)
        h = ctypes.c_void_p()
        _lib.check(self.lib.smg_create(self.device.index, self.max_samples, self.H, ctypes.byref(h)))
        self.h = h
        self._sig = {}
        self._param_lists = {}
        self._staged = {}
        self._mean_std = (0.01, 0.03)   # Trainer.image_mean / image_std, set by the Trainer before a fused step
        self.set_precision(precision)

    def __del__(self):
        try:
            if getattr(self, "h", None) is not None and self.h:
                self.lib.smg_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ------------------------------------------------------------------ misc
    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def set_precision(self, precision):
        if getattr(self, "precision", None) != precision:
            self._sig = {}  # the packed weight layouts depend on the precision: re-pack on the next sync
        self.precision = precision
        _lib.check(self.lib.smg_set_precision(self.h, PRECISIONS[precision]))
        # pack only what this precision needs, plus the data-gradient layout used by the training step
        _lib.check(self.lib.smg_set_pack_layouts(self.h, (1 << PRECISIONS[precision]) | 8))

    def launch_count(self):
        return int(self.lib.smg_launch_count(self.h))

    def workspace_bytes(self):
        return int(self.lib.smg_workspace_bytes(self.h))

    # ------------------------------------------------------------------ weights
    def _device_params(self, key, params):
        """fp32 contiguous tensors on this device (parameters living on the CPU are staged)."""
        out = []
        for i, p in enumerate(params):
            t = p.detach()
            if t.device != self.device or t.dtype != torch.float32 or not t.is_contiguous():
                t = t.to(self.device, torch.float32).contiguous()
            out.append(t)
        self._staged[key] = out  # keep staged copies alive until the packing kernels ran
        return out

    def sync_weights(self, model, force=False, style=None):
        """Re-pack the model's parameters if they changed (load_state_dict, optimizer step, .cuda()).
        `style` restricts the check to the trunk and head that style routes to (the per-call cost on the inference
        path: 368 (data_ptr, version) pairs instead of 1104).  The Parameter lists are cached per (model, submodule):
        torch keeps Parameter identity across .cuda() / load_state_dict / optimizer steps.  The fused training step updates
        the parameters in place from a library kernel (no version bump): the trainer counts those updates in
        `model._smg_epoch` (per trunk / head), which is part of the signature, so that OTHER handles of the same model re-pack too."""
        n_out = 3 if model.__class__.__name__ == "reactive_net" else 1
        epochs = getattr(model, "_smg_epoch", None) or {}
        tids, hids = range(len(TRUNK_ATTRS)), range(len(HEAD_ATTRS))
        if style is not None:
            t_only, h_only = STYLE_ROUTE[int(style)]
            tids, hids = (t_only,), (h_only,)
        for tid in tids:
            key = ("t", tid)
            sub = getattr(model, TRUNK_ATTRS[tid])
            params = self._param_lists.get((id(model), key, id(sub)))
            if params is None or force:
                params = self._param_lists[(id(model), key, id(sub))] = trunk_param_list(sub)
            sig = _param_signature(model, epochs.get(key, 0), params)
            if force or self._sig.get(key) != sig:
                dev = self._device_params(key, params)
                _lib.check(self.lib.smg_set_trunk_weights(self.h, tid, _ptr_array(dev), len(dev), self._stream()))
                self._sig[key] = sig
        for hid in hids:
            key = ("h", hid)
            sub = getattr(model, HEAD_ATTRS[hid])
            params = self._param_lists.get((id(model), key, id(sub)))
            if params is None or force:
                params = self._param_lists[(id(model), key, id(sub))] = head_param_list(sub)
            sig = _param_signature(model, epochs.get(key, 0), params)
            if force or self._sig.get(key) != sig:
                dev = self._device_params(key, params)
                _lib.check(self.lib.smg_set_head_weights(self.h, hid, _ptr_array(dev), n_out, self._stream()))
                self._sig[key] = sig
        self.n_out = n_out

    def mark_synced(self, model, style):
        """The packed weights of `style`'s trunk and head on THIS handle are current (the fused training step re-packs them
        itself): record the signature without packing again."""
        epochs = getattr(model, "_smg_epoch", None) or {}
        tid, hid = STYLE_ROUTE[int(style)]
        for key, attr in ((("t", tid), TRUNK_ATTRS[tid]), (("h", hid), HEAD_ATTRS[hid])):
            params = self._param_lists.get((id(model), key, id(getattr(model, attr))))
            if params is not None:
                self._sig[key] = _param_signature(model, epochs.get(key, 0), params)


    # ------------------------------------------------------------------ K1
    def prep(self, heightmaps, mean, std):
        """[n,hs,hs] float64 (cuda) -> [n,3,H,H] float32: code/trainer.py:165-191."""
        hm = heightmaps.to(self.device, torch.float64).contiguous()
        n, hs = hm.shape[0], hm.shape[-1]
        out = torch.empty((n, 3, self.H, self.H), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.smg_prep(self.h, hm.data_ptr(), n, hs, float(mean), float(std), out.data_ptr(), self._stream()))
        return out

    def rotate(self, image, rot_idx, num_rotations):
        """[3,H,H] float32 -> [len(rot_idx),3,H,H]: code/models.py:371-382."""
        img = image.to(self.device, torch.float32).contiguous().view(3, self.H, self.H)
        rot = (ctypes.c_int * len(rot_idx))(*[int(r) for r in rot_idx])
        out = torch.empty((len(rot_idx), 3, self.H, self.H), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.smg_rotate(self.h, img.data_ptr(), rot, len(rot_idx), int(num_rotations), out.data_ptr(),
                                       self._stream()))
        return out

    def rotate_index_map(self, rot_idx, num_rotations):
        out = torch.empty((self.H, self.H), dtype=torch.int32, device=self.device)
        _lib.check(self.lib.smg_rotate_index_map(self.h, int(rot_idx), int(num_rotations), out.data_ptr(), self._stream()))
        return out

    # ------------------------------------------------------------------ trunk / Q
    def trunk_forward(self, trunk_id, x, want_bn_stats=False):
        """`trunk.features(x)` for x [n,3,H,H] -> [n,1024,H/32,H/32] (+ optional per-sample BN stats)."""
        x = x.to(self.device, torch.float32).contiguous()
        n = x.shape[0]
        feat = torch.empty((n, 1024, self.H // 32, self.H // 32), dtype=torch.float32, device=self.device)
        mean = var = None
        pm = pv = None
        if want_bn_stats:
            mean = torch.empty((n, TRUNK_BN_CHANNELS), dtype=torch.float32, device=self.device)
            var = torch.empty_like(mean)
            pm, pv = mean.data_ptr(), var.data_ptr()
        _lib.check(self.lib.smg_trunk_forward(self.h, int(trunk_id), x.data_ptr(), n, feat.data_ptr(), pm, pv, self._stream()))
        return (feat, mean, var) if want_bn_stats else feat

    def qforward(self, style, scene, masks, rot_idx, num_rotations, want_bn_stats=False):
        """Q for every (mask, rotation): [n_masks, n_rot, n_out].  scene [3,H,H], masks [n,3,H,H] float32."""
        tid, hid = STYLE_ROUTE[int(style)]
        scene = scene.to(self.device, torch.float32).contiguous()
        masks = masks.to(self.device, torch.float32).contiguous().view(-1, 3, self.H, self.H)
        n_masks, n_rot = masks.shape[0], len(rot_idx)
        rot = (ctypes.c_int * n_rot)(*[int(r) for r in rot_idx])
        q = torch.empty((n_masks, n_rot, self.n_out), dtype=torch.float32, device=self.device)
        mean = var = None
        pm = pv = None
        if want_bn_stats:
            mean = torch.empty((n_rot + n_masks, TRUNK_BN_CHANNELS), dtype=torch.float32, device=self.device)
            var = torch.empty_like(mean)
            pm, pv = mean.data_ptr(), var.data_ptr()
        _lib.check(self.lib.smg_qforward(self.h, tid, hid, scene.data_ptr(), masks.data_ptr(), n_masks, rot, n_rot,
                                         int(num_rotations), q.data_ptr(), pm, pv, self._stream()))
        return (q, mean, var) if want_bn_stats else q

    def qforward_maps(self, style, scene_hm, mask_hms, mean, std, rot_idx, num_rotations, want_bn_stats=False):
        """Same from 224x224 float64 heightmaps already on the device (fuses code/trainer.py:165-191)."""
        tid, hid = STYLE_ROUTE[int(style)]
        n_masks, hs = mask_hms.shape[0], mask_hms.shape[-1]
        n_rot = len(rot_idx)
        rot = (ctypes.c_int * n_rot)(*[int(r) for r in rot_idx])
        q = torch.empty((n_masks, n_rot, self.n_out), dtype=torch.float32, device=self.device)
        mean_t = var_t = None
        pm = pv = None
        if want_bn_stats:
            mean_t = torch.empty((n_rot + n_masks, TRUNK_BN_CHANNELS), dtype=torch.float32, device=self.device)
            var_t = torch.empty_like(mean_t)
            pm, pv = mean_t.data_ptr(), var_t.data_ptr()
        _lib.check(self.lib.smg_qforward_maps(self.h, tid, hid, scene_hm.data_ptr(), mask_hms.data_ptr(), n_masks, hs,
                                              float(mean), float(std), rot, n_rot, int(num_rotations), q.data_ptr(),
                                              pm, pv, self._stream()))
        return (q, mean_t, var_t) if want_bn_stats else q

    def qpartials(self, style, scene_hm, rot_idx, num_rotations, mask_hms, mean, std):
        """Per-sample halves of the head's first convolution for the LISTED rotations of `scene_hm` and the masked heightmaps
        `mask_hms` [n,hs,hs] (either list may be empty): P [len(rot_idx) + n, 400, 64].  Gathered partials of all ranks are
        paired by `qcombine` - the multi-GPU split of one decision (smg_qpartials / smg_qcombine)."""
        tid, hid = STYLE_ROUTE[int(style)]
        n_rot = len(rot_idx)
        n_masks = 0 if mask_hms is None else int(mask_hms.shape[0])
        hs = int(scene_hm.shape[-1])
        rot = (ctypes.c_int * max(n_rot, 1))(*[int(r) for r in rot_idx])
        p = torch.empty((n_rot + n_masks, 400, 64), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.smg_qpartials(self.h, tid, hid, scene_hm.data_ptr(), rot, n_rot, int(num_rotations),
                                          mask_hms.data_ptr() if n_masks else None, n_masks, hs, float(mean), float(std),
                                          p.data_ptr(), self._stream()))
        return p

    def qcombine(self, style, p_scene, p_mask):
        """Q [n_masks, n_rot, n_out] from per-sample partials (scene partials [n_rot,400,64], mask partials [n_masks,400,64])."""
        hid = STYLE_ROUTE[int(style)][1]
        p_scene, p_mask = p_scene.contiguous(), p_mask.contiguous()
        q = torch.empty((p_mask.shape[0], p_scene.shape[0], self.n_out), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.smg_qcombine(self.h, hid, p_scene.data_ptr(), int(p_scene.shape[0]), p_mask.data_ptr(),
                                         int(p_mask.shape[0]), q.data_ptr(), self._stream()))
        return q

    def qforward_maps_batch(self, style, scene_hms, mask_hms, mean, std, rot_idx, num_rotations, want_bn_stats=False):
        """G independent units in one batch: scene_hms [G,hs,hs], mask_hms [G,M,hs,hs] float64 on the device
        -> Q [G, M, n_rot, n_out].  Same per-unit results as qforward_maps (BatchNorm is per sample).  With want_bn_stats
        also the per-sample batch statistics [G*(n_rot+M), C] (all rotated scenes first, then all masked scenes)."""
        tid, hid = STYLE_ROUTE[int(style)]
        G, M, hs = mask_hms.shape[0], mask_hms.shape[1], mask_hms.shape[-1]
        n_rot = len(rot_idx)
        rot = (ctypes.c_int * n_rot)(*[int(r) for r in rot_idx])
        q = torch.empty((G, M, n_rot, self.n_out), dtype=torch.float32, device=self.device)
        mean_t = var_t = None
        pm = pv = None
        if want_bn_stats:
            mean_t = torch.empty((G * (n_rot + M), TRUNK_BN_CHANNELS), dtype=torch.float32, device=self.device)
            var_t = torch.empty_like(mean_t)
            pm, pv = mean_t.data_ptr(), var_t.data_ptr()
        _lib.check(self.lib.smg_qforward_maps_batch(self.h, tid, hid, scene_hms.data_ptr(), mask_hms.data_ptr(), G, M, hs,
                                                    float(mean), float(std), rot, n_rot, int(num_rotations), q.data_ptr(),
                                                    pm, pv, self._stream()))
        return (q, mean_t, var_t) if want_bn_stats else q

    # ------------------------------------------------------------------ training (code/trainer.py:278-384)
    def qforward_train(self, style, scene, mask, rot, num_rotations):
        """One grad-enabled Q evaluation (rotation `rot`); keeps what smg_qbackward needs inside the handle.
        Returns (q [n_out], bn_mean, bn_var [2, TRUNK_BN_CHANNELS])."""
        tid, hid = STYLE_ROUTE[int(style)]
        scene = scene.to(self.device, torch.float32).contiguous()
        mask = mask.to(self.device, torch.float32).contiguous()
        q = torch.empty((self.n_out,), dtype=torch.float32, device=self.device)
        mean = torch.empty((2, TRUNK_BN_CHANNELS), dtype=torch.float32, device=self.device)
        var = torch.empty_like(mean)
        _lib.check(self.lib.smg_qforward_train(self.h, tid, hid, scene.data_ptr(), mask.data_ptr(), int(rot),
                                               int(num_rotations), q.data_ptr(), mean.data_ptr(), var.data_ptr(),
                                               self._stream()))
        return q, mean, var

    def train_pass_id(self):
        """Stamp of the pending grad-enabled pass on this handle, -1 if none (any other forward invalidates it)."""
        if getattr(self, "h", None) is None or not self.h:
            return -1
        return int(self.lib.smg_train_pass_id(self.h))

    def train_step(self, style, scene_hm, mask_hm, rot, num_rotations, loss_kind, label, class_weight, ptrs, n_tensors,
                   step, lr=1e-4, beta1=0.9, beta2=0.999, eps=1e-8, want_bn_stats=True, grads_only=False):
        """One whole training step (smg_train_step): pre-processing, grad-enabled Q pass, loss, backward, Adam, re-pack.
        scene_hm / mask_hm: [hs,hs] float64 on this device; ptrs: (params, grads, exp_avg, exp_avg_sq) ctypes pointer
        arrays of `n_tensors` device tensors.  Returns (loss [1], q [n_out], bn_mean, bn_var [2, C] or None)."""
        tid, hid = STYLE_ROUTE[int(style)]
        args = _lib.TrainStepArgs(tid, hid, int(rot), int(num_rotations), int(scene_hm.shape[-1]), int(loss_kind), int(step),
                                  float(label), float(self._mean_std[0]), float(self._mean_std[1]),
                                  (ctypes.c_float * 3)(*[float(w) for w in class_weight]), float(lr), float(beta1),
                                  float(beta2), float(eps), 1 if grads_only else 0)
        loss = torch.empty((1,), dtype=torch.float32, device=self.device)
        q = torch.empty((self.n_out,), dtype=torch.float32, device=self.device)
        mean = var = None
        pm = pv = None
        if want_bn_stats:
            mean = torch.empty((2, TRUNK_BN_CHANNELS), dtype=torch.float32, device=self.device)
            var = torch.empty_like(mean)
            pm, pv = mean.data_ptr(), var.data_ptr()
        _lib.check(self.lib.smg_train_step(self.h, ctypes.byref(args), scene_hm.data_ptr(), mask_hm.data_ptr(), ptrs[0],
                                           ptrs[1], ptrs[2], ptrs[3], int(n_tensors), loss.data_ptr(), q.data_ptr(), pm, pv,
                                           self._stream()))
        return loss, q, mean, var

    def head_bn_stats(self, n_pairs):
        """(mean, biased var) [n_pairs, 2, 64] of the head's BatchNorm2d(64) for the last Q pass on this handle."""
        out = torch.empty((n_pairs, 2, 64), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.smg_head_bn_stats(self.h, out.data_ptr(), int(n_pairs), self._stream()))
        return out

    def qbackward(self, dq, trunk_grads, head_grads):
        """dLoss/dQ [n_out] -> gradients written into the given tensors (smg_set_*_weights order)."""
        dq = dq.to(self.device, torch.float32).contiguous()
        _lib.check(self.lib.smg_qbackward(self.h, dq.data_ptr(), _ptr_array(trunk_grads), _ptr_array(head_grads),
                                          self._stream()))

    def adam_step(self, params, grads, exp_avg, exp_avg_sq, step, lr=1e-4, beta1=0.9, beta2=0.999, eps=1e-8):
        """Fused Adam over a list of tensors (torch.optim.Adam semantics, code/trainer.py:99)."""
        n = len(params)
        numel = (ctypes.c_int64 * n)(*[p.numel() for p in params])
        _lib.check(self.lib.smg_adam_step(self.h, _ptr_array(params), _ptr_array(grads), _ptr_array(exp_avg),
                                          _ptr_array(exp_avg_sq), numel, n, int(step), float(lr), float(beta1),
                                          float(beta2), float(eps), self._stream()))

    def adam_step_ptrs(self, ptrs, numel, n, step, lr=1e-4, beta1=0.9, beta2=0.999, eps=1e-8):
        """Same with pre-built ctypes pointer arrays (params, grads, exp_avg, exp_avg_sq) and element counts."""
        _lib.check(self.lib.smg_adam_step(self.h, ptrs[0], ptrs[1], ptrs[2], ptrs[3], numel, int(n), int(step), float(lr),
                                          float(beta1), float(beta2), float(eps), self._stream()))

    def debug_read(self, what, sample, shape):
        out = torch.empty(shape, dtype=torch.float32, device=self.device)
        _lib.check(self.lib.smg_debug_read(self.h, what.encode(), int(sample), out.data_ptr(), out.numel(), self._stream()))
        return out

    def profile_enable(self, enable=True):
        _lib.check(self.lib.smg_profile_enable(self.h, int(bool(enable))))

    def profile_read(self):
        """Per kernel class (stem, conv1x1, conv3x3, other): dict(ms, launches, flops, bytes) since profile_enable."""
        ms = (ctypes.c_double * 4)()
        ln = (ctypes.c_int64 * 4)()
        fl = (ctypes.c_double * 4)()
        by = (ctypes.c_double * 4)()
        _lib.check(self.lib.smg_profile_read(self.h, ms, ln, fl, by))
        names = ("stem", "conv1x1", "conv3x3", "other")
        return {names[i]: {"ms": ms[i], "launches": int(ln[i]), "flops": fl[i], "bytes": by[i]} for i in range(4)}

    def debug_conv(self, precision, x_nhwc, cin, scale, shift, relu, pool, w_oihw, out_cstride=None, out_coff=0,
                   want_stats=True):
        """One generic-conv launch on caller tensors (unit-test hook).  Returns (out NHWC, stats [n,C,2] f64)."""
        n, hin, _, cstride = x_nhwc.shape
        cout, taps = w_oihw.shape[0], w_oihw.shape[2] * w_oihw.shape[3]
        hout = hin // 2 if pool else hin
        out_cstride = out_cstride or cout
        out = torch.zeros((n, hout, hout, out_cstride), dtype=torch.float32, device=self.device)
        stats = torch.zeros((n, out_cstride, 2), dtype=torch.float64, device=self.device) if want_stats else None
        _lib.check(self.lib.smg_debug_conv(
            self.h, PRECISIONS[precision], x_nhwc.contiguous().data_ptr(), n, hin, cin, cstride,
            scale.contiguous().data_ptr(), shift.contiguous().data_ptr(), int(relu), int(pool), taps,
            w_oihw.contiguous().data_ptr(), cout, out.data_ptr(), out_cstride, out_coff,
            stats.data_ptr() if want_stats else None, self._stream()))
        return out, stats

    # ------------------------------------------------------------------ K9 / K11 / K12
    def argmax(self, q):
        q = q.to(self.device, torch.float32).contiguous().view(-1)
        val = torch.empty(1, dtype=torch.float32, device=self.device)
        idx = torch.empty(1, dtype=torch.int32, device=self.device)
        _lib.check(self.lib.smg_argmax(self.h, q.data_ptr(), q.numel(), val.data_ptr(), idx.data_ptr(), self._stream()))
        return val, idx

    def heightmap(self, depth, intrinsics, pose):
        """depth [480,640] float64 -> (depth_heightmap [224,224], depth_mask [448,448], A_htor [3,3])."""
        d = depth.to(self.device, torch.float64).contiguous()
        K = np.ascontiguousarray(intrinsics, dtype=np.float64)
        P = np.ascontiguousarray(pose, dtype=np.float64)
        A = np.empty(9, dtype=np.float64)
        o224 = torch.empty((224, 224), dtype=torch.float64, device=self.device)
        o448 = torch.empty((448, 448), dtype=torch.float64, device=self.device)
        dp = ctypes.POINTER(ctypes.c_double)
        _lib.check(self.lib.smg_heightmap(self.h, d.data_ptr(), K.ctypes.data_as(dp), P.ctypes.data_as(dp), o224.data_ptr(),
                                          o448.data_ptr(), A.ctypes.data_as(dp), self._stream()))
        return o224, o448, A.reshape(3, 3)

    def heightmap_color(self, color):
        """color [480,640,3] uint8 -> (color_heightmap [224,224,3], color_mask [448,448,3]) uint8 (cv2 fixed-point remap)."""
        c = color.to(self.device, torch.uint8).contiguous()
        if tuple(c.shape) != (480, 640, 3):
            raise ValueError("heightmap_color expects a [480,640,3] uint8 image, got %s" % (tuple(c.shape),))
        o224 = torch.empty((224, 224, 3), dtype=torch.uint8, device=self.device)
        o448 = torch.empty((448, 448, 3), dtype=torch.uint8, device=self.device)
        _lib.check(self.lib.smg_heightmap_color(self.h, c.data_ptr(), o224.data_ptr(), o448.data_ptr(), self._stream()))
        return o224, o448

    def resize_masks(self, masks, size_out=224):
        """[n,s,s] float32 -> [n,size_out,size_out]: bilinear, align_corners=True (code/masks.py:51)."""
        m = masks.to(self.device, torch.float32).contiguous()
        n, s = int(m.shape[0]), int(m.shape[-1])
        out = torch.empty((n, size_out, size_out), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.smg_resize_masks(self.h, m.data_ptr() if n else None, n, s, int(size_out), out.data_ptr() if n else None,
                                             self._stream()))
        return out

    def geometry(self, mode, depth_img, A_htor, intrinsics, pose, boxes=None, centers=None, best=0, flag=0, pix=None):
        """smg_geometry: mode 0 global_position, 1 grasp angle / opening, 2 suction direction -> numpy [5]."""
        d = depth_img if torch.is_tensor(depth_img) else torch.from_numpy(np.ascontiguousarray(depth_img, dtype=np.float64))
        d = d.to(self.device, torch.float64).contiguous()
        dp = ctypes.POINTER(ctypes.c_double)

        def arr(a, n=None):
            if a is None:
                return None, None
            v = np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1))
            return v, v.ctypes.data_as(dp)

        A, pA = arr(np.asarray(A_htor)[:3, :3])
        K, pK = arr(np.asarray(intrinsics)[:3, :3])
        P, pP = arr(np.asarray(pose)[:4, :4])
        B, pB = arr(boxes)
        C, pC = arr(centers)
        X, pX = arr(pix)
        n = 0 if boxes is None else int(np.asarray(boxes).shape[0])
        out = np.zeros(5, dtype=np.float64)
        _lib.check(self.lib.smg_geometry(self.h, int(mode), d.data_ptr(), int(d.shape[0]), int(d.shape[1]), pA, pK, pP, pB, pC, n,
                                         int(best), int(bool(flag)), pX, out.ctypes.data_as(dp), self._stream()))
        return out

    def nms(self, boxes, co_thresh, min_area, max_area):
        b = boxes.to(self.device, torch.float32).contiguous().view(-1, 4)
        n = b.shape[0]
        keep = torch.empty(max(n, 1), dtype=torch.int32, device=self.device)
        cnt = torch.zeros(1, dtype=torch.int32, device=self.device)
        _lib.check(self.lib.smg_nms(self.h, b.data_ptr() if n else None, n, float(co_thresh), float(min_area),
                                    float(max_area), keep.data_ptr(), cnt.data_ptr(), self._stream()))
        return keep, cnt


def bump_weight_epoch(model, style):
    """The fused training step changed the parameters of `style`'s trunk and head in place (no torch version bump): make every
    handle of this model see it at its next sync_weights."""
    tid, hid = STYLE_ROUTE[int(style)]
    epochs = getattr(model, "_smg_epoch", None)
    if epochs is None:
        epochs = {}
        object.__setattr__(model, "_smg_epoch", epochs)
    for key in (("t", tid), ("h", hid)):
        epochs[key] = epochs.get(key, 0) + 1


def drop_engine(owner):
    """Release the engines of a model that is going away."""
    for key in [k for k in _engines if k[2] == owner]:
        _engines.pop(key).__del__()


def stateless_engine(device=0):
    """A handle without trunk workspace for the kernels that need none (heightmap, NMS, argmax)."""
    return get_engine(device, 0, 640, owner="stateless")


def get_engine(device=0, max_samples=32, H=640, precision=None, owner=None):
    """Engine for (device, owner); `owner` separates models with different weights (model vs model_target).
    The workspace grows on demand."""
    if isinstance(device, torch.device):
        device = device.index or 0
    key = (int(device), int(H), owner)
    eng = _engines.get(key)
    if eng is None or eng.max_samples < max_samples:
        prec = precision or (eng.precision if eng is not None else "fp32")
        if eng is not None:
            eng.__del__()
        eng = Engine(device, max_samples, H, prec)
        _engines[key] = eng
    elif precision is not None and eng.precision != precision:
        eng.set_precision(precision)
    return eng
