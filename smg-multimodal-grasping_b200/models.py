"""Drop-in `reinforcement_net` / `reactive_net` whose forward runs on libsmg_b200.so.

Boundary kept from the reference:
  * class names, constructor `(use_cuda)`, attribute names (`suction_depth_trunk`,
    `grasp_depth_trunk`, `gs_depth_trunk`, `suctionnet_val`, `graspnet_val`, `gsnet_val`,
    `gnum_rotations`, `snum_rotations`, `gra_prob`, `suc_prob`, `gs_prob`) and the 2217
    state_dict keys (code/models.py:15-69, :301-358), so reference snapshots load unchanged;
  * `forward(input_depth_data, m_input_depth_data, style=0, is_volatile=False,
    specific_rotation=-1)` with the reference's three branches and return types
    (code/models.py:361-586, :72-296).
The parameter containers are ordinary torch modules (torchvision's DenseNet-121 tree is
used for storage and key names only; its forward is never called).  The arithmetic is
done by the CUDA library; without it (or without a GPU) forward raises - no fallback.
"""
import weakref
from collections import OrderedDict

import torch
import torch.nn as nn
import torchvision

from . import engine as _engine

# running-stat momentum of nn.BatchNorm2d (torch default), used for the EMA side effect
_BN_MOMENTUM = 0.1


def _make_head(prefix, n_out):
    return nn.Sequential(OrderedDict([
        (prefix + '-norm0', nn.BatchNorm2d(2048)),
        (prefix + '-relu0', nn.ReLU(inplace=True)),
        (prefix + '-conv0', nn.Conv2d(2048, 64, kernel_size=1, stride=1, bias=False)),
        (prefix + '-norm1', nn.BatchNorm2d(64)),
        (prefix + '-relu1', nn.ReLU(inplace=True)),
        (prefix + '-conv1', nn.Conv2d(64, n_out, kernel_size=20, stride=1, bias=False)),
    ]))


class _SmgNet(nn.Module):
    """Shared implementation; `N_OUT` = 1 (Q value) or 3 (class logits)."""

    N_OUT = 1
    precision = "fp32"          # "fp32" | "tf32" | "bf16": arithmetic mode of the convolutions
    update_running_stats = True  # reproduce BatchNorm's running_mean/var side effect

    def __init__(self, use_cuda):
        super().__init__()
        self.use_cuda = use_cuda
        # same construction order as the reference so that a given torch seed yields the same
        # random-init weights (code/models.py:308-353); no network here, so no ImageNet weights
        self.suction_depth_trunk = torchvision.models.densenet.densenet121(weights=None)
        self.grasp_depth_trunk = torchvision.models.densenet.densenet121(weights=None)
        self.gs_depth_trunk = torchvision.models.densenet.densenet121(weights=None)
        self.gnum_rotations = 1
        self.snum_rotations = 1
        self.suctionnet_val = _make_head('suction-val', self.N_OUT)
        self.graspnet_val = _make_head('grasp-val', self.N_OUT)
        # the reference re-uses the 'grasp-val-*' names for the ES head (code/models.py:336-343)
        self.gsnet_val = _make_head('grasp-val', self.N_OUT)
        for name, m in self.named_modules():
            if 'suction-' in name or 'grasp-' in name or 'gs-' in name:
                if isinstance(m, nn.Conv2d):
                    nn.init.kaiming_normal_(m.weight.data)
                elif isinstance(m, nn.BatchNorm2d):
                    m.weight.data.fill_(1)
                    m.bias.data.zero_()
        self.gra_prob = []
        self.suc_prob = []
        self.gs_prob = []

    # ------------------------------------------------------------------ engine plumbing
    def _device_index(self):
        p = next(self.parameters())
        if p.is_cuda:
            return p.device.index
        return torch.cuda.current_device()

    def _engine(self, n_samples, style=None):
        if getattr(self, "_finalizer_owner", None) != id(self):  # also true for a deepcopy of a live model
            object.__setattr__(self, "_finalizer_owner", id(self))
            weakref.finalize(self, _engine.drop_engine, id(self))
        eng = _engine.get_engine(self._device_index(), max(n_samples, 18), 640, self.precision, owner=id(self))
        eng.sync_weights(self, style=style)
        return eng

    def _bn_modules(self, trunk):
        return [m for m in trunk.features.modules() if isinstance(m, nn.BatchNorm2d)]

    @torch.no_grad()
    def _apply_running_stats(self, trunk, mean, var, order):
        """BatchNorm2d train-mode side effect: running = (1-m)*running + m*batch for each trunk call in
        `order` (sample indices; the reference calls trunk(scene_r) then trunk(mask) per rotation)."""
        k = len(order)
        # the per-sample weights depend on the call pattern only: built once per pattern and kept on the device (a fresh
        # torch.tensor(..., device=cuda) is a synchronous pageable copy that waits for the whole pass in front of it)
        cache = self.__dict__.setdefault("_bn_weight_cache", {})
        key = ("trunk", mean.shape[0], tuple(order), str(mean.device))
        w = cache.get(key)
        if w is None:
            wl = [0.0] * mean.shape[0]
            for i, s in enumerate(order):
                wl[s] += _BN_MOMENTUM * (1 - _BN_MOMENTUM) ** (k - 1 - i)
            if len(cache) > 64:
                cache.clear()
            w = cache[key] = torch.tensor(wl, dtype=torch.float64, device=mean.device)
        self._apply_running_sums(trunk, (w[:, None] * mean.double()).sum(0), (w[:, None] * var.double()).sum(0), k)

    @torch.no_grad()
    def _apply_running_sums(self, trunk, bm, bv, k):
        """running = (1-m)^k running + (weighted sums of the k passes' batch means / biased variances, float64 [C])."""
        decay = (1 - _BN_MOMENTUM) ** k
        mean = bm
        mods = self._bn_modules(trunk)
        sizes = [m.num_features for m in mods]
        # unbiased variance: n/(n-1) with n = H*W of that layer
        key = (id(trunk), mean.device)
        cache = self.__dict__.setdefault("_bn_unbias", {})
        if key not in cache:
            cache[key] = torch.cat([torch.full((c,), n / (n - 1.0), dtype=torch.float64)
                                    for c, n in zip(sizes, _bn_counts(640))]).to(mean.device)
        bm = bm.float()
        bv = (bv * cache[key]).float()
        rm, rv = [m.running_mean for m in mods], [m.running_var for m in mods]
        # a handful of multi-tensor launches instead of ~500 tiny ones
        torch._foreach_mul_(rm, decay)
        torch._foreach_add_(rm, list(torch.split(bm, sizes)))
        torch._foreach_mul_(rv, decay)
        torch._foreach_add_(rv, list(torch.split(bv, sizes)))
        torch._foreach_add_([m.num_batches_tracked for m in mods], k)

    @torch.no_grad()
    def _apply_head_running_stats(self, head, trunk, var, pairs, head_bn1):
        """Running statistics of the head's two BatchNorm2d for the head calls listed in `pairs`
        [(scene sample, mask sample), ...] in the reference's call order (code/models.py:386-387).
        BN(2048) normalises cat(norm5(x_s), norm5(x_m)): its batch mean is norm5.bias and its batch variance
        gamma5^2 var/(var+eps) of the respective sample; BN(64)'s statistics come from the head kernel."""
        bns = [m for m in head.children() if isinstance(m, nn.BatchNorm2d)]
        norm0, norm1 = bns[0], bns[1]
        norm5 = trunk.features.norm5
        k = len(pairs)
        cache = self.__dict__.setdefault("_bn_weight_cache", {})
        key = ("head", tuple(pairs), str(var.device))
        ent = cache.get(key)
        if ent is None:
            if len(cache) > 64:
                cache.clear()
            ent = cache[key] = (torch.tensor([_BN_MOMENTUM * (1 - _BN_MOMENTUM) ** (k - 1 - i) for i in range(k)],
                                             dtype=torch.float64, device=var.device),
                                torch.tensor([p[0] for p in pairs], device=var.device),
                                torch.tensor([p[1] for p in pairs], device=var.device))
        w, si, mi = ent
        decay = (1 - _BN_MOMENTUM) ** k
        n = 400.0
        v5 = var[:, -1024:].double()
        var_z = norm5.weight.double() ** 2 * v5 / (v5 + 1e-5)                       # [samples, 1024]
        bv = torch.cat([(w[:, None] * var_z[si]).sum(0), (w[:, None] * var_z[mi]).sum(0)]) * (n / (n - 1.0))
        bm = torch.cat([norm5.bias.double(), norm5.bias.double()]) * w.sum()
        norm0.running_mean.mul_(decay).add_(bm.to(norm0.running_mean))
        norm0.running_var.mul_(decay).add_(bv.to(norm0.running_var))
        h1 = head_bn1.double()                                                       # [k, 2, 64] in call order
        norm1.running_mean.mul_(decay).add_((w[:, None] * h1[:, 0]).sum(0).to(norm1.running_mean))
        norm1.running_var.mul_(decay).add_(((w[:, None] * h1[:, 1]).sum(0) * (n / (n - 1.0))).to(norm1.running_var))
        norm0.num_batches_tracked += k
        norm1.num_batches_tracked += k

    # ------------------------------------------------------------------ forward
    def forward(self, input_depth_data, m_input_depth_data, style=0, is_volatile=False, specific_rotation=-1):
        if not is_volatile:
            from .autograd import q_forward_with_grad
            return q_forward_with_grad(self, input_depth_data, m_input_depth_data, style, specific_rotation)
        if specific_rotation == -1:
            if style == 0:
                rots, nrot = list(range(self.gnum_rotations)), self.gnum_rotations
            elif style == 1:
                rots, nrot = list(range(self.snum_rotations)), self.snum_rotations
            else:
                rots, nrot = [0], self.gnum_rotations          # code/models.py:418
        else:
            # the reference derives the angle from gnum_rotations for every style and pins the ES
            # primitive to rotation 0 (code/models.py:469,491,545)
            rots, nrot = [0 if style == 2 else int(specific_rotation)], self.gnum_rotations
        q = self._q(input_depth_data, m_input_depth_data, style, rots, nrot)  # [n_rot, C]
        outs = [q[i].view(1, self.N_OUT, 1, 1) for i in range(len(rots))]
        return outs if specific_rotation == -1 else outs[0]

    @torch.no_grad()
    def _q(self, scene, mask, style, rots, nrot):
        eng = self._engine(len(rots) + 1, style)
        scene = scene.reshape(3, 640, 640)
        mask = mask.reshape(1, 3, 640, 640)
        if self.update_running_stats:
            q, mean, var = eng.qforward(style, scene, mask, rots, nrot, want_bn_stats=True)
            trunk = getattr(self, _engine.TRUNK_ATTRS[_engine.STYLE_ROUTE[style][0]])
            order = []
            for i in range(len(rots)):
                order += [i, len(rots)]
            self._apply_running_stats(trunk, mean, var, order)
            head = getattr(self, _engine.HEAD_ATTRS[_engine.STYLE_ROUTE[style][1]])
            self._apply_head_running_stats(head, trunk, var, [(i, len(rots)) for i in range(len(rots))],
                                           eng.head_bn_stats(len(rots)))
        else:
            q = eng.qforward(style, scene, mask, rots, nrot)
        return q[0]


def _bn_counts(H):
    """H*W seen by each of the 121 trunk BatchNorm layers, in module order."""
    counts = [(H // 2) ** 2]
    hw = H // 4
    for b, nl in enumerate((6, 12, 24, 16)):
        counts += [hw * hw] * (2 * nl + 1)  # norm1/norm2 per layer + transition norm / norm5
        hw //= 2
    return counts


class reactive_net(_SmgNet):
    """Reactive policy: 3-class logits per (scene, mask, rotation) (code/models.py:15-296)."""
    N_OUT = 3


class reinforcement_net(_SmgNet):
    """DRL policy: scalar Q per (scene, mask, rotation) (code/models.py:301-586)."""
    N_OUT = 1
