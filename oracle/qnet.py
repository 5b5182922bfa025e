"""CPU restatement of the reference Q-network path.  TEST INFRASTRUCTURE ONLY.

Follows (file:line relative to /root/reference):
  * trainer pre-processing            code/trainer.py:162-191
  * nearest rotation                  code/models.py:371-382 (F.affine_grid + F.grid_sample)
  * DenseNet-121 `.features`          torchvision/models/densenet.py (third party, torchvision
                                      0.26.0 here; the reference does not pin it) as called at
                                      code/models.py:384-385
  * heads                             code/models.py:316-343 (RL), :28-55 (reactive)
  * forward branches                  code/models.py:361-586 / :72-296
  * losses                            code/trainer.py:284-299 (reactive CE), :345-348 (Huber)

Everything is written with functional torch ops on a plain state_dict so it can
run in float32 (the reference's precision) or float64 (error yard-stick).
BatchNorm always uses the statistics of the one sample in flight (the
reference never leaves train mode, code/trainer.py:95).

Parity pin: tests/test_oracle_golden.py checks this file against outputs of the
unmodified reference modules (tests/golden/, made by tests/golden/make_golden.py).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

BLOCK_CONFIG = (6, 12, 24, 16)  # torchvision densenet121
GROWTH = 32
BN_EPS = 1e-5

STYLE_TRUNK = {0: "grasp_depth_trunk", 1: "suction_depth_trunk", 2: "gs_depth_trunk"}
# style 2 feeds the gs trunk features to the *suction* head (code/models.py:434,507,582)
STYLE_HEAD = {0: ("graspnet_val", "grasp-val"), 1: ("suctionnet_val", "suction-val"),
              2: ("suctionnet_val", "suction-val")}


# ----------------------------------------------------------------------------
# trainer pre-processing (code/trainer.py:162-191)
# ----------------------------------------------------------------------------
def preprocess(depth_heightmap, mean=0.01, std=0.03):
    """224x224 float64 heightmap -> float32 [1,3,640,640] network input.

    zoom x2 order 0 (== np.repeat on both axes, code/trainer.py:165), zero pad to
    ceil(448*sqrt(2)/32)*32 = 640 (code/trainer.py:169-173), replicate to 3
    channels and normalise in float64 (code/trainer.py:176-185), cast to float32
    (code/trainer.py:188).  The published mean/std literals are 0/0 (-> NaN,
    SURVEY.md section 0.4); mean 0.01 / std 0.03 is the harness substitution.
    """
    d = np.asarray(depth_heightmap, dtype=np.float64)
    d2 = np.repeat(np.repeat(d, 2, axis=0), 2, axis=1)
    diag = float(d2.shape[0]) * np.sqrt(2)
    diag = np.ceil(diag / 32) * 32
    pad = int((diag - d2.shape[0]) / 2)
    d2 = np.pad(d2, pad, "constant", constant_values=0)
    x = (d2 - mean) / std
    x = np.stack([x, x, x], axis=0)[None]
    return torch.from_numpy(x.astype(np.float32))


# ----------------------------------------------------------------------------
# rotation (code/models.py:371-382)
# ----------------------------------------------------------------------------
def rotation_theta(rotate_idx, num_rotations):
    """code/models.py:372-376: float64 angle -> 2x3 matrix -> float32."""
    rotate_theta = np.radians(rotate_idx * (360 / num_rotations))
    m = np.asarray([[np.cos(-rotate_theta), np.sin(-rotate_theta), 0],
                    [-np.sin(-rotate_theta), np.cos(-rotate_theta), 0]])
    return m.astype(np.float32)


def rotate_nearest(x, rotate_idx, num_rotations):
    """F.affine_grid(align_corners=True) + F.grid_sample(mode='nearest', zero padding)."""
    theta = torch.from_numpy(rotation_theta(rotate_idx, num_rotations))[None].to(device=x.device, dtype=x.dtype)
    grid = F.affine_grid(theta, list(x.shape), align_corners=True)
    return F.grid_sample(x, grid, mode="nearest", align_corners=True)


def rotate_index_map(H, rotate_idx, num_rotations):
    """Explicit float32 restatement of the index arithmetic behind rotate_nearest.

    Returns int32 [H,H] source linear indices (-1 = outside -> 0).  Follows
    torch: base grid linspace(-1,1,H) (symmetric evaluation), grid = base @ theta^T
    in float32, unnormalise ((g+1)/2)*(H-1), nearbyint (ties to even), bounds test.
    The float32 accumulation order of the K=3 product is the one that matches
    torch CPU bit-for-bit in tests/test_oracle_golden.py::test_rotate_index_map.
    """
    th = rotation_theta(rotate_idx, num_rotations)
    lin = linspace_sym(H)
    bx = lin[None, :].repeat(H, axis=0)
    by = lin[:, None].repeat(H, axis=1)
    gx = affine_dot(bx, by, th[0])
    gy = affine_dot(bx, by, th[1])
    half = np.float32((H - 1)) / np.float32(2)
    ix = (gx + np.float32(1)) * half
    iy = (gy + np.float32(1)) * half
    ixn = np.rint(ix).astype(np.int64)
    iyn = np.rint(iy).astype(np.int64)
    ok = (ixn >= 0) & (ixn < H) & (iyn >= 0) & (iyn < H)
    idx = np.where(ok, iyn * H + ixn, -1)
    return idx.astype(np.int32)


def _fma32(a, b, c):
    """float32 fused multiply-add a*b+c with a single rounding (oracle/fma_helper.c)."""
    from ._fma import fma32
    return fma32(a, b, c)


def linspace_sym(n):
    """torch.linspace(-1, 1, n) in float32 as torch CPU evaluates it (probe-verified, torch 2.11):
    step = 2f/(n-1)f; fma(step, i, -1) below the midpoint, fma(-step, n-1-i, 1) from it on."""
    step = np.float32(2) / np.float32(n - 1)
    i = np.arange(n)
    lo = _fma32(step, i.astype(np.float32), np.float32(-1))
    hi = _fma32(-step, (n - 1 - i).astype(np.float32), np.float32(1))
    return np.where(i < n // 2, lo, hi).astype(np.float32)


def affine_dot(bx, by, row):
    """float32 K=3 product of F.affine_grid's bmm as torch CPU evaluates it (probe-verified):
    acc = bx*r0 (rounded); acc = fma(by, r1, acc); acc = fma(1, r2, acc)."""
    r0, r1, r2 = np.float32(row[0]), np.float32(row[1]), np.float32(row[2])
    acc = (bx * r0).astype(np.float32)
    acc = _fma32(by, r1, acc)
    return _fma32(np.float32(1), r2, acc)


# ----------------------------------------------------------------------------
# DenseNet-121 features, functional, train-mode BN
# ----------------------------------------------------------------------------
def _bn(x, sd, name, relu):
    y = F.batch_norm(x, None, None, sd[name + ".weight"], sd[name + ".bias"],
                     training=True, momentum=0.0, eps=BN_EPS)
    return F.relu(y) if relu else y


def densenet_features(sd, prefix, x, taps=None):
    """`densenet121().features(x)` with per-sample train-mode BN.

    sd: state_dict (or dict of tensors) of the whole net; prefix e.g.
    'grasp_depth_trunk.features.'.  x: [N,3,H,W]; N>1 is evaluated sample by
    sample so BN statistics stay per sample like the reference's batch-1 calls.
    taps: optional dict that receives intermediate activations of sample 0.
    """
    if x.shape[0] > 1:
        return torch.cat([densenet_features(sd, prefix, x[i:i + 1], taps if i == 0 else None)
                          for i in range(x.shape[0])], 0)
    p = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    y = F.conv2d(x, p["conv0.weight"], stride=2, padding=3)
    if taps is not None:
        taps["conv0"] = y
    y = _bn(y, p, "norm0", True)
    y = F.max_pool2d(y, 3, 2, 1)
    if taps is not None:
        taps["pool0"] = y
    for b, nl in enumerate(BLOCK_CONFIG, start=1):
        for l in range(1, nl + 1):
            q = "denseblock%d.denselayer%d." % (b, l)
            t = _bn(y, p, q + "norm1", True)
            t = F.conv2d(t, p[q + "conv1.weight"])
            if taps is not None and l == 1:
                taps["b%d_l1_conv1" % b] = t
            t = _bn(t, p, q + "norm2", True)
            t = F.conv2d(t, p[q + "conv2.weight"], padding=1)
            y = torch.cat([y, t], 1)
        if taps is not None:
            taps["block%d" % b] = y
        if b < 4:
            q = "transition%d." % b
            y = _bn(y, p, q + "norm", True)
            y = F.conv2d(y, p[q + "conv.weight"])
            y = F.avg_pool2d(y, 2, 2)
            if taps is not None:
                taps["trans%d" % b] = y
    y = _bn(y, p, "norm5", False)  # ReLU lives in DenseNet.forward, which the reference never calls
    return y


def head(sd, head_attr, head_name, feat, taps=None):
    """BN2048-ReLU-1x1(2048->64)-BN64-ReLU-20x20 valid conv (code/models.py:316-343, :28-55)."""
    q = head_attr + "." + head_name + "-"
    y = _bn(feat, sd, q + "norm0", True)
    y = F.conv2d(y, sd[q + "conv0.weight"])
    if taps is not None:
        taps["head_conv0"] = y
    y = _bn(y, sd, q + "norm1", True)
    return F.conv2d(y, sd[q + "conv1.weight"])


def q_forward(sd, x_scene, x_mask, style, rotations, num_rotations, taps=None):
    """Q (or logits) for the given rotation indices.

    Restates one loop body of code/models.py:371-389: rotate the scene, run the
    style's trunk on the rotated scene and on the *un-rotated* masked scene,
    concatenate and apply the style's head.  Returns a list of [1,C,1,1].
    """
    trunk = STYLE_TRUNK[style] + ".features."
    head_attr, head_name = STYLE_HEAD[style]
    out = []
    f_m = None
    for r in rotations:
        rot = rotate_nearest(x_scene, r, num_rotations)
        f_s = densenet_features(sd, trunk, rot, taps)
        if f_m is None or torch.is_grad_enabled():
            f_m = densenet_features(sd, trunk, x_mask)  # identical for every rotation
        feat = torch.cat((f_s, f_m), dim=1)
        if taps is not None:
            taps["feat"] = feat
        out.append(head(sd, head_attr, head_name, feat, taps))
        taps = None
    return out


def model_forward(sd, x_scene, x_mask, style=0, is_volatile=False, specific_rotation=-1,
                  gnum_rotations=1, snum_rotations=1):
    """`reinforcement_net.forward` / `reactive_net.forward` (code/models.py:361-586, :72-296).

    Branches: (volatile, -1) -> list over all rotations (style 2: rotation 0 only,
    code/models.py:418); (volatile, r) -> tensor; (grad, r) -> tensor with grad.
    specific_rotation angles always use gnum_rotations, also for suction
    (code/models.py:469,545); style 2 always uses rotation 0 (code/models.py:491,567).
    """
    if is_volatile and specific_rotation == -1:
        with torch.no_grad():
            if style == 0:
                return q_forward(sd, x_scene, x_mask, 0, range(gnum_rotations), gnum_rotations)
            if style == 1:
                return q_forward(sd, x_scene, x_mask, 1, range(snum_rotations), snum_rotations)
            return q_forward(sd, x_scene, x_mask, 2, [0], gnum_rotations)
    rot = 0 if style == 2 else specific_rotation
    if is_volatile:
        with torch.no_grad():
            return q_forward(sd, x_scene, x_mask, style, [rot], gnum_rotations)[0]
    return q_forward(sd, x_scene, x_mask, style, [rot], gnum_rotations)[0]


# ----------------------------------------------------------------------------
# losses (code/trainer.py:284-299, :345-348; code/utils.py:306-313)
# ----------------------------------------------------------------------------
def huber_loss(q, label_value):
    """Hand-written Huber, delta 1, on the scalar Q (code/trainer.py:345-348)."""
    d = q - label_value
    if abs(float(d)) < 1:
        return 0.5 * (d ** 2)
    return abs(d) - 0.5


def reactive_loss(logits, label_value):
    """CrossEntropyLoss2d(weight=[1,1,0]) on [1,3,1,1] logits (code/trainer.py:40-45,296-299)."""
    w = torch.tensor([1.0, 1.0, 0.0], dtype=logits.dtype)
    target = torch.full((1, 1, 1), int(label_value), dtype=torch.long)
    return F.nll_loss(F.log_softmax(logits.view(1, 3, 1, 1), dim=1), target, weight=w, reduction="mean")


def backprop_grads(sd, x_scene, x_mask, style, rotation, label_value, method="reinforcement",
                   gnum_rotations=1):
    """Loss and gradients of one `Trainer.backprop` step (code/trainer.py:278-384).

    Returns (loss float, {state_dict key: grad}) for the trunk + head the sample touches.
    """
    trunk = STYLE_TRUNK[style] + ".features."
    head_attr, _ = STYLE_HEAD[style]
    leaves = {}
    for k, v in sd.items():
        if (k.startswith(trunk) or k.startswith(head_attr + ".")) and v.is_floating_point() \
                and "running_" not in k:
            leaves[k] = v.detach().clone().requires_grad_(True)
    sd2 = dict(sd)
    sd2.update(leaves)
    with torch.enable_grad():
        out = model_forward(sd2, x_scene, x_mask, style, False, rotation, gnum_rotations)
        if method == "reinforcement":
            loss = huber_loss(out[0, 0, 0, 0], label_value)
        else:
            loss = reactive_loss(out, label_value)
        loss = loss.sum()
        loss.backward()
    return float(loss), {k: v.grad for k, v in leaves.items() if v.grad is not None}


def adam_step(param, grad, m, v, step, lr=1e-4, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.Adam single-tensor update (code/trainer.py:99), step counted from 1."""
    m = b1 * m + (1 - b1) * grad
    v = b2 * v + (1 - b2) * grad * grad
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    return param - (lr / bc1) * m / denom, m, v
