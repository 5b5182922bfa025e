"""Drop-in pieces of the reference's utils.py that sit on the hot path.

  get_heightmap        code/utils.py:38-68   (K11 kernel for the float64 depth path)
  CrossEntropyLoss2d   code/utils.py:306-313 (reactive loss; scalar host-side math on the 3 logits)
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import engine as _engine


def get_heightmap(color_img, depth_img, cam_intrinsics, cam_pose, workspace_limits, heightmap_resolution, device=None):
    """Same 5-tuple as the reference: (color_heightmap, depth_heightmap, color_mask, depth_mask, A_htor).

    `workspace_limits` / `heightmap_resolution` are accepted and ignored exactly like the reference does.
    The depth outputs (what the Q pass consumes) come from the K11 gather kernel and are bit-identical to the
    reference's numpy + cv2 result.  The two colour warps (consumed only by Mask R-CNN / logging, outside the
    Q path) reproduce cv2's 15-bit fixed-point bilinear remap bit for bit; `color_img=None` skips them (None, None).
    """
    eng = _engine.stateless_engine(torch.cuda.current_device() if device is None else device)
    d = torch.from_numpy(np.ascontiguousarray(depth_img, dtype=np.float64))
    o224, o448, A = eng.heightmap(d, np.asarray(cam_intrinsics, dtype=np.float64)[:3, :3], np.asarray(cam_pose, dtype=np.float64))
    c224 = c448 = None
    if color_img is not None:
        img = torch.from_numpy(np.ascontiguousarray(color_img, dtype=np.uint8).reshape(480, 640, 3))
        c224, c448 = (t.cpu().numpy() for t in eng.heightmap_color(img))
    return c224, o224.cpu().numpy(), c448, o448.cpu().numpy(), A


def _depth_on_device(depth_img, eng):
    """The camera depth image as a float64 device tensor; the last upload is kept, the step loop passes the same array to
    several geometry calls in a row (code/main.py:245-262)."""
    key = (id(depth_img), getattr(depth_img, "shape", None))
    cache = _depth_on_device.__dict__
    if cache.get("key") != key or cache.get("dev") != eng.device:
        cache["key"], cache["dev"] = key, eng.device
        cache["val"] = torch.from_numpy(np.ascontiguousarray(depth_img, dtype=np.float64)).to(eng.device)
        cache["ref"] = depth_img       # keeps id() unique while cached
    return cache["val"]


def global_position(pix_mask_position, A_htor, cam_intrinsics, cam_pose, depth_img, device=None):
    """Heightmap pixel (_, row, col) -> robot-frame xyz (reference: code/utils.py:70-81)."""
    eng = _engine.stateless_engine(torch.cuda.current_device() if device is None else device)
    out = eng.geometry(0, _depth_on_device(depth_img, eng), A_htor, cam_intrinsics, cam_pose, pix=pix_mask_position)
    return out[:3].copy()


def get_best_grasp_angle(is_pe, box_mask_cors, bestg_id, A_htor, cam_intrinsics, cam_pose, depth_img, device=None):
    """(grasp centre xyz, jaw rotation angle, opening distance); pre-enveloping when `is_pe` (reference: code/utils.py:316-366)."""
    eng = _engine.stateless_engine(torch.cuda.current_device() if device is None else device)
    out = eng.geometry(1, _depth_on_device(depth_img, eng), A_htor, cam_intrinsics, cam_pose, boxes=box_mask_cors,
                       best=int(bestg_id[0]), flag=is_pe)
    angle = float(out[3]) if is_pe else 0
    return out[:3].copy(), angle, float(out[4])


def get_best_suction_angle(is_oo, objects_number, masks_cter, box_mask_cors, bests_id, A_htor, cam_intrinsics, cam_pose,
                           depth_img, device=None):
    """(suction centre xyz, approach direction in radians); orientation optimisation over the other objects when `is_oo`
    (reference: code/utils.py:370-612)."""
    eng = _engine.stateless_engine(torch.cuda.current_device() if device is None else device)
    boxes = np.asarray(box_mask_cors, dtype=np.float64)[:objects_number]
    out = eng.geometry(2, _depth_on_device(depth_img, eng), A_htor, cam_intrinsics, cam_pose, boxes=boxes,
                       centers=np.asarray(masks_cter, dtype=np.float64)[:objects_number], best=int(bests_id[0]), flag=is_oo)
    return out[:3].copy(), float(out[3])


class CrossEntropyLoss2d(nn.Module):
    """NLLLoss(log_softmax) over [1,3,1,1] logits with class weights (code/utils.py:306-313)."""

    def __init__(self, weight=None, size_average=True):
        super().__init__()
        self.weight = weight
        self.reduction = "mean" if size_average else "sum"

    def forward(self, inputs, targets):
        return F.nll_loss(F.log_softmax(inputs, dim=1), targets, weight=self.weight, reduction=self.reduction)
