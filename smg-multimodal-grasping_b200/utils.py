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


class CrossEntropyLoss2d(nn.Module):
    """NLLLoss(log_softmax) over [1,3,1,1] logits with class weights (code/utils.py:306-313)."""

    def __init__(self, weight=None, size_average=True):
        super().__init__()
        self.weight = weight
        self.reduction = "mean" if size_average else "sum"

    def forward(self, inputs, targets):
        return F.nll_loss(F.log_softmax(inputs, dim=1), targets, weight=self.weight, reduction=self.reduction)
