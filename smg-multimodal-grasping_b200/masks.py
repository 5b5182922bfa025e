"""The weight-free half of the detector front end (reference: code/masks.py:31-85).

The reference runs a pretrained torchvision Mask R-CNN on the CPU (code/masks.py:15-16; needs downloaded weights, out of
scope here) and then post-processes its raw output: score filter, bilinear resize of the soft masks from the 448x448
detector grid to the 224x224 heightmap grid, box halving and the in-order IoU NMS.  `postprocess_detections` is that
second half on the GPU (resize kernel + `NMS.py_cpu_nms`), taking the detector's raw `pred` dict entries as inputs, so a
caller that owns a detector can plug it in front of the Q pass.  The contour / min-area-rectangle box extraction of
`instance_segmentation` (code/masks.py:137-161, cv2.findContours + cv2.minAreaRect) is not rebuilt: see DESIGN.md.
"""
import numpy as np
import torch

from . import NMS
from . import engine as _engine


def resize_soft_masks(masks, size_out=224, device=None):
    """F.interpolate(masks, size=[224,224], mode="bilinear", align_corners=True) of [n,1,448,448] / [n,448,448] float32 masks
    (code/masks.py:51) -> numpy [n,224,224]."""
    eng = _engine.stateless_engine(torch.cuda.current_device() if device is None else device)
    m = torch.as_tensor(masks, dtype=torch.float32)
    m = m.reshape(-1, m.shape[-2], m.shape[-1])
    return eng.resize_masks(m, size_out).cpu().numpy()


def postprocess_detections(pred_masks, pred_boxes, pred_scores, threshold, image_size=448, device=None):
    """What `get_prediction` does after the detector call (code/masks.py:36-83).

    pred_masks [n,1,s,s] float32 soft masks, pred_boxes [n,4] (x1,y1,x2,y2), pred_scores [n] sorted descending - the entries
    of torchvision's detection output.  Returns (masks_initial bool [k,s,s], masks float32 [k,s/2,s/2], boxes [k,2,2] in the
    half-resolution grid, kept indices, number) like the reference (all None / 0 when no score passes `threshold`)."""
    scores = [float(v) for v in np.asarray(pred_scores).reshape(-1)]
    passing = [i for i, v in enumerate(scores) if v > threshold]
    if not passing:
        return None, None, None, [], 0
    last = passing[-1]                                       # the reference keeps everything up to the LAST passing score
    m = torch.as_tensor(pred_masks, dtype=torch.float32)
    m = m.reshape(-1, m.shape[-2], m.shape[-1])
    half = image_size // 2
    soft = resize_soft_masks(m, half, device)[:last + 1]
    initial = (m > 0.5).numpy()[:last + 1]
    b = np.asarray(pred_boxes, dtype=np.float32).reshape(-1, 4)[:last + 1]
    boxes = np.stack([b[:, 0:2], b[:, 2:4]], axis=1) / 2     # [(x1,y1),(x2,y2)] halved (code/masks.py:69-71)
    area = image_size * image_size / 4
    keep = NMS.py_cpu_nms(boxes, scores[:last + 1], 0.40, area / 60, area / 5, device=device)
    return initial[keep], soft[keep], boxes[keep], keep, len(keep)
