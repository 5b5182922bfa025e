"""Drop-in `py_cpu_nms` (reference: code/NMS.py:8-59) running the K12 kernel on the GPU."""
import numpy as np
import torch

from . import engine as _engine


def py_cpu_nms(boxes, pred_score, co_thresh, min_area, max_area, device=None):
    """boxes [N,2,2] float32 ((x1,y1),(x2,y2)); returns the kept indices as a Python list (index order)."""
    n = len(pred_score)
    if n == 0:
        return []
    eng = _engine.stateless_engine(torch.cuda.current_device() if device is None else device)
    b = torch.from_numpy(np.ascontiguousarray(np.asarray(boxes)[:n], dtype=np.float32))
    keep, cnt = eng.nms(b, co_thresh, min_area, max_area)
    k = int(cnt.item())
    return [int(v) for v in keep[:k].cpu().tolist()]
