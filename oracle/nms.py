"""CPU restatement of NMS.py_cpu_nms.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/code/NMS.py:8-59: strict area filter
`min_area < (x2-x1)*(y2-y1) < max_area` (:19-21), areas with the +1 pixel
convention (:23), greedy suppression in INDEX order (no score sort, :28-40),
suppress when IoU > co_thresh.  `pred_score` only supplies the length (:14).
All arithmetic stays in the dtype of `boxes` (float32 from the detector).
Pinned against the reference in tests/test_oracle_golden.py (tests/golden/nms.npz).
"""
import numpy as np


def nms(boxes, pred_score, co_thresh, min_area, max_area):
    boxes = np.asarray(boxes)
    n = len(pred_score)
    x1, y1, x2, y2 = boxes[:, 0, 0], boxes[:, 0, 1], boxes[:, 1, 0], boxes[:, 1, 1]
    cand = []
    for i in range(n):
        area = (x2[i] - x1[i]) * (y2[i] - y1[i])
        if area > min_area and area < max_area:
            cand.append(i)
    areas = (x2 - x1 + 1) * (y2 - y1 + 1)
    alive = list(cand)
    keep = []
    while alive:
        i = alive[0]
        keep.append(int(i))
        rest = []
        for j in alive[1:]:
            xx1 = max(x1[i], x1[j])
            yy1 = max(y1[i], y1[j])
            xx2 = min(x2[i], x2[j])
            yy2 = min(y2[i], y2[j])
            w = max(boxes.dtype.type(0.0), xx2 - xx1 + 1)
            h = max(boxes.dtype.type(0.0), yy2 - yy1 + 1)
            inter = w * h
            ovr = inter / (areas[i] + areas[j] - inter)
            if ovr <= co_thresh:
                rest.append(j)
        alive = rest
    return keep
