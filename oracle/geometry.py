"""CPU restatement of the PE / OO geometry that follows the action argmax.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/code/utils.py:70-81 (`global_position`), :316-366 (`get_best_grasp_angle`) and :370-540
(`get_best_suction_angle`), restructured the way csrc/geometry.cu evaluates them (segments + rounds) so that the two can
be compared step by step.  Pinned on tests/golden/golden_r02.json (recorded from the unmodified reference functions).
"""
import math

import numpy as np


def global_position(pix, A, K, P, depth):
    """utils.py:70-81: heightmap pixel (_, row, col) -> camera pixel (int() truncation) -> depth -> robot frame."""
    row, col = float(pix[1]), float(pix[2])
    den = col * A[2, 0] + row * A[2, 1] + A[2, 2]
    px = int((col * A[0, 0] + row * A[0, 1] + A[0, 2]) / den)
    py = int((col * A[1, 0] + row * A[1, 1] + A[1, 2]) / den)
    z = depth[py][px]
    cam = np.asarray([(px - K[0][2]) * (z / K[0][0]), (py - K[1][2]) * (z / K[1][1]), z])
    return np.asarray(P)[0:3, 0:3] @ cam + np.asarray(P)[0:3, 3]


def _centre(box, i):
    return [0, int(sum(box[i][j][1] for j in range(4)) / 4), int(sum(box[i][j][0] for j in range(4)) / 4)]


def grasp_angle(is_pe, box, best, A, K, P, depth):
    """utils.py:316-366."""
    c = global_position(_centre(box, best), A, K, P, depth)
    angle, open_d = 0, 2.0
    if is_pe:
        q = [global_position([0, int(box[best][i][1]), int(box[best][i][0])], A, K, P, depth) for i in range(4)]
        d01 = math.sqrt((q[0][0] - q[1][0]) ** 2 + (q[0][1] - q[1][1]) ** 2)
        d12 = math.sqrt((q[2][0] - q[1][0]) ** 2 + (q[2][1] - q[1][1]) ** 2)
        if d01 > d12:
            open_d, a, b, d = d12 * min(1.2, d01 / d12), q[0], q[1], d01
        else:
            open_d, a, b, d = d01 * min(1.2, d12 / d01), q[2], q[1], d12
        if a[1] == b[1]:
            angle = 0
        elif a[1] > b[1]:
            angle = math.acos((a[0] - b[0]) / d)
        else:
            angle = math.acos((b[0] - a[0]) / d)
    return c, angle, open_d


def _segments(val):
    """utils.py:467-476: run-length segments; the last one is appended only if it did not start at bin 359."""
    seg, start, cur = [], 0, val[0]
    for i in range(360):
        if val[i] != cur:
            seg.append((cur, start, i - 1))
            cur, start = val[i], i
        if i == 359 and start != i:
            seg.append((cur, start, i))
    return seg


def _vote(ov, best):
    val = np.ones(360)
    for i in range(len(ov)):
        if i == best or ov[i][2] == 1.0:
            continue
        a0, a1 = int(180 * ov[i][0] / np.pi), int(180 * ov[i][1] / np.pi)
        if abs(ov[i][0] - ov[i][1]) <= np.pi:
            val[a0:a1] *= ov[i][2]
        else:
            val[:a0] *= ov[i][2]
            val[a1:] *= ov[i][2]
    return val


def suction_angle(is_oo, n, cter, box, best, A, K, P, depth):
    """utils.py:370-540."""
    c = global_position(_centre(box, best), A, K, P, depth)
    if not is_oo:
        return c, np.deg2rad(0)
    ctr = [global_position([0, cter[i][1], cter[i][0]], A, K, P, depth) for i in range(n)]
    height = [max([ctr[i][2]] + [global_position([0, int(box[i][j][1]), int(box[i][j][0])], A, K, P, depth)[2] for j in range(4)])
              for i in range(n)]
    dist = [math.sqrt((ctr[i][0] - ctr[best][0]) ** 2 + (ctr[i][1] - ctr[best][1]) ** 2) for i in range(n)]
    cx, cy = cter[best][0], cter[best][1]
    ov = np.ones((n, 3))
    for o in range(n):
        if o == best:
            continue
        ap = []
        for k in range(4):
            x, y = box[o][k][0], box[o][k][1]
            a = 0.0
            if x == cx:
                a = np.pi if y > cy else 0.0
            if y == cy:
                a = np.pi / 2 if x < cx else 3 * np.pi / 2
            if x < cx:
                if y < cy:
                    a = math.atan((cx - x) / (cy - y))
                elif y > cy:
                    a = np.pi / 2 + math.atan((y - cy) / (cx - x))
            if x > cx:
                if y < cy:
                    a = 3 * np.pi / 2 + math.atan((cy - y) / (x - cx))
                elif y > cy:
                    a = np.pi + math.atan((x - cx) / (y - cy))
            ap.append(a)
        amax = 0
        for i in range(3):
            for j in range(i + 1, 4):
                d = min(abs(ap[i] - ap[j]), 2 * np.pi - abs(ap[i] - ap[j]))
                if d > amax:
                    amax, ov[o][0], ov[o][1] = d, min(ap[i], ap[j]), max(ap[i], ap[j])
    for i in range(n):
        ov[i][2] = math.exp(-max(0., height[i] - height[best]) / max(0.001, dist[i]))
    vals = sorted(set(list(ov[:, 2]) + [1.0]), reverse=True)
    val = _vote(ov, best)
    seg = _segments(val)
    selected = None
    for rnd in range(len(vals)):
        if min(s[0] for s in seg) >= 0.95:
            selected = 0.
            break
        if val[1] == val[359] and seg[0][0] >= 1.0:
            left, right = seg[0][2], seg[-1][2] - seg[-1][1]
            if left + right >= 45:
                selected = left - (left + right) // 2 if left > right else seg[-1][1] + (left + right) // 2
                break
        best_len, best_mid = -1, 0
        for v, lo, hi in seg:
            if v >= 1.0 and hi - lo >= 45 and hi - lo >= best_len:
                best_len, best_mid = hi - lo, (lo + hi) // 2
        if best_len >= 0:
            selected = best_mid
            break
        for i in range(n):
            if abs(ov[i][2] - vals[rnd + 1]) < 0.001:
                ov[i][2] = 1.
        val = _vote(ov, best)
        seg = _segments(val)
    return c, np.deg2rad(selected)
