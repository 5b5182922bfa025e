"""Seeded synthetic inputs for tests and bench (no simulator, no datasets).

Shapes follow the reference's data path: 224x224 float64 depth heightmaps
(code/utils.py:41), K soft object masks in [0,1] made the way
code/masks.py:51 makes them (bilinear 448->224 resize of a detector mask),
scene = depth * sum(masks) (code/main.py:145-151), objects on the reference's
3x4 drop grid (code/robot.py:51-61), a 640x480 camera depth image with the
reference's hard-coded intrinsics (code/robot.py:99).  SURVEY.md section 8(d).
"""
import numpy as np

CAM_INTRINSICS = np.asarray([[618.62, 0, 320], [0, 618.62, 240], [0, 0, 1]], dtype=np.float64)
WORKSPACE_LIMITS = np.asarray([[-0.724, -0.276], [-0.224, 0.224], [-0.0001, 0.4]])


def _footprints(rs, K, cluttered, size=448):
    """K binary footprints [K,size,size] of random-yaw boxes / ellipsoids on the 3x4 drop grid."""
    # 1 px of the 224 heightmap ~ 400/224 camera px ~ 2 mm; grid spacing in metres -> 448-px units
    sx, sy = (0.10, 0.10) if cluttered else (0.14, 0.10)
    px_per_m = size / 0.448
    cells = [(i, j) for i in range(3) for j in range(4)]
    order = rs.permutation(len(cells))[:K]
    yy, xx = np.mgrid[0:size, 0:size].astype(np.float64)
    out = np.zeros((K, size, size), dtype=np.float64)
    heights = rs.uniform(0.02, 0.10, size=K)
    for n, ci in enumerate(order):
        i, j = cells[ci]
        cx = size / 2 + (j - 1.5) * sx * px_per_m + rs.uniform(-6, 6)
        cy = size / 2 + (i - 1.0) * sy * px_per_m + rs.uniform(-6, 6)
        a = rs.uniform(20, 45)  # half-extent in 448 px == footprint 20-45 px in the 224 map
        b = rs.uniform(20, 45)
        yaw = rs.uniform(0, np.pi)
        u = (xx - cx) * np.cos(yaw) + (yy - cy) * np.sin(yaw)
        v = -(xx - cx) * np.sin(yaw) + (yy - cy) * np.cos(yaw)
        if rs.randint(0, 2) == 0:
            out[n] = ((np.abs(u) <= a) & (np.abs(v) <= b)).astype(np.float64)
        else:
            out[n] = ((u / a) ** 2 + (v / b) ** 2 <= 1.0).astype(np.float64)
    return out, heights


def _resize_half_bilinear(m):
    """Bilinear 2x downsample with align_corners=True (what code/masks.py:51 does to detector masks)."""
    n = m.shape[-1]
    o = n // 2
    pos = np.arange(o, dtype=np.float64) * (n - 1) / (o - 1)
    i0 = np.floor(pos).astype(np.int64)
    i1 = np.minimum(i0 + 1, n - 1)
    w = pos - i0
    rows = m[..., i0, :] * (1 - w)[:, None] + m[..., i1, :] * w[:, None]
    return rows[..., :, i0] * (1 - w) + rows[..., :, i1] * w


def make_scene(seed, num_objects=4, cluttered=False):
    """Return dict(depth [224,224] f64, masks [K,224,224] f32, scene [224,224] f64, heights [K])."""
    rs = np.random.RandomState(seed)
    fp, heights = _footprints(rs, num_objects, cluttered)
    depth448 = np.zeros((448, 448), dtype=np.float64)
    for n in range(num_objects):
        depth448 = np.maximum(depth448, fp[n] * heights[n])
    depth = depth448[::2, ::2].copy()
    masks = _resize_half_bilinear(fp).astype(np.float32)
    mask_all = masks.astype(np.float64).sum(0)
    scene = depth * mask_all  # code/main.py:145-151
    return {"depth": depth, "masks": masks, "scene": scene, "heights": heights}


def masked_scene(scene, masks, ids):
    """scene x (sum of the listed masks): code/main.py:160 (single object) / :186 (ES pair)."""
    m = np.zeros_like(scene)
    for i in ids:
        m = m + masks[i].astype(np.float64)
    return scene * m


def make_camera(seed, num_objects=10, cluttered=True):
    """640x480 depth (metres, float64), colour uint8, intrinsics, top-down pose for get_heightmap."""
    rs = np.random.RandomState(seed)
    fp, heights = _footprints(rs, num_objects, cluttered, size=400)
    depth = np.full((480, 640), 0.8, dtype=np.float64)
    top = np.zeros((400, 400), dtype=np.float64)
    for n in range(num_objects):
        top = np.maximum(top, fp[n] * heights[n])
    depth[0:400, 110:510] -= top
    depth += rs.uniform(-2e-4, 2e-4, size=depth.shape)
    color = rs.randint(0, 256, size=(480, 640, 3)).astype(np.uint8)
    cam_pose = np.asarray([[1, 0, 0, -0.5], [0, -1, 0, 0.0], [0, 0, -1, 0.8], [0, 0, 0, 1]], dtype=np.float64)
    # a small tilt so the rigid transform is not a pure axis flip
    t = 0.02
    rot = np.asarray([[1, 0, 0], [0, np.cos(t), -np.sin(t)], [0, np.sin(t), np.cos(t)]])
    cam_pose[:3, :3] = rot @ cam_pose[:3, :3]
    return {"color": color, "depth": depth, "intrinsics": CAM_INTRINSICS.copy(), "pose": cam_pose}


def make_boxes(seed, n=100):
    """n float32 detector boxes [n,2,2] in 224-space, roughly score-sorted clusters (for NMS)."""
    rs = np.random.RandomState(seed)
    centers = rs.uniform(20, 204, size=(max(1, n // 6), 2))
    boxes = np.zeros((n, 2, 2), dtype=np.float32)
    for i in range(n):
        c = centers[rs.randint(0, len(centers))] + rs.normal(0, 4, size=2)
        wh = rs.uniform(8, 120, size=2)
        boxes[i, 0] = c - wh / 2
        boxes[i, 1] = c + wh / 2
    scores = np.sort(rs.uniform(0.01, 1.0, size=n))[::-1].astype(np.float32)
    return boxes, scores
