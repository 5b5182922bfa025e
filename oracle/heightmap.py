"""CPU restatement of utils.get_pointcloud + utils.get_heightmap.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/code/utils.py:12-35 (get_pointcloud) and :38-68
(get_heightmap).  The reference delegates the arithmetic to third-party code
that is not in its tree: numpy `np.dot` (OpenBLAS dgemm) and opencv-python
(unpinned; 4.13.0 in this image) `cv2.getPerspectiveTransform`,
`cv2.warpPerspective` (INTER_LINEAR, BORDER_CONSTANT 0).  This file restates
their published algorithms in numpy; tests/test_oracle_golden.py pins it
bit-for-bit against the reference's own output (tests/golden/heightmap_*.npz).

Depth path recipe:
  z_w  = fma(R22, z, fma(R21, y, R20*x)) + t2           (float64, dgemm K=3 order)
  M    = getPerspectiveTransform(src quad -> dst square)  (8x8 LU solve, float64)
  Minv = cv2 invert(M) (3x3 adjugate / determinant, float64)
  per destination pixel (x,y): X0 = M00*x + M01*y + M02 (etc.), s = 32/W0,
  X = rint(X0*s), Y = rint(Y0*s)  -> 5 fractional bits; taps at (X>>5, Y>>5);
  weights = float32 products of (1-a) and a with a = (X&31)/32;
  out = sum tap*weight accumulated in float64 in tap order (y0x0,y0x1,y1x0,y1x1),
  taps outside the source image contribute the border value 0.
"""
import numpy as np

from ._fma import fma64 as _fma64

HEIGHTMAP_SIZE = 224
COLORMASK_SIZE = 448
SRC_QUAD = np.array([[110, 0], [110, 400], [510, 400], [510, 0]], np.float32)  # code/utils.py:49-55

INTER_BITS = 5
INTER_TAB_SIZE = 1 << INTER_BITS


def pointcloud_xyz(depth_img, K):
    """code/utils.py:18-21: x=(u-cx)*(d/fx), y=(v-cy)*(d/fy), z=d in float64."""
    im_h, im_w = depth_img.shape
    pix_x, pix_y = np.meshgrid(np.linspace(0, im_w - 1, im_w), np.linspace(0, im_h - 1, im_h))
    x = np.multiply(pix_x - K[0][2], depth_img / K[0][0])
    y = np.multiply(pix_y - K[1][2], depth_img / K[1][1])
    return x, y, depth_img.copy()


def world_z(depth_img, K, cam_pose):
    """Row 2 of R*p + t (code/utils.py:47,60-61) in the dgemm accumulation order of this image's numpy."""
    x, y, z = pointcloud_xyz(depth_img, K)
    R = cam_pose[0:3, 0:3]
    acc = R[2, 0] * x
    acc = _fma64(R[2, 1], y, acc)
    acc = _fma64(R[2, 2], z, acc)
    return acc + cam_pose[2, 3]


def get_perspective_transform(src, dst):
    """cv2.getPerspectiveTransform: solve the 8x8 system with LU (partial pivoting), float64."""
    a = np.zeros((8, 8), dtype=np.float64)
    b = np.zeros(8, dtype=np.float64)
    for i in range(4):
        a[i, 0] = a[i + 4, 3] = src[i][0]
        a[i, 1] = a[i + 4, 4] = src[i][1]
        a[i, 2] = a[i + 4, 5] = 1
        a[i, 6] = -float(src[i][0]) * float(dst[i][0])
        a[i, 7] = -float(src[i][1]) * float(dst[i][0])
        a[i + 4, 6] = -float(src[i][0]) * float(dst[i][1])
        a[i + 4, 7] = -float(src[i][1]) * float(dst[i][1])
        b[i] = dst[i][0]
        b[i + 4] = dst[i][1]
    x = _lu_solve(a, b)
    return np.append(x, 1.0).reshape(3, 3)


def _lu_solve(a, b):
    """cv::LU (hal::LU64f) forward elimination with partial pivoting + back substitution."""
    a = a.copy()
    b = b.copy()
    n = a.shape[0]
    for i in range(n):
        k = i
        for j in range(i + 1, n):
            if abs(a[j, i]) > abs(a[k, i]):
                k = j
        if k != i:
            a[[i, k], i:] = a[[k, i], i:]
            b[[i, k]] = b[[k, i]]
        d = -1 / a[i, i]
        for j in range(i + 1, n):
            alpha = a[j, i] * d
            for kk in range(i + 1, n):
                a[j, kk] += alpha * a[i, kk]
            b[j] += alpha * b[i]
    for i in range(n - 1, -1, -1):
        s = b[i]
        for k in range(i + 1, n):
            s -= a[i, k] * b[k]
        b[i] = s / a[i, i]
    return b


def invert3x3(m):
    """cv::invert for a 3x3 double matrix (DECOMP_LU fast path: adjugate * 1/det)."""
    d = (m[0, 0] * (m[1, 1] * m[2, 2] - m[1, 2] * m[2, 1])
         - m[0, 1] * (m[1, 0] * m[2, 2] - m[1, 2] * m[2, 0])
         + m[0, 2] * (m[1, 0] * m[2, 1] - m[1, 1] * m[2, 0]))
    d = 1.0 / d
    t = np.empty((3, 3), dtype=np.float64)
    t[0, 0] = (m[1, 1] * m[2, 2] - m[1, 2] * m[2, 1]) * d
    t[0, 1] = (m[0, 2] * m[2, 1] - m[0, 1] * m[2, 2]) * d
    t[0, 2] = (m[0, 1] * m[1, 2] - m[0, 2] * m[1, 1]) * d
    t[1, 0] = (m[1, 2] * m[2, 0] - m[1, 0] * m[2, 2]) * d
    t[1, 1] = (m[0, 0] * m[2, 2] - m[0, 2] * m[2, 0]) * d
    t[1, 2] = (m[0, 2] * m[1, 0] - m[0, 0] * m[1, 2]) * d
    t[2, 0] = (m[1, 0] * m[2, 1] - m[1, 1] * m[2, 0]) * d
    t[2, 1] = (m[0, 1] * m[2, 0] - m[0, 0] * m[2, 1]) * d
    t[2, 2] = (m[0, 0] * m[1, 1] - m[0, 1] * m[1, 0]) * d
    return t


def warp_perspective_f64(src, M, size):
    """cv2.warpPerspective(src float64, M, (size,size)) with INTER_LINEAR / BORDER_CONSTANT(0)."""
    Minv = invert3x3(np.asarray(M, dtype=np.float64))
    h, w = src.shape
    ys, xs = np.mgrid[0:size, 0:size].astype(np.float64)
    X0 = Minv[0, 0] * xs + Minv[0, 1] * ys + Minv[0, 2]
    Y0 = Minv[1, 0] * xs + Minv[1, 1] * ys + Minv[1, 2]
    W0 = Minv[2, 0] * xs + Minv[2, 1] * ys + Minv[2, 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        s = np.where(W0 != 0, INTER_TAB_SIZE / W0, 0.0)
    fX = np.clip(X0 * s, -2147483648.0, 2147483647.0)
    fY = np.clip(Y0 * s, -2147483648.0, 2147483647.0)
    X = np.rint(fX).astype(np.int64)
    Y = np.rint(fY).astype(np.int64)
    sx = X >> INTER_BITS
    sy = Y >> INTER_BITS
    ax = ((X & (INTER_TAB_SIZE - 1)).astype(np.float32)) * np.float32(1.0 / INTER_TAB_SIZE)
    ay = ((Y & (INTER_TAB_SIZE - 1)).astype(np.float32)) * np.float32(1.0 / INTER_TAB_SIZE)
    one = np.float32(1)
    w00 = ((one - ay) * (one - ax)).astype(np.float32)
    w01 = ((one - ay) * ax).astype(np.float32)
    w10 = (ay * (one - ax)).astype(np.float32)
    w11 = (ay * ax).astype(np.float32)

    def tap(yy, xx):
        ok = (yy >= 0) & (yy < h) & (xx >= 0) & (xx < w)
        return np.where(ok, src[np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)], 0.0)

    out = tap(sy, sx) * w00.astype(np.float64)
    out = out + tap(sy, sx + 1) * w01.astype(np.float64)
    out = out + tap(sy + 1, sx) * w10.astype(np.float64)
    out = out + tap(sy + 1, sx + 1) * w11.astype(np.float64)
    return out


def warp_perspective_u8(src, M, size):
    """cv2.warpPerspective(src uint8 [h,w,C], M, (size,size)) with INTER_LINEAR / BORDER_CONSTANT(0): same source
    coordinates as the float path, but cv2's 8-bit remap interpolates in 15-bit fixed point (imgwarp.cpp: BilinearTab_i,
    FixedPtCast<int, uchar, INTER_REMAP_COEF_BITS>): integer weights (32-ay)(32-ax)*32 ... that sum to 32768, result
    (sum + 2^14) >> 15.  The entry for ax = ay = 0 would be 32768 and saturates to (32767, 0, 0, 1) in cv2's table
    (initInterTab2D's sum correction lands on the last tap) - reproduced here although it cannot change an 8-bit result."""
    Minv = invert3x3(np.asarray(M, dtype=np.float64))
    h, w = src.shape[:2]
    ys, xs = np.mgrid[0:size, 0:size].astype(np.float64)
    X0 = Minv[0, 0] * xs + Minv[0, 1] * ys + Minv[0, 2]
    Y0 = Minv[1, 0] * xs + Minv[1, 1] * ys + Minv[1, 2]
    W0 = Minv[2, 0] * xs + Minv[2, 1] * ys + Minv[2, 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        s = np.where(W0 != 0, INTER_TAB_SIZE / W0, 0.0)
    X = np.rint(np.clip(X0 * s, -2147483648.0, 2147483647.0)).astype(np.int64)
    Y = np.rint(np.clip(Y0 * s, -2147483648.0, 2147483647.0)).astype(np.int64)
    sx, sy = X >> INTER_BITS, Y >> INTER_BITS
    ax, ay = X & (INTER_TAB_SIZE - 1), Y & (INTER_TAB_SIZE - 1)
    w00 = (32 - ay) * (32 - ax) * 32
    w01 = (32 - ay) * ax * 32
    w10 = ay * (32 - ax) * 32
    w11 = ay * ax * 32
    origin = (ax == 0) & (ay == 0)
    w00 = np.where(origin, 32767, w00)
    w11 = np.where(origin, 1, w11)

    def tap(yy, xx):
        ok = (yy >= 0) & (yy < h) & (xx >= 0) & (xx < w)
        v = src[np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)].astype(np.int64)
        return np.where(ok[..., None], v, 0)

    acc = (tap(sy, sx) * w00[..., None] + tap(sy, sx + 1) * w01[..., None] + tap(sy + 1, sx) * w10[..., None] +
           tap(sy + 1, sx + 1) * w11[..., None])
    return np.clip((acc + (1 << 14)) >> 15, 0, 255).astype(np.uint8)


def get_heightmap_color(color_img):
    """Colour outputs of utils.get_heightmap (code/utils.py:62,64): (color_heightmap 224^2, color_mask 448^2) uint8."""
    hs, cs = HEIGHTMAP_SIZE, COLORMASK_SIZE
    dst_h = np.array([[0, 0], [0, hs], [hs, hs], [hs, 0]], np.float32)
    dst_m = np.array([[0, 0], [0, cs], [cs, cs], [cs, 0]], np.float32)
    A_h = get_perspective_transform(SRC_QUAD, dst_h)
    A_m = get_perspective_transform(SRC_QUAD, dst_m)
    img = np.ascontiguousarray(color_img, dtype=np.uint8).reshape(480, 640, 3)
    return warp_perspective_u8(img, A_h, hs), warp_perspective_u8(img, A_m, cs)


def get_heightmap_depth(depth_img, cam_intrinsics, cam_pose):
    """Depth outputs of utils.get_heightmap: (depth_heightmap 224^2, depth_mask 448^2, A_htor)."""
    zw = world_z(np.asarray(depth_img, np.float64), cam_intrinsics, cam_pose).reshape(480, 640)
    hs, cs = HEIGHTMAP_SIZE, COLORMASK_SIZE
    dst_h = np.array([[0, 0], [0, hs], [hs, hs], [hs, 0]], np.float32)
    dst_m = np.array([[0, 0], [0, cs], [cs, cs], [cs, 0]], np.float32)
    A_h = get_perspective_transform(SRC_QUAD, dst_h)
    A_m = get_perspective_transform(SRC_QUAD, dst_m)
    d224 = warp_perspective_f64(zw, A_h, hs)
    d448 = warp_perspective_f64(zw, A_m, cs)
    A_htor = get_perspective_transform(dst_h, SRC_QUAD)
    return d224, d448, A_htor
