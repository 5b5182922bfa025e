// heightmap.cu - K11: utils.get_pointcloud + utils.get_heightmap, depth path
// (/root/reference/code/utils.py:12-35, :38-68), bit-exact with the reference's
// numpy + OpenBLAS + cv2.warpPerspective result (recipe: SURVEY.md section 8(a) row H1,
// restated and pinned in oracle/heightmap.py).
//
// The reference back-projects all 307 200 pixels, transforms them (np.dot), keeps world z
// as a 480x640 image and warps it twice (224^2 and 448^2, INTER_LINEAR, constant-0 border).
// Here one gather kernel computes, per destination pixel, the four bilinear taps' world z on
// the fly: no intermediate point cloud, depth read once through L2.  HBM-bound:
// 2.46 MB in, 2.01 MB out.  All arithmetic uses explicit round-to-nearest intrinsics so the
// compiler cannot contract a*b+c differently from the reference's operation order.
#include "smg_internal.cuh"

namespace smg {

struct HmParams {
    double minv224[9];
    double minv448[9];
    double fx, fy, cx, cy;
    double r20, r21, r22, t2;
};

__device__ __forceinline__ double world_z(const double* __restrict__ depth, int v, int u, const HmParams& p) {
    const double d = depth[v * 640 + u];
    const double x = __dmul_rn(__dsub_rn((double)u, p.cx), __ddiv_rn(d, p.fx));   // code/utils.py:19
    const double y = __dmul_rn(__dsub_rn((double)v, p.cy), __ddiv_rn(d, p.fy));   // code/utils.py:20
    double acc = __dmul_rn(p.r20, x);                                             // dgemm K=3, row 2
    acc = __fma_rn(p.r21, y, acc);
    acc = __fma_rn(p.r22, d, acc);
    return __dadd_rn(acc, p.t2);                                                  // code/utils.py:47
}

__global__ void heightmap_kernel(const double* __restrict__ depth, HmParams p, double* __restrict__ out224,
                                 double* __restrict__ out448) {
    const int n224 = 224 * 224, n448 = 448 * 448;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n224 + n448; i += gridDim.x * blockDim.x) {
        const bool small = i < n224;
        const int size = small ? 224 : 448;
        const int j = small ? i : i - n224;
        const double* M = small ? p.minv224 : p.minv448;
        const int dy = j / size, dx = j - dy * size;
        const double xs = (double)dx, ys = (double)dy;
        const double X0 = __dadd_rn(__dadd_rn(__dmul_rn(M[0], xs), __dmul_rn(M[1], ys)), M[2]);
        const double Y0 = __dadd_rn(__dadd_rn(__dmul_rn(M[3], xs), __dmul_rn(M[4], ys)), M[5]);
        const double W0 = __dadd_rn(__dadd_rn(__dmul_rn(M[6], xs), __dmul_rn(M[7], ys)), M[8]);
        const double sc = W0 != 0.0 ? __ddiv_rn(32.0, W0) : 0.0;
        double fX = __dmul_rn(X0, sc), fY = __dmul_rn(Y0, sc);
        fX = fmax(-2147483648.0, fmin(2147483647.0, fX));
        fY = fmax(-2147483648.0, fmin(2147483647.0, fY));
        const long long X = __double2ll_rn(fX), Y = __double2ll_rn(fY);
        const long long sx = X >> 5, sy = Y >> 5;
        const float ax = __fmul_rn((float)(X & 31), 1.0f / 32.0f);
        const float ay = __fmul_rn((float)(Y & 31), 1.0f / 32.0f);
        const float w00 = __fmul_rn(__fsub_rn(1.0f, ay), __fsub_rn(1.0f, ax));
        const float w01 = __fmul_rn(__fsub_rn(1.0f, ay), ax);
        const float w10 = __fmul_rn(ay, __fsub_rn(1.0f, ax));
        const float w11 = __fmul_rn(ay, ax);
        double t00 = 0.0, t01 = 0.0, t10 = 0.0, t11 = 0.0;
        const bool y0 = sy >= 0 && sy < 480, y1 = sy + 1 >= 0 && sy + 1 < 480;
        const bool x0 = sx >= 0 && sx < 640, x1 = sx + 1 >= 0 && sx + 1 < 640;
        if (y0 && x0) t00 = world_z(depth, (int)sy, (int)sx, p);
        if (y0 && x1) t01 = world_z(depth, (int)sy, (int)sx + 1, p);
        if (y1 && x0) t10 = world_z(depth, (int)sy + 1, (int)sx, p);
        if (y1 && x1) t11 = world_z(depth, (int)sy + 1, (int)sx + 1, p);
        double o = __dmul_rn(t00, (double)w00);
        o = __dadd_rn(o, __dmul_rn(t01, (double)w01));
        o = __dadd_rn(o, __dmul_rn(t10, (double)w10));
        o = __dadd_rn(o, __dmul_rn(t11, (double)w11));
        (small ? out224 : out448)[j] = o;
    }
}

// Colour path (code/utils.py:62,64): cv2.warpPerspective of the uint8 camera image with the same two transforms.  Same
// source coordinates as above; cv2's 8-bit remap interpolates in 15-bit fixed point (imgwarp.cpp BilinearTab_i +
// FixedPtCast<int, uchar, 15>): integer weights (32-ay)(32-ax)*32 ... summing to 32768, result (sum + 2^14) >> 15; the
// table entry for ax = ay = 0 saturates to (32767, 0, 0, 1).  Pinned against cv2 4.13 in tests/test_oracle_golden.py.
__global__ void heightmap_color_kernel(const uint8_t* __restrict__ color, HmParams p, uint8_t* __restrict__ out224,
                                       uint8_t* __restrict__ out448) {
    const int n224 = 224 * 224, n448 = 448 * 448;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n224 + n448; i += gridDim.x * blockDim.x) {
        const bool small = i < n224;
        const int size = small ? 224 : 448;
        const int j = small ? i : i - n224;
        const double* M = small ? p.minv224 : p.minv448;
        const int dy = j / size, dx = j - dy * size;
        const double xs = (double)dx, ys = (double)dy;
        const double X0 = __dadd_rn(__dadd_rn(__dmul_rn(M[0], xs), __dmul_rn(M[1], ys)), M[2]);
        const double Y0 = __dadd_rn(__dadd_rn(__dmul_rn(M[3], xs), __dmul_rn(M[4], ys)), M[5]);
        const double W0 = __dadd_rn(__dadd_rn(__dmul_rn(M[6], xs), __dmul_rn(M[7], ys)), M[8]);
        const double sc = W0 != 0.0 ? __ddiv_rn(32.0, W0) : 0.0;
        double fX = __dmul_rn(X0, sc), fY = __dmul_rn(Y0, sc);
        fX = fmax(-2147483648.0, fmin(2147483647.0, fX));
        fY = fmax(-2147483648.0, fmin(2147483647.0, fY));
        const long long X = __double2ll_rn(fX), Y = __double2ll_rn(fY);
        const long long sx = X >> 5, sy = Y >> 5;
        const int ax = (int)(X & 31), ay = (int)(Y & 31);
        int w00 = (32 - ay) * (32 - ax) * 32, w01 = (32 - ay) * ax * 32, w10 = ay * (32 - ax) * 32, w11 = ay * ax * 32;
        if (ax == 0 && ay == 0) { w00 = 32767; w11 = 1; }
        const bool y0 = sy >= 0 && sy < 480, y1 = sy + 1 >= 0 && sy + 1 < 480;
        const bool x0 = sx >= 0 && sx < 640, x1 = sx + 1 >= 0 && sx + 1 < 640;
        uint8_t* o = (small ? out224 : out448) + (size_t)j * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            int acc = 1 << 14;
            if (y0 && x0) acc += w00 * color[((int)sy * 640 + (int)sx) * 3 + c];
            if (y0 && x1) acc += w01 * color[((int)sy * 640 + (int)sx + 1) * 3 + c];
            if (y1 && x0) acc += w10 * color[(((int)sy + 1) * 640 + (int)sx) * 3 + c];
            if (y1 && x1) acc += w11 * color[(((int)sy + 1) * 640 + (int)sx + 1) * 3 + c];
            const int v = acc >> 15;
            o[c] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
        }
    }
}

// ---- host: cv2.getPerspectiveTransform (8x8 LU with partial pivoting) and cv2.invert (3x3) ----
// compiled with -ffp-contract=off so the double operations round exactly like numpy / OpenCV's C++.
static void lu_solve8(double a[8][8], double b[8]) {
    const int n = 8;
    for (int i = 0; i < n; ++i) {
        int k = i;
        for (int j = i + 1; j < n; ++j)
            if (fabs(a[j][i]) > fabs(a[k][i])) k = j;
        if (k != i) {
            for (int j = i; j < n; ++j) { double t = a[i][j]; a[i][j] = a[k][j]; a[k][j] = t; }
            double t = b[i]; b[i] = b[k]; b[k] = t;
        }
        const double d = -1 / a[i][i];
        for (int j = i + 1; j < n; ++j) {
            const double alpha = a[j][i] * d;
            for (int kk = i + 1; kk < n; ++kk) {
                const double prod = alpha * a[i][kk];
                a[j][kk] += prod;
            }
            const double pb = alpha * b[i];
            b[j] += pb;
        }
    }
    for (int i = n - 1; i >= 0; --i) {
        double s = b[i];
        for (int k = i + 1; k < n; ++k) {
            const double prod = a[i][k] * b[k];
            s -= prod;
        }
        b[i] = s / a[i][i];
    }
}

void perspective_transform(const float src[4][2], const float dst[4][2], double M[9]) {
    double a[8][8] = {{0}};
    double b[8];
    for (int i = 0; i < 4; ++i) {
        a[i][0] = a[i + 4][3] = src[i][0];
        a[i][1] = a[i + 4][4] = src[i][1];
        a[i][2] = a[i + 4][5] = 1;
        a[i][6] = -(double)src[i][0] * (double)dst[i][0];
        a[i][7] = -(double)src[i][1] * (double)dst[i][0];
        a[i + 4][6] = -(double)src[i][0] * (double)dst[i][1];
        a[i + 4][7] = -(double)src[i][1] * (double)dst[i][1];
        b[i] = dst[i][0];
        b[i + 4] = dst[i][1];
    }
    lu_solve8(a, b);
    for (int i = 0; i < 8; ++i) M[i] = b[i];
    M[8] = 1.0;
}

void invert3x3(const double* m, double* t) {
#define A(i, j) m[(i)*3 + (j)]
    double d = A(0, 0) * (A(1, 1) * A(2, 2) - A(1, 2) * A(2, 1)) - A(0, 1) * (A(1, 0) * A(2, 2) - A(1, 2) * A(2, 0)) +
               A(0, 2) * (A(1, 0) * A(2, 1) - A(1, 1) * A(2, 0));
    d = 1.0 / d;
    t[0] = (A(1, 1) * A(2, 2) - A(1, 2) * A(2, 1)) * d;
    t[1] = (A(0, 2) * A(2, 1) - A(0, 1) * A(2, 2)) * d;
    t[2] = (A(0, 1) * A(1, 2) - A(0, 2) * A(1, 1)) * d;
    t[3] = (A(1, 2) * A(2, 0) - A(1, 0) * A(2, 2)) * d;
    t[4] = (A(0, 0) * A(2, 2) - A(0, 2) * A(2, 0)) * d;
    t[5] = (A(0, 2) * A(1, 0) - A(0, 0) * A(1, 2)) * d;
    t[6] = (A(1, 0) * A(2, 1) - A(1, 1) * A(2, 0)) * d;
    t[7] = (A(0, 1) * A(2, 0) - A(0, 0) * A(2, 1)) * d;
    t[8] = (A(0, 0) * A(1, 1) - A(0, 1) * A(1, 0)) * d;
#undef A
}

int launch_heightmap(smg_handle* h, const double* depth, const double* K, const double* pose, double* out224,
                     double* out448, double* host_A_htor, cudaStream_t st) {
    const float src[4][2] = {{110, 0}, {110, 400}, {510, 400}, {510, 0}};  // code/utils.py:49-50,55
    const float d224[4][2] = {{0, 0}, {0, 224}, {224, 224}, {224, 0}};
    const float d448[4][2] = {{0, 0}, {0, 448}, {448, 448}, {448, 0}};
    double M224[9], M448[9];
    perspective_transform(src, d224, M224);
    perspective_transform(src, d448, M448);
    HmParams p;
    invert3x3(M224, p.minv224);
    invert3x3(M448, p.minv448);
    p.fx = K[0]; p.fy = K[4]; p.cx = K[2]; p.cy = K[5];
    p.r20 = pose[8]; p.r21 = pose[9]; p.r22 = pose[10]; p.t2 = pose[11];
    if (host_A_htor) perspective_transform(d224, src, host_A_htor);  // code/utils.py:66
    const int total = 224 * 224 + 448 * 448;
    heightmap_kernel<<<(total + 255) / 256, 256, 0, st>>>(depth, p, out224, out448);
    h->launches++;
    SMG_CUDA(cudaGetLastError());
    return SMG_OK;
}

int launch_heightmap_color(smg_handle* h, const uint8_t* color, uint8_t* out224, uint8_t* out448, cudaStream_t st) {
    const float src[4][2] = {{110, 0}, {110, 400}, {510, 400}, {510, 0}};
    const float d224[4][2] = {{0, 0}, {0, 224}, {224, 224}, {224, 0}};
    const float d448[4][2] = {{0, 0}, {0, 448}, {448, 448}, {448, 0}};
    double M224[9], M448[9];
    perspective_transform(src, d224, M224);
    perspective_transform(src, d448, M448);
    HmParams p = {};
    invert3x3(M224, p.minv224);
    invert3x3(M448, p.minv448);
    const int total = 224 * 224 + 448 * 448;
    heightmap_color_kernel<<<(total + 255) / 256, 256, 0, st>>>(color, p, out224, out448);
    h->launches++;
    SMG_CUDA(cudaGetLastError());
    return SMG_OK;
}

}  // namespace smg
