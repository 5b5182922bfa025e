// geometry.cu - the pre-enveloping / orientation-optimisation geometry that follows the action argmax
// (/root/reference/code/utils.py:70-81 global_position, :316-366 get_best_grasp_angle, :370-612 get_best_suction_angle),
// kept on the device so that a decision never leaves the GPU between the Q tables and the motion parameters.
// The work is O(objects x 360) scalar double arithmetic: ONE thread of one CTA runs it (the chain is sequential by
// construction: run-length segments of a 360-bin table, relaxed value by value until an opening of >= 45 degrees exists).
//
// What the reference computes, restated:
//   global_position      pixel (row r, col c) of the 224 heightmap -> camera pixel through the homography A_htor (truncated
//                        towards zero), depth lookup, pinhole back-projection, rigid transform into the robot frame.
//   grasp angle / width  centre = mean of the min-area box corners; with PE: the box's side lengths in the robot frame give
//                        the opening (short side x min(1.2, long/short)) and the jaw angle (acos of the long side's x-slope).
//   suction angle        with OO: every other object blocks the angular interval its box subtends around the target's
//                        centre, weighted by exp(-height difference / distance); the widest fully free interval (>= 45
//                        degrees) is chosen, objects being dropped from the weakest blocker upwards until one exists.
#include "smg_internal.cuh"

namespace smg {

namespace {

constexpr int kMaxObj = 32;
constexpr double kPi = 3.141592653589793;

struct GeoIn {
    double A[9];        // A_htor
    double K[9];        // camera intrinsics
    double P[16];       // camera pose
    double box[kMaxObj][4][2];
    double cter[kMaxObj][2];
    double pix[3];
    int n, best, flag, mode, img_h, img_w;
};

struct Pos {
    double x, y, z;
    int ok;
};

// utils.py:70-81.  pix = (_, row, col); int() truncates towards zero
__device__ Pos global_position(const GeoIn& g, const double* __restrict__ depth, double row, double col) {
    const double den = col * g.A[6] + row * g.A[7] + g.A[8];
    const int px = (int)((col * g.A[0] + row * g.A[1] + g.A[2]) / den);
    const int py = (int)((col * g.A[3] + row * g.A[4] + g.A[5]) / den);
    Pos o;
    o.ok = px >= 0 && px < g.img_w && py >= 0 && py < g.img_h;
    const double z = o.ok ? depth[(size_t)py * g.img_w + px] : 0.0;
    const double cx = ((double)px - g.K[2]) * (z / g.K[0]);
    const double cy = ((double)py - g.K[5]) * (z / g.K[4]);
    o.x = g.P[0] * cx + g.P[1] * cy + g.P[2] * z + g.P[3];
    o.y = g.P[4] * cx + g.P[5] * cy + g.P[6] * z + g.P[7];
    o.z = g.P[8] * cx + g.P[9] * cy + g.P[10] * z + g.P[11];
    return o;
}

__device__ Pos box_centre(const GeoIn& g, const double* depth, int id) {
    // mean of the four corners, truncated like np.array(...).astype(int) (utils.py:318-322)
    const double r = (g.box[id][0][1] + g.box[id][1][1] + g.box[id][2][1] + g.box[id][3][1]) / 4;
    const double c = (g.box[id][0][0] + g.box[id][1][0] + g.box[id][2][0] + g.box[id][3][0]) / 4;
    return global_position(g, depth, (double)(long long)r, (double)(long long)c);
}

// run-length segments of the 360-bin table exactly as utils.py:467-476 builds them: a segment is closed when the value
// changes; the last one is appended only if it did not start at bin 359
__device__ int segments(const double* val, double* seg_val, int* seg_lo, int* seg_hi) {
    int n = 0, start = 0;
    double cur = val[0];
    for (int i = 0; i < 360; ++i) {
        if (val[i] != cur) {
            seg_val[n] = cur; seg_lo[n] = start; seg_hi[n] = i - 1; ++n;
            cur = val[i];
            start = i;
        }
        if (i == 359 && start != i) {
            seg_val[n] = cur; seg_lo[n] = start; seg_hi[n] = i; ++n;
        }
    }
    return n;
}

__device__ void vote(const GeoIn& g, const double (*ov)[3], double* val) {
    for (int i = 0; i < 360; ++i) val[i] = 1.0;
    for (int i = 0; i < g.n; ++i) {
        if (i == g.best || ov[i][2] == 1.0) continue;
        const int a0 = (int)(180 * ov[i][0] / kPi), a1 = (int)(180 * ov[i][1] / kPi);
        if (fabs(ov[i][0] - ov[i][1]) <= kPi) {
            for (int a = a0; a < a1; ++a) val[a] *= ov[i][2];
        } else {
            for (int a = 0; a < a0; ++a) val[a] *= ov[i][2];
            for (int a = a1; a < 360; ++a) val[a] *= ov[i][2];
        }
    }
}

__global__ void geometry_kernel(GeoIn g, const double* __restrict__ depth, double* __restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    // out: [0..2] position, [3] angle, [4] open distance, [5] status (0 ok, 1 pixel outside the camera image)
    int bad = 0;
    if (g.mode == 0) {
        const Pos p = global_position(g, depth, g.pix[1], g.pix[2]);
        out[0] = p.x; out[1] = p.y; out[2] = p.z; out[3] = 0; out[4] = 0; out[5] = p.ok ? 0 : 1;
        return;
    }
    const int b = g.best;
    const Pos c = box_centre(g, depth, b);
    bad |= !c.ok;
    out[0] = c.x; out[1] = c.y; out[2] = c.z;
    if (g.mode == 1) {
        // ---------------------------------------------------------------- get_best_grasp_angle (utils.py:316-366)
        double angle = 0.0, open_d = 2.0;   // 2 m: "larger than the threshold"
        if (g.flag) {
            Pos q[4];
            for (int i = 0; i < 4; ++i) {
                q[i] = global_position(g, depth, (double)(long long)g.box[b][i][1], (double)(long long)g.box[b][i][0]);
                bad |= !q[i].ok;
            }
            const double d01 = sqrt((q[0].x - q[1].x) * (q[0].x - q[1].x) + (q[0].y - q[1].y) * (q[0].y - q[1].y));
            const double d12 = sqrt((q[2].x - q[1].x) * (q[2].x - q[1].x) + (q[2].y - q[1].y) * (q[2].y - q[1].y));
            if (d01 > d12) {
                open_d = d12 * fmin(1.2, d01 / d12);
                if (q[0].y == q[1].y) angle = 0;
                else if (q[0].y > q[1].y) angle = acos((q[0].x - q[1].x) / d01);
                else angle = acos((q[1].x - q[0].x) / d01);
            } else {
                open_d = d01 * fmin(1.2, d12 / d01);
                if (q[2].y == q[1].y) angle = 0;
                else if (q[2].y > q[1].y) angle = acos((q[2].x - q[1].x) / d12);
                else angle = acos((q[1].x - q[2].x) / d12);
            }
        }
        out[3] = angle; out[4] = open_d; out[5] = bad;
        return;
    }
    // -------------------------------------------------------------------- get_best_suction_angle (utils.py:370-540)
    double selected = 0.0;
    if (g.flag) {
        __shared__ double val[360], seg_val[361];
        __shared__ int seg_lo[361], seg_hi[361];
        double ov[kMaxObj][3], height[kMaxObj], dist[kMaxObj];
        Pos ctr[kMaxObj];
        for (int i = 0; i < g.n; ++i) {
            ctr[i] = global_position(g, depth, g.cter[i][1], g.cter[i][0]);
            bad |= !ctr[i].ok;
            double hmax = ctr[i].z;
            for (int j = 0; j < 4; ++j) {
                const Pos q = global_position(g, depth, (double)(long long)g.box[i][j][1], (double)(long long)g.box[i][j][0]);
                bad |= !q.ok;
                hmax = fmax(hmax, q.z);
            }
            height[i] = hmax;
            ov[i][0] = ov[i][1] = ov[i][2] = 1.0;
        }
        for (int i = 0; i < g.n; ++i)
            dist[i] = sqrt((ctr[i].x - ctr[b].x) * (ctr[i].x - ctr[b].x) + (ctr[i].y - ctr[b].y) * (ctr[i].y - ctr[b].y));
        const double cx = g.cter[b][0], cy = g.cter[b][1];
        for (int o = 0; o < g.n; ++o) {
            if (o == b) continue;
            double ap[4];
            for (int k = 0; k < 4; ++k) {
                const double x = g.box[o][k][0], y = g.box[o][k][1];
                double a = 0.0;
                if (x == cx) a = y > cy ? kPi : 0.0;
                if (y == cy) a = x < cx ? kPi / 2 : 3 * kPi / 2;
                if (x < cx) {
                    if (y < cy) a = atan((cx - x) / (cy - y));
                    else if (y > cy) a = kPi / 2 + atan((y - cy) / (cx - x));
                }
                if (x > cx) {
                    if (y < cy) a = 3 * kPi / 2 + atan((cy - y) / (x - cx));
                    else if (y > cy) a = kPi + atan((x - cx) / (y - cy));
                }
                ap[k] = a;
            }
            double amax = 0.0;
            for (int i = 0; i < 3; ++i)
                for (int j = i + 1; j < 4; ++j) {
                    const double d = fmin(fabs(ap[i] - ap[j]), 2 * kPi - fabs(ap[i] - ap[j]));
                    if (d > amax) {
                        amax = d;
                        ov[o][0] = fmin(ap[i], ap[j]);
                        ov[o][1] = fmax(ap[i], ap[j]);
                    }
                }
        }
        for (int i = 0; i < g.n; ++i) ov[i][2] = exp(-fmax(0.0, height[i] - height[b]) / fmax(0.001, dist[i]));
        // distinct blocking values in descending order, 1.0 first (object_val_set / pre_sorted, utils.py:478-483)
        double vals[kMaxObj + 1];
        int nv = 0;
        vals[nv++] = 1.0;
        for (int i = 0; i < g.n; ++i) {
            bool seen = false;
            for (int j = 0; j < nv; ++j) seen |= vals[j] == ov[i][2];
            if (!seen) vals[nv++] = ov[i][2];
        }
        for (int i = 1; i < nv; ++i)
            for (int j = i; j > 0 && vals[j] > vals[j - 1]; --j) { const double t = vals[j]; vals[j] = vals[j - 1]; vals[j - 1] = t; }
        vote(g, ov, val);
        int nseg = segments(val, seg_val, seg_lo, seg_hi);
        bool found = false;
        for (int round = 0; round < nv && !found; ++round) {
            double vmin = seg_val[0];
            for (int i = 1; i < nseg; ++i) vmin = fmin(vmin, seg_val[i]);
            if (vmin >= 0.95) { selected = 0.0; found = true; break; }
            // the free interval that wraps around 0 degrees (utils.py:490-499)
            if (val[1] == val[359] && seg_val[0] >= 1.0) {
                const int left = seg_hi[0], right = seg_hi[nseg - 1] - seg_lo[nseg - 1];
                if (left + right >= 45) {
                    selected = left > right ? left - (left + right) / 2 : seg_lo[nseg - 1] + (left + right) / 2;
                    found = true;
                    break;
                }
            }
            // otherwise the longest fully free interval of at least 45 degrees (ties: the last one in argsort order)
            int best_len = -1, best_mid = 0;
            for (int i = 0; i < nseg; ++i) {
                const int len = seg_hi[i] - seg_lo[i];
                if (seg_val[i] >= 1.0 && len >= 45 && len >= best_len) { best_len = len; best_mid = (seg_lo[i] + seg_hi[i]) / 2; }
            }
            if (best_len >= 0) { selected = best_mid; found = true; break; }
            // nothing: stop counting the objects whose weight is the next one down, and vote again (utils.py:524-530)
            if (round + 1 >= nv) break;
            for (int i = 0; i < g.n; ++i)
                if (fabs(ov[i][2] - vals[round + 1]) < 0.001) ov[i][2] = 1.0;
            vote(g, ov, val);
            nseg = segments(val, seg_val, seg_lo, seg_hi);
        }
        if (!found) bad |= 2;   // the reference would raise IndexError here
    }
    out[3] = selected * (kPi / 180.0);   // np.deg2rad
    out[4] = 0; out[5] = bad;
}

}  // namespace

int launch_geometry(smg_handle* h, int mode, const double* dev_depth, int img_h, int img_w, const double* A, const double* K,
                    const double* P, const double* boxes, const double* centers, int n, int best, int flag, const double* pix,
                    double* dev_out, cudaStream_t st) {
    SMG_CHECK(n >= 0 && n <= kMaxObj, SMG_ERR_INVALID, "geometry: %d objects (max %d)", n, kMaxObj);
    SMG_CHECK(mode == 0 || (best >= 0 && best < n), SMG_ERR_INVALID, "geometry: best id %d of %d", best, n);
    GeoIn g{};
    for (int i = 0; i < 9; ++i) { g.A[i] = A[i]; g.K[i] = K[i]; }
    for (int i = 0; i < 16; ++i) g.P[i] = P[i];
    for (int i = 0; i < n; ++i) {
        for (int j = 0; j < 4; ++j) { g.box[i][j][0] = boxes[(i * 4 + j) * 2]; g.box[i][j][1] = boxes[(i * 4 + j) * 2 + 1]; }
        if (centers) { g.cter[i][0] = centers[2 * i]; g.cter[i][1] = centers[2 * i + 1]; }
    }
    if (pix) { g.pix[0] = pix[0]; g.pix[1] = pix[1]; g.pix[2] = pix[2]; }
    g.n = n; g.best = best; g.flag = flag; g.mode = mode; g.img_h = img_h; g.img_w = img_w;
    geometry_kernel<<<1, 32, 0, st>>>(g, dev_depth, dev_out);
    h->launches++;
    SMG_CUDA(cudaGetLastError());
    return SMG_OK;
}

}  // namespace smg
