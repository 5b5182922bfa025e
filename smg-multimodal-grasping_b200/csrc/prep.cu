// prep.cu - K1: the input stage of the Q pass.  HBM-bound, one pass, fully coalesced writes.
//
//   smg_prep    code/trainer.py:165-191   nearest zoom x2 + zero pad + (x-mean)/std + 3 channels
//   smg_rotate  code/models.py:371-382    F.affine_grid(align_corners=True) + F.grid_sample(nearest)
//
// The rotation index arithmetic replicates torch's float32 evaluation order exactly
// (probe-verified against torch 2.11 CPU, see oracle/qnet.py::rotate_index_map):
//   base   = linspace(-1,1,H):  fma(step,i,-1) for i < H/2, fma(-step,H-1-i,1) otherwise
//   grid   = fma(1,t2, fma(by,t1, bx*t0))      (bmm K=3 accumulation order)
//   index  = nearbyint((grid+1) * ((H-1)/2))   (ties to even), zero outside [0,H)
#include "smg_internal.cuh"

#include <math.h>

namespace smg {

__global__ void prep_kernel(const double* __restrict__ hm, int n, int hs, double mean, double stddev,
                            float* __restrict__ out, int H, int channels) {
    const int pad = (H - 2 * hs) / 2;
    const size_t plane = (size_t)H * H;
    const size_t total = (size_t)n * plane;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int s = (int)(i / plane);
        const int rem = (int)(i - (size_t)s * plane);
        const int y = rem / H, x = rem - y * H;
        const int yy = y - pad, xx = x - pad;
        double v = 0.0;
        if (yy >= 0 && yy < 2 * hs && xx >= 0 && xx < 2 * hs) v = hm[((size_t)s * hs + (yy >> 1)) * hs + (xx >> 1)];
        const float f = (float)((v - mean) / stddev);
        float* o = out + (size_t)s * channels * plane + rem;
        for (int c = 0; c < channels; ++c) o[c * plane] = f;
    }
}

__device__ __forceinline__ float base_coord(int i, int H, float step) {
    return (i < H / 2) ? __fmaf_rn(step, (float)i, -1.0f) : __fmaf_rn(-step, (float)(H - 1 - i), 1.0f);
}

__device__ __forceinline__ int rotate_src_index(int x, int y, int H, const float* t) {
    const float step = __fdiv_rn(2.0f, (float)(H - 1));
    const float half = __fdiv_rn((float)(H - 1), 2.0f);
    const float bx = base_coord(x, H, step);
    const float by = base_coord(y, H, step);
    const float gx = __fmaf_rn(1.0f, t[2], __fmaf_rn(by, t[1], __fmul_rn(bx, t[0])));
    const float gy = __fmaf_rn(1.0f, t[5], __fmaf_rn(by, t[4], __fmul_rn(bx, t[3])));
    const float ix = rintf(__fmul_rn(__fadd_rn(gx, 1.0f), half));
    const float iy = rintf(__fmul_rn(__fadd_rn(gy, 1.0f), half));
    if (ix >= 0.0f && ix < (float)H && iy >= 0.0f && iy < (float)H) return (int)iy * H + (int)ix;
    return -1;
}

// the same index arithmetic with the per-launch constants (step = 2/(H-1), half = (H-1)/2, both rounded to float exactly
// like the divisions above) and the row coordinate hoisted: what the fused input kernel runs per pixel
__device__ __forceinline__ bool rotate_src_xy(float bx, float by, int H, float half, const float* t, int& sx, int& sy) {
    const float gx = __fmaf_rn(1.0f, t[2], __fmaf_rn(by, t[1], __fmul_rn(bx, t[0])));
    const float gy = __fmaf_rn(1.0f, t[5], __fmaf_rn(by, t[4], __fmul_rn(bx, t[3])));
    const float ix = rintf(__fmul_rn(__fadd_rn(gx, 1.0f), half));
    const float iy = rintf(__fmul_rn(__fadd_rn(gy, 1.0f), half));
    sx = (int)ix;
    sy = (int)iy;
    return ix >= 0.0f && ix < (float)H && iy >= 0.0f && iy < (float)H;
}

struct RotTheta {
    float t[32][6];
};

__global__ void rotate_kernel(const float* __restrict__ in, int n_rot, RotTheta th, float* __restrict__ out, int H,
                              int channels) {
    const size_t plane = (size_t)H * H;
    const size_t total = (size_t)n_rot * plane;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / plane);
        const int rem = (int)(i - (size_t)r * plane);
        const int y = rem / H, x = rem - y * H;
        const int src = rotate_src_index(x, y, H, th.t[r]);
        float* o = out + (size_t)r * channels * plane + rem;
        for (int c = 0; c < channels; ++c) o[c * plane] = src >= 0 ? __ldg(in + c * plane + src) : 0.f;
    }
}

// K1 for the heightmap entry points: two launches for a whole pass, whatever the number of units.
//   normalize_hm_kernel  (x - mean)/std of every 224x224 float64 heightmap -> float32, computed ONCE per heightmap pixel
//                        (code/trainer.py:176-188: float64 arithmetic, then the cast) - 50 k pixels per map, L2-resident;
//   prep_rotate_kernel   sample z < n_scene_samples is scene (z / n_rot) at rotation (z % n_rot), the remaining samples are
//                        the masked heightmaps unrotated: the value of network-input pixel (x, y) is the normalised
//                        heightmap value at the source pixel of the zoom x2 + pad (code/trainer.py:165-173) composed with
//                        the nearest rotation (code/models.py:371-382); the padded / rotated 640x640 image in between is
//                        never materialised.  One plane per sample (the three channels the reference feeds are
//                        identical), four pixels per thread, one 16-byte store each: HBM-write bound.
__global__ void normalize_hm_kernel(const double* __restrict__ scene_hm, int n_scene, const double* __restrict__ mask_hm,
                                    int n_mask, int px, double mean, double stddev, float* __restrict__ out) {
    const size_t total = (size_t)(n_scene + n_mask) * px;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t m = i / px;
        const double v = m < (size_t)n_scene ? scene_hm[i] : mask_hm[i - (size_t)n_scene * px];
        out[i] = (float)((v - mean) / stddev);
    }
}

__global__ void __launch_bounds__(256)
prep_rotate_kernel(const float* __restrict__ norm_hm, int groups, int n_rot, int n_scene_samples, int n_samples,
                   const __grid_constant__ RotTheta th, int hs, float pad_val, float* __restrict__ out, int H, float step, float half) {
    // blockIdx.y = sample: no 64-bit index arithmetic per pixel; the matrix table is a __grid_constant__ parameter, so the
    // sample's row is read in place (a plain by-value array indexed dynamically is first copied to every thread's local memory)
    const int z = blockIdx.y;
    const bool is_scene = z < n_scene_samples;
    const float* t = th.t[is_scene ? z % n_rot : 0];
    const int pad = (H - 2 * hs) / 2;
    const int quads = H * H / 4;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= quads) return;
    const int rem = i * 4;
    const int y = rem / H, x0 = rem - y * H;
    const float* hm = norm_hm + (size_t)(is_scene ? z / n_rot : groups + z - n_scene_samples) * hs * hs;
    const float by = base_coord(y, H, step);
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int sy = y, sx = x0 + j;
        bool inside = true;
        if (is_scene) inside = rotate_src_xy(base_coord(x0 + j, H, step), by, H, half, t, sx, sy);
        const int yy = sy - pad, xx = sx - pad;
        const bool in_hm = inside && yy >= 0 && yy < 2 * hs && xx >= 0 && xx < 2 * hs;
        const float f = in_hm ? __ldg(hm + (size_t)(yy >> 1) * hs + (xx >> 1)) : pad_val;
        v[j] = inside ? f : 0.f;   // grid_sample pads with zeros OUTSIDE the (already normalised) image
    }
    *reinterpret_cast<float4*>(out + (size_t)z * H * H + rem) = make_float4(v[0], v[1], v[2], v[3]);
}

__global__ void rotate_index_kernel(RotTheta th, int32_t* __restrict__ out, int H) {
    const int total = H * H;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int y = i / H, x = i - y * H;
        out[i] = rotate_src_index(x, y, H, th.t[0]);
    }
}

// code/models.py:372-376: float64 angle -> 2x3 affine matrix -> float32
void rotation_theta(int rot_idx, int num_rot, float* t6) {
    const double deg = (double)rot_idx * (360.0 / (double)num_rot);
    const double th = deg * (M_PI / 180.0);  // np.radians
    const double c = cos(-th), s = sin(-th);
    t6[0] = (float)c;
    t6[1] = (float)s;
    t6[2] = 0.0f;
    t6[3] = (float)(-s);
    t6[4] = (float)c;
    t6[5] = 0.0f;
}

int launch_prep(smg_handle* h, const double* hm, int n, int hm_size, double mean, double stddev, float* out,
                int channels, cudaStream_t st) {
    SMG_CHECK(2 * hm_size <= h->H, SMG_ERR_INVALID, "smg_prep: 2*hm_size %d exceeds H %d", 2 * hm_size, h->H);
    const size_t total = (size_t)n * h->H * h->H;
    const int threads = 256;
    const int blocks = (int)((total + threads - 1) / threads);
    prep_kernel<<<blocks < h->num_sms * 16 ? blocks : h->num_sms * 16, threads, 0, st>>>(hm, n, hm_size, mean, stddev,
                                                                                         out, h->H, channels);
    h->launches++;
    SMG_CUDA(cudaGetLastError());
    return SMG_OK;
}

int launch_rotate(smg_handle* h, const float* in, const int* host_rot, int n_rot, int num_rot, float* out,
                  int channels, cudaStream_t st) {
    for (int base = 0; base < n_rot; base += 32) {
        const int cnt = n_rot - base < 32 ? n_rot - base : 32;
        RotTheta th;
        for (int i = 0; i < cnt; ++i) rotation_theta(host_rot[base + i], num_rot, th.t[i]);
        const size_t total = (size_t)cnt * h->H * h->H;
        const int threads = 256;
        const int blocks = (int)((total + threads - 1) / threads);
        rotate_kernel<<<blocks < h->num_sms * 16 ? blocks : h->num_sms * 16, threads, 0, st>>>(
            in, cnt, th, out + (size_t)base * channels * h->H * h->H, h->H, channels);
        h->launches++;
        SMG_CUDA(cudaGetLastError());
    }
    return SMG_OK;
}

int launch_prep_rotate(smg_handle* h, const double* scene_hm, int groups, const int* host_rot, int n_rot, int num_rot,
                       const double* mask_hm, int n_mask_samples, int hm_size, double mean, double stddev, float* out, cudaStream_t st) {
    SMG_CHECK(2 * hm_size <= h->H && h->H % 4 == 0 && n_rot >= 1 && n_rot <= 32, SMG_ERR_INVALID,
              "prep_rotate: hm_size %d / H %d / %d rotations", hm_size, h->H, n_rot);
    SMG_CHECK((size_t)(groups + n_mask_samples) * hm_size * hm_size <= (size_t)(h->max_samples > 3 ? h->max_samples : 3) * h->H * h->H,
              SMG_ERR_STATE, "prep_rotate: %d heightmaps do not fit the staging buffer", groups + n_mask_samples);
    RotTheta th;
    for (int i = 0; i < n_rot; ++i) rotation_theta(host_rot[i], num_rot, th.t[i]);
    const int px = hm_size * hm_size;
    float* norm = h->scene_tmp;   // [groups + n_mask_samples][hm_size^2] normalised heightmaps
    {
        const size_t total = (size_t)(groups + n_mask_samples) * px;
        const int blocks = (int)((total + 255) / 256);
        normalize_hm_kernel<<<blocks < h->num_sms * 8 ? blocks : h->num_sms * 8, 256, 0, st>>>(scene_hm, groups, mask_hm, n_mask_samples,
                                                                                            px, mean, stddev, norm);
        h->launches++;
    }
    const int n_samples = groups * n_rot + n_mask_samples;
    const int quads = h->H * h->H / 4;
    prep_rotate_kernel<<<dim3((quads + 255) / 256, n_samples), 256, 0, st>>>(
        norm, groups, n_rot, groups * n_rot, n_samples, th, hm_size, (float)((0.0 - mean) / stddev), out, h->H,
        2.0f / (float)(h->H - 1), (float)(h->H - 1) / 2.0f);
    h->launches++;
    SMG_CUDA(cudaGetLastError());
    return SMG_OK;
}

int launch_rotate_index_map(smg_handle* h, int rot, int num_rot, int32_t* out, cudaStream_t st) {
    RotTheta th;
    rotation_theta(rot, num_rot, th.t[0]);
    const int total = h->H * h->H;
    rotate_index_kernel<<<(total + 255) / 256, 256, 0, st>>>(th, out, h->H);
    h->launches++;
    SMG_CUDA(cudaGetLastError());
    return SMG_OK;
}

// Soft-mask resize of the detector front end (/root/reference/code/masks.py:51): F.interpolate(masks, size=[224, 224],
// mode="bilinear", align_corners=True) of float32 [n,1,448,448] masks.  torch's float32 recipe: scale = (in-1)/(out-1);
// src = scale * dst; i0 = (int)src; lambda1 = src - i0; lambda0 = 1 - lambda1; out = l0y*(l0x*p00 + l1x*p01) + l1y*(l0x*p10 + l1x*p11).
__global__ void resize_bilinear_ac_kernel(const float* __restrict__ in, float* __restrict__ out, int n, int hin, int hout) {
    const float scale = hout > 1 ? (float)(hin - 1) / (float)(hout - 1) : 0.f;
    const size_t total = (size_t)n * hout * hout;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % hout), y = (int)((i / hout) % hout);
        const size_t m = i / ((size_t)hout * hout);
        const float sy = scale * (float)y, sx = scale * (float)x;
        const int y0 = (int)sy, x0 = (int)sx;
        const int yp = y0 < hin - 1 ? 1 : 0, xp = x0 < hin - 1 ? 1 : 0;
        const float ly1 = sy - (float)y0, ly0 = 1.f - ly1, lx1 = sx - (float)x0, lx0 = 1.f - lx1;
        const float* p = in + m * hin * hin + (size_t)y0 * hin + x0;
        const float top = __fadd_rn(__fmul_rn(lx0, p[0]), __fmul_rn(lx1, p[xp]));
        const float bot = __fadd_rn(__fmul_rn(lx0, p[(size_t)yp * hin]), __fmul_rn(lx1, p[(size_t)yp * hin + xp]));
        out[i] = __fadd_rn(__fmul_rn(ly0, top), __fmul_rn(ly1, bot));
    }
}

int launch_resize_masks(smg_handle* h, const float* in, int n, int hin, int hout, float* out, cudaStream_t st) {
    const size_t total = (size_t)n * hout * hout;
    if (total == 0) return SMG_OK;
    int blocks = (int)((total + 255) / 256);
    if (blocks > h->num_sms * 16) blocks = h->num_sms * 16;
    resize_bilinear_ac_kernel<<<blocks, 256, 0, st>>>(in, out, n, hin, hout);
    h->launches++;
    SMG_CUDA(cudaGetLastError());
    return SMG_OK;
}

}  // namespace smg
