// head.cu - the Q heads and the action argmax.
//
// Reference: /root/reference/code/models.py:316-343 (RL heads), :28-55 (reactive heads):
//   BN(2048,train) - ReLU - conv1x1 2048->64 - BN(64,train) - ReLU - conv 20x20 valid 64->{1,3}
// applied to cat(trunk(rotated scene), trunk(masked scene)) (code/models.py:386-387).
//
// Decomposition used here (exact algebra, fp32 rounding differs at the 1e-7 level):
//   * channels [0,1024) of the concatenation depend only on the scene sample, [1024,2048) only on
//     the mask sample, and BN statistics are per channel, so BN(2048)+ReLU+conv1x1 splits into
//     P_scene[r] = W[:, :1024] relu(bn(f_scene[r]))  and  P_mask[k] = W[:, 1024:] relu(bn(f_mask[k]));
//     the 1x1 conv of a (rotation r, object k) pair is P_scene[r] + P_mask[k].  The two partial
//     products are computed once per SAMPLE by the generic conv kernel (prologue mode 1).
//   * norm5 (the trunk's last BN, no ReLU) followed by the head's BN is a per-channel affine map of
//     the raw block-4 activations; head_prepare composes both from the (sum, sumsq) statistics.
//   * head_tail: per pair, BN(64) statistics over the 400 pixels, ReLU, 25 600-term dot product.
#include "smg_internal.cuh"

namespace smg {

__global__ void head_prepare_kernel(int n, const double* __restrict__ stats, int stats_stride, double cnt,
                                    const float* __restrict__ g5, const float* __restrict__ b5,
                                    const float* __restrict__ gh, const float* __restrict__ bh,
                                    float* __restrict__ scale, float* __restrict__ shift) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * kFeatC) return;
    const int s = i / kFeatC, c = i - s * kFeatC;
    const double* st = stats + 2 * ((size_t)s * stats_stride + c);
    const double mean = st[0] / cnt;
    double var = st[1] / cnt - mean * mean;
    if (var < 0) var = 0;
    const double rstd5 = 1.0 / sqrt(var + (double)kBnEps);
    // z = g5*(x-mean)*rstd5 + b5 has batch mean b5 and biased variance g5^2 * var * rstd5^2
    const double var_z = (double)g5[c] * g5[c] * var * rstd5 * rstd5;
    const double rstd_h = 1.0 / sqrt(var_z + (double)kBnEps);
    const double sc = (double)gh[c] * rstd_h * (double)g5[c] * rstd5;
    scale[i] = (float)sc;
    shift[i] = (float)((double)bh[c] - mean * sc);
}

__global__ void norm5_export_kernel(int n, const float* __restrict__ x, const double* __restrict__ stats,
                                    int stats_stride, int hw, const float* __restrict__ g5,
                                    const float* __restrict__ b5, float* __restrict__ out) {
    // out NCHW [n,1024,hw,hw]; thread = (s, c, p)
    const size_t total = (size_t)n * kFeatC * hw * hw;
    const double cnt = (double)hw * hw;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int p = (int)(i % (hw * hw));
        const int c = (int)((i / (hw * hw)) % kFeatC);
        const int s = (int)(i / ((size_t)hw * hw * kFeatC));
        const double* st = stats + 2 * ((size_t)s * stats_stride + c);
        const double mean = st[0] / cnt;
        double var = st[1] / cnt - mean * mean;
        if (var < 0) var = 0;
        const float sc = g5[c] * (float)(1.0 / sqrt(var + (double)kBnEps));
        const float sh = b5[c] - (float)mean * sc;
        out[i] = fmaf(x[((size_t)s * hw * hw + p) * kFeatC + c], sc, sh);
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float group_sum(const float (*r)[64], int c) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) t += r[i][c];
    return t;
}

// one CTA per (mask k, rotation r) pair.  p: [n_rot + n_masks][400][64]
__global__ void __launch_bounds__(1024)
head_tail_kernel(const float* __restrict__ p, const float* __restrict__ p_mask, int n_rot, int n_masks, int npix,
                 const float* __restrict__ g1,
                 const float* __restrict__ b1, const float* __restrict__ w1, int n_out, float* __restrict__ q,
                 float* __restrict__ bn1_stats) {
    __shared__ float red[2][16][64];   // 64 channels x 16 pixel groups: 25 pixels per thread and pass
    __shared__ float s_sc[64], s_sh[64];
    __shared__ float s_part[32][4];
    // samples: [groups x n_rot] rotated scenes, then [groups x n_masks] masked scenes; blockIdx.z = group (unit)
    const int r = blockIdx.x, k0 = blockIdx.y, grp = blockIdx.z, groups = gridDim.z;
    const int k = grp * n_masks + k0;   // mask index over all groups == row of q / bn1_stats
    const int tid = threadIdx.x, c = tid & 63, g = tid >> 6;
    const float* ps = p + (size_t)(grp * n_rot + r) * npix * 64;
    const float* pm = p_mask + (size_t)k * npix * 64;   // masked-scene partials: [groups x n_masks], mask index k over all groups
    // pass 1: mean
    float su = 0.f;
    for (int px = g; px < npix; px += 16) su += ps[px * 64 + c] + pm[px * 64 + c];
    red[0][g][c] = su;
    __syncthreads();
    const float mean = group_sum(red[0], c) / (float)npix;
    // pass 2: biased variance around the mean
    float sq = 0.f;
    for (int px = g; px < npix; px += 16) {
        const float d = ps[px * 64 + c] + pm[px * 64 + c] - mean;
        sq = fmaf(d, d, sq);
    }
    red[1][g][c] = sq;
    __syncthreads();
    if (tid < 64) {
        const float var = group_sum(red[1], c) / (float)npix;
        const float sc = g1[c] * rsqrtf(var + kBnEps);
        s_sc[c] = sc;
        s_sh[c] = b1[c] - mean * sc;
        if (bn1_stats != nullptr) {  // batch statistics of the head's BN(64) for the running-stat side effect
            float* o = bn1_stats + ((size_t)k * n_rot + r) * 128;
            o[c] = mean;
            o[64 + c] = var;
        }
    }
    __syncthreads();
    const float sc = s_sc[c], sh = s_sh[c];
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int px = g; px < npix; px += 16) {
        const float y = fmaxf(fmaf(ps[px * 64 + c] + pm[px * 64 + c], sc, sh), 0.f);
#pragma unroll
        for (int o = 0; o < 4; ++o)
            if (o < n_out) acc[o] = fmaf(y, w1[((size_t)o * npix + px) * 64 + c], acc[o]);
    }
    const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
    for (int o = 0; o < 4; ++o) {
        const float v = warp_sum(acc[o]);
        if (lane == 0) s_part[warp][o] = v;
    }
    __syncthreads();
    if (tid < n_out) {
        float v = 0.f;
        for (int w = 0; w < 32; ++w) v += s_part[w][tid];
        q[((size_t)k * n_rot + r) * n_out + tid] = v;
    }
}

// first-max-wins argmax (np.argmax, code/main.py:172-173,195)
__global__ void argmax_kernel(const float* __restrict__ q, int n, float* __restrict__ out, int32_t* __restrict__ out_idx) {
    __shared__ float sv[32];
    __shared__ int si[32];
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float v = q[i];
        if (v > best || (v == best && i < bi)) { best = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x < 32) {
        const int nw = blockDim.x >> 5;
        best = threadIdx.x < nw ? sv[threadIdx.x] : -INFINITY;
        bi = threadIdx.x < nw ? si[threadIdx.x] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (threadIdx.x == 0) { out[0] = best; out_idx[0] = bi; }
    }
}

__global__ void nhwc_to_nchw_kernel(const float* __restrict__ in, int npix, int c, int cstride, float* __restrict__ out) {
    const size_t total = (size_t)npix * c;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int p = (int)(i % npix);
        const int ch = (int)(i / npix);
        out[i] = in[(size_t)p * cstride + ch];
    }
}

// per-sample BN batch statistics (mean, biased var) from (sum, sumsq): block = (region, sample)
__global__ void bn_export_all_kernel(const BnRegion* __restrict__ regions, int total, float* __restrict__ mean,
                                     float* __restrict__ var) {
    const BnRegion r = regions[blockIdx.x];
    const int s = blockIdx.y;
    for (int c = threadIdx.x; c < r.count; c += blockDim.x) {
        const double* st = r.stats + 2 * ((size_t)s * r.stride + c);
        const double m = st[0] / r.cnt;
        double v = st[1] / r.cnt - m * m;
        if (v < 0) v = 0;
        mean[(size_t)s * total + r.out_off + c] = (float)m;
        var[(size_t)s * total + r.out_off + c] = (float)v;
    }
}

int launch_head_prepare(smg_handle* h, int n, const double* stats4, int stats_stride, const BnP& norm5,
                        const BnP& hnorm0, int half, float* scale, float* shift, cudaStream_t st) {
    const int hw = h->geom[3].hw;
    const int total = n * kFeatC;
    head_prepare_kernel<<<(total + 255) / 256, 256, 0, st>>>(n, stats4, stats_stride, (double)hw * hw, norm5.gamma,
                                                             norm5.beta, hnorm0.gamma + half * kFeatC,
                                                             hnorm0.beta + half * kFeatC, scale, shift);
    h->launches++;
    SMG_CUDA(cudaGetLastError());
    return SMG_OK;
}

int launch_norm5_export(smg_handle* h, int n, const float* block4, const double* stats4, int stats_stride,
                        const BnP& norm5, float* out_nchw, cudaStream_t st) {
    const int hw = h->geom[3].hw;
    const size_t total = (size_t)n * kFeatC * hw * hw;
    int blocks = (int)((total + 255) / 256);
    if (blocks > h->num_sms * 8) blocks = h->num_sms * 8;
    norm5_export_kernel<<<blocks, 256, 0, st>>>(n, block4, stats4, stats_stride, hw, norm5.gamma, norm5.beta, out_nchw);
    h->launches++;
    SMG_CUDA(cudaGetLastError());
    return SMG_OK;
}

int launch_head_tail(smg_handle* h, const float* p, const float* p_mask, int n_rot, int n_masks, const HeadW& hw, float* q,
                     cudaStream_t st, int groups) {
    const int npix = h->geom[3].hw * h->geom[3].hw;
    SMG_CHECK(hw.n_out >= 1 && hw.n_out <= 4, SMG_ERR_INVALID, "head_tail: n_out %d", hw.n_out);
    dim3 grid(n_rot, n_masks, groups);
    const bool fits = (size_t)groups * n_rot * n_masks * 128 <= h->head_bn1_floats;
    head_tail_kernel<<<grid, 1024, 0, st>>>(p, p_mask, n_rot, n_masks, npix, hw.norm1.gamma, hw.norm1.beta, hw.conv1, hw.n_out, q,
                                           fits ? h->head_bn1 : nullptr);
    h->head_bn1_pairs = fits ? groups * n_rot * n_masks : 0;
    h->launches++;
    SMG_CUDA(cudaGetLastError());
    return SMG_OK;
}

int launch_argmax(smg_handle* h, const float* q, int n, float* out, int32_t* out_idx, cudaStream_t st) {
    SMG_CHECK(n > 0, SMG_ERR_INVALID, "argmax: empty table");
    argmax_kernel<<<1, 256, 0, st>>>(q, n, out, out_idx);
    h->launches++;
    SMG_CUDA(cudaGetLastError());
    return SMG_OK;
}

int launch_nhwc_to_nchw(smg_handle* h, const float* in, int hw, int c, int cstride, float* out, cudaStream_t st) {
    const size_t total = (size_t)hw * hw * c;
    int blocks = (int)((total + 255) / 256);
    if (blocks > h->num_sms * 8) blocks = h->num_sms * 8;
    nhwc_to_nchw_kernel<<<blocks, 256, 0, st>>>(in, hw * hw, c, cstride, out);
    h->launches++;
    SMG_CUDA(cudaGetLastError());
    return SMG_OK;
}

int launch_bn_export_all(smg_handle* h, int n, const void* dev_regions, int n_regions, int total, float* mean, float* var,
                         cudaStream_t st) {
    bn_export_all_kernel<<<dim3(n_regions, n), 256, 0, st>>>(reinterpret_cast<const BnRegion*>(dev_regions), total, mean, var);
    h->launches++;
    SMG_CUDA(cudaGetLastError());
    return SMG_OK;
}

}  // namespace smg
