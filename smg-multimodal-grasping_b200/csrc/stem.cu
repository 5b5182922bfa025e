// stem.cu - K2/K4: densenet `features.conv0` (7x7 stride 2 pad 3, 3->64) and
// `norm0 + relu0 + pool0` (train-mode BN, ReLU, 3x3 stride-2 max pool), torchvision
// densenet.py as called at /root/reference/code/models.py:384-385.
//
// conv0 is a K=147 (K=49 when the three input channels are identical) direct convolution on the
// CUDA cores (4% of the trunk's MACs, Cin=3 does not tile onto the tensor cores); it writes the raw
// output NHWC and accumulates
// the per-(sample,channel) sum / sum-of-squares that norm0 needs.  pool0 applies
// norm0+ReLU on the fly, max-pools, writes channels [0,64) of the dense-block-1 buffer
// and accumulates THEIR statistics (needed by every norm1 of block 1).
#include "smg_internal.cuh"

namespace smg {

constexpr int C0_TILE = 16;                       // 16x16 output pixels per CTA
constexpr int C0_PATCH = 2 * C0_TILE + 5;         // 37 input rows/cols
constexpr int C0_PLD = 49;                        // padded patch row (2*PLD mod 32 == 2: <= 2-way bank conflicts)
constexpr int C0_STAGE_LD = 65;                   // padded row of the output staging tile

// CIN = 3: the general case (any [n,3,H,H] input).  CIN = 1: the three input channels are identical
// (Trainer.forward replicates the depth map, code/trainer.py:178-181), so the 7x7 weights are summed
// over the input channel once at pack time and K drops from 147 to 49.
// thread = 4 consecutive output pixels x 16 output channels: per (c,kh) it loads the 13 input values its
// 4 pixels need once and 7 x 4 broadcast LDS.128 of weights feed 7 x 64 FMAs (FMA-bound, not LDS-bound).
template <int CIN>
__global__ void __launch_bounds__(256, 2)
conv0_kernel(const float* __restrict__ in, const float* __restrict__ w, float* __restrict__ out,
             double* __restrict__ stats, int H, int stats_stride) {
    constexpr int K = CIN * 49;
    extern __shared__ float sm[];
    float* s_w = sm;                                   // [K][64]
    float* s_in = s_w + K * 64;                        // [CIN][37][40]
    float* s_out = sm;                                 // [256][65] staging, aliases weights + patch after the MACs
    const int Ho = H / 2;
    const int tiles = Ho / C0_TILE;
    const int s = blockIdx.y;
    const int ty = blockIdx.x / tiles, tx = blockIdx.x % tiles;
    const int tid = threadIdx.x;

    for (int i = tid; i < K * 64; i += 256) s_w[i] = w[i];
    const int iy0 = ty * C0_TILE * 2 - 3, ix0 = tx * C0_TILE * 2 - 3;
    const float* inp = in + (size_t)s * CIN * H * H;
    for (int i = tid; i < CIN * C0_PATCH * C0_PLD; i += 256) {
        const int c = i / (C0_PATCH * C0_PLD);
        const int r = i - c * C0_PATCH * C0_PLD;
        const int py = r / C0_PLD, px = r - py * C0_PLD;
        const int y = iy0 + py, x = ix0 + px;
        float v = 0.f;
        if (px < C0_PATCH && y >= 0 && y < H && x >= 0 && x < H) v = inp[((size_t)c * H + y) * H + x];
        s_in[i] = v;
    }
    __syncthreads();

    const int cg = tid >> 6;            // 16-channel group (warp-uniform)
    const int pg = tid & 63;            // pixel group: row oy, 4 pixels starting at ox0
    const int oy = pg >> 2, ox0 = (pg & 3) * 4;
    float acc[4][16];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[p][j] = 0.f;
    for (int c = 0; c < CIN; ++c) {
        for (int kh = 0; kh < 7; ++kh) {
            const float* irow = s_in + (c * C0_PATCH + (2 * oy + kh)) * C0_PLD + 2 * ox0;
            float iv[13];
#pragma unroll
            for (int t = 0; t < 13; ++t) iv[t] = irow[t];
            const float* wrow = s_w + ((c * 7 + kh) * 7) * 64 + cg * 16;
#pragma unroll
            for (int kw = 0; kw < 7; ++kw) {
                const float4* w4 = reinterpret_cast<const float4*>(wrow + kw * 64);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 ww = w4[j];
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        const float v = iv[2 * p + kw];
                        acc[p][4 * j + 0] = fmaf(v, ww.x, acc[p][4 * j + 0]);
                        acc[p][4 * j + 1] = fmaf(v, ww.y, acc[p][4 * j + 1]);
                        acc[p][4 * j + 2] = fmaf(v, ww.z, acc[p][4 * j + 2]);
                        acc[p][4 * j + 3] = fmaf(v, ww.w, acc[p][4 * j + 3]);
                    }
                }
            }
        }
    }
    __syncthreads();  // all reads of s_w / s_in done: the staging tile may overwrite them
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int j = 0; j < 16; ++j) s_out[(oy * C0_TILE + ox0 + p) * C0_STAGE_LD + cg * 16 + j] = acc[p][j];
    __syncthreads();
    float* outp = out + (size_t)s * Ho * Ho * 64;
    for (int i = tid; i < 256 * 64; i += 256) {
        const int p = i >> 6, c = i & 63;
        const int py = p / C0_TILE, px = p - py * C0_TILE;
        outp[((size_t)(ty * C0_TILE + py) * Ho + tx * C0_TILE + px) * 64 + c] = s_out[p * C0_STAGE_LD + c];
    }
    // column sums: thread = (channel c, row group g of 64 pixels)
    {
        const int c = tid & 63, g = tid >> 6;
        float su = 0.f, sq = 0.f;
        for (int p = g * 64; p < g * 64 + 64; ++p) {
            const float v = s_out[p * C0_STAGE_LD + c];
            su += v;
            sq = fmaf(v, v, sq);
        }
        __syncthreads();
        float* red = sm + 256 * C0_STAGE_LD;  // behind the staging tile
        red[tid] = su;
        red[256 + tid] = sq;
        __syncthreads();
        if (tid < 64) {
            const double a = (double)red[tid] + (double)red[64 + tid] + (double)red[128 + tid] + (double)red[192 + tid];
            const double b = (double)red[256 + tid] + (double)red[320 + tid] + (double)red[384 + tid] +
                             (double)red[448 + tid];
            double* st = stats + 2 * ((size_t)s * stats_stride + tid);
            atomicAdd(st, a);
            atomicAdd(st + 1, b);
        }
    }
}

// norm0 + ReLU + maxpool 3x3/2 pad 1.  thread = (pixel, 4-channel group); CTA = 16 pixel lanes x 16 groups.
__global__ void __launch_bounds__(256)
pool0_kernel(const float* __restrict__ conv0, const double* __restrict__ stats_in, int stats_in_stride,
             const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ out,
             int out_cstride, double* __restrict__ stats_out, int stats_out_stride, int Hc, int pix_per_cta) {
    __shared__ __align__(16) float s_sc[64];
    __shared__ __align__(16) float s_sh[64];
    __shared__ double s_sum[16][64], s_sq[16][64];
    const int s = blockIdx.y;
    const int tid = threadIdx.x;
    const int Hp = Hc / 2;
    if (tid < 64) {
        const double* st = stats_in + 2 * ((size_t)s * stats_in_stride + tid);
        const double cnt = (double)Hc * Hc;
        const double m = st[0] / cnt;
        double var = st[1] / cnt - m * m;
        if (var < 0) var = 0;
        const float sc = gamma[tid] * (float)(1.0 / sqrt(var + (double)kBnEps));
        s_sc[tid] = sc;
        s_sh[tid] = beta[tid] - (float)m * sc;
    }
    __syncthreads();
    const int cg = tid & 15, pl = tid >> 4;
    const float4 sc = *reinterpret_cast<const float4*>(&s_sc[cg * 4]);
    const float4 sh = *reinterpret_cast<const float4*>(&s_sh[cg * 4]);
    const float* cin = conv0 + (size_t)s * Hc * Hc * 64;
    float* o = out + (size_t)s * Hp * Hp * out_cstride;
    // statistics in double from the first addition: the consumers form var = E[x^2] - mean^2, and a channel whose spread is
    // small against its mean (a large BatchNorm bias) loses (mean/std)^2 of the precision of these sums
    double su[4] = {0.0, 0.0, 0.0, 0.0}, sq[4] = {0.0, 0.0, 0.0, 0.0};
    // a CTA owns a 16 x 16 block of pooled pixels (pix_per_cta = 256): its 3x3/2 windows cover 33 x 33 conv0 pixels, 1.06
    // reads per unique input (256 consecutive pixels of a row needed 5 full conv0 rows: 1.4x, ncu: 2.53 GB read for 1.78 GB)
    const int tiles_x = (Hp + 15) >> 4;
    const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
    for (int q = pl; q < pix_per_cta; q += 16) {
        const int py = ty * 16 + (q >> 4), px = tx * 16 + (q & 15);
        if (py >= Hp || px >= Hp) continue;
        const int p = py * Hp + px;
        float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            const int y = 2 * py - 1 + dy;
            if (y < 0 || y >= Hc) continue;
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const int x = 2 * px - 1 + dx;
                if (x < 0 || x >= Hc) continue;
                const float4 v = *reinterpret_cast<const float4*>(cin + ((size_t)y * Hc + x) * 64 + cg * 4);
                m.x = fmaxf(m.x, fmaxf(fmaf(v.x, sc.x, sh.x), 0.f));
                m.y = fmaxf(m.y, fmaxf(fmaf(v.y, sc.y, sh.y), 0.f));
                m.z = fmaxf(m.z, fmaxf(fmaf(v.z, sc.z, sh.z), 0.f));
                m.w = fmaxf(m.w, fmaxf(fmaf(v.w, sc.w, sh.w), 0.f));
            }
        }
        *reinterpret_cast<float4*>(o + (size_t)p * out_cstride + cg * 4) = m;
        const float mv[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            su[j] += (double)mv[j];
            sq[j] = fma((double)mv[j], (double)mv[j], sq[j]);
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        s_sum[pl][cg * 4 + j] = su[j];
        s_sq[pl][cg * 4 + j] = sq[j];
    }
    __syncthreads();
    if (tid < 64) {
        double a = 0, b = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            a += s_sum[i][tid];
            b += s_sq[i][tid];
        }
        double* st = stats_out + 2 * ((size_t)s * stats_out_stride + tid);
        atomicAdd(st, a);
        atomicAdd(st + 1, b);
    }
}

template <int CIN>
static int launch_conv0_t(smg_handle* h, const float* in, int n, const float* w, float* out, double* stats,
                          cudaStream_t st) {
    const int Ho = h->H / 2;
    const size_t operands = (size_t)(CIN * 49 * 64 + CIN * C0_PATCH * C0_PLD);
    const size_t staging = (size_t)256 * C0_STAGE_LD + 512;
    const size_t smem = (operands > staging ? operands : staging) * sizeof(float);
    SMG_TRY(ensure_dyn_smem(h, (const void*)conv0_kernel<CIN>, (int)smem));
    dim3 grid((Ho / C0_TILE) * (Ho / C0_TILE), n);
    conv0_kernel<CIN><<<grid, 256, smem, st>>>(in, w, out, stats, h->H, 64);
    h->launches++;
    SMG_CUDA(cudaGetLastError());
    return SMG_OK;
}

int launch_conv0(smg_handle* h, const float* in, int cin, int n, const float* w, float* out, double* stats,
                 cudaStream_t st) {
    SMG_CHECK((h->H / 2) % C0_TILE == 0, SMG_ERR_INVALID, "conv0: H/2 must be a multiple of %d", C0_TILE);
    if (cin == 1) return launch_conv0_t<1>(h, in, n, w, out, stats, st);
    return launch_conv0_t<3>(h, in, n, w, out, stats, st);
}

int launch_pool0(smg_handle* h, int n, const float* conv0, const double* stats_in, const float* gamma,
                 const float* beta, float* out, int out_cstride, double* stats_out, cudaStream_t st) {
    const int Hc = h->H / 2, Hp = Hc / 2;
    const int pix_per_cta = 256;                 // one 16 x 16 block of pooled pixels
    const int tiles = (Hp + 15) / 16;
    dim3 grid(tiles * tiles, n);
    pool0_kernel<<<grid, 256, 0, st>>>(conv0, stats_in, 64, gamma, beta, out, out_cstride, stats_out, out_cstride, Hc,
                                       pix_per_cta);
    h->launches++;
    SMG_CUDA(cudaGetLastError());
    return SMG_OK;
}

}  // namespace smg
