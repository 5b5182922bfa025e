// backward.cu - gradient kernels of the training step (Trainer.backprop,
// /root/reference/code/trainer.py:278-384: one grad-enabled forward at a given rotation, loss.backward()).
//
// The reference differentiates through both trunk passes (rotated scene and masked scene share the
// weights) and the head with autograd; every BatchNorm is in train mode, so each BN backward needs the two
// per-(sample, channel) reductions  S1 = sum(dz), S2 = sum(dz * xhat)  before it can produce
//     dx = gamma * rstd * (dz - S1/n - xhat * S2/n),  dgamma = sum_s S2,  dbeta = sum_s S1.
// Round-1 implementation: fp32 CUDA-core kernels (exactness first).  Per dense layer, in reverse order:
//   3x3 dgrad (conv_ffma on flipped weights) -> BN2 reduce -> BN2 apply (-> dy1) -> 3x3 wgrad -> 1x1 wgrad
//   -> 1x1 dgrad -> BN1 reduce -> BN1 apply (accumulates into the block gradient buffer).
// The activations a = relu(bn(x)) are never stored: every kernel recomputes them from the raw buffers and
// the (sum, sumsq) statistics saved by the forward pass.
#include "umma_common.cuh"

namespace smg {

// ------------------------------------------------------------------------------------------------
// BatchNorm(+ReLU) backward: reduce and apply
// ------------------------------------------------------------------------------------------------

__device__ __forceinline__ void bn_channel_consts(const BnBwd& a, int s, int c, float& mean, float& rstd, float& sc,
                                                  float& sh) {
    const double cnt = (double)a.hw * a.hw;
    const double* st = a.stats + 2 * ((size_t)s * a.stats_stride + c);
    const double m = st[0] / cnt;
    double var = st[1] / cnt - m * m;
    if (var < 0) var = 0;
    const double r = 1.0 / sqrt(var + (double)kBnEps);
    mean = (float)m;
    rstd = (float)r;
    sc = a.gamma[c] * rstd;
    sh = a.beta[c] - mean * sc;
}

__device__ __forceinline__ float bn_da(const BnBwd& a, const float* da_s, int p, int c) {
    if (a.da_pooled) {
        const int y = p / a.hw, x = p - y * a.hw;
        const int hp = a.hw >> 1;
        return 0.25f * da_s[((size_t)(y >> 1) * hp + (x >> 1)) * a.da_cstride + c];
    }
    return da_s[(size_t)p * a.da_cstride + c];
}

// One CTA = a chunk of pixels of one sample x ALL channels.  A thread owns four consecutive channels (float4 loads and
// stores) and every PL-th pixel of the chunk; Q = C/4 channel quads x PL = 256/Q pixel lanes.
//   APPLY = 0: S1 += dz, S2 += dz * xhat per channel -> shared-memory combine over the pixel lanes -> one double atomic
//              pair per channel and CTA.
//   APPLY = 1: dx = gamma * rstd * (dz - S1/n - xhat * S2/n); the CTA (0, sample 0) also writes the parameter gradients
//              dgamma = sum_s S2, dbeta = sum_s S1 (the reductions are complete once this kernel runs).
template <int APPLY>
__global__ void __launch_bounds__(256)
bn_bwd_kernel(BnBwd a, int S, float* __restrict__ dgamma, float* __restrict__ dbeta) {
    extern __shared__ __align__(16) float4 red[];   // APPLY = 0: [2][PL][Q]
    pdl_launch_dependents();                        // see conv_umma.cu: the backward chain is launch-latency bound
    pdl_wait();
    const int s = blockIdx.y;
    const int Q = a.C >> 2;
    const int PL = 256 / Q;
    const int q = threadIdx.x % Q, pl = threadIdx.x / Q;
    const bool active = pl < PL;
    const int npix = a.hw * a.hw;
    const int p0 = blockIdx.x * a.pix_per_cta;
    const int p1 = min(p0 + a.pix_per_cta, npix);
    const int hw_da = a.da_pooled ? (a.hw >> 1) : a.hw;
    const float* da_s = a.da + (size_t)s * hw_da * hw_da * a.da_cstride + 4 * q;
    const float* x_s = a.x + (size_t)s * npix * a.x_cstride + 4 * q;
    const float inv_n = 1.0f / (float)npix;
    float mean[4], rstd[4], sc[4], sh[4], m1[4], m2[4], gr[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        bn_channel_consts(a, s, 4 * q + j, mean[j], rstd[j], sc[j], sh[j]);
        m1[j] = m2[j] = gr[j] = 0.f;
        if (APPLY) {
            const double* sm = a.sums + 2 * ((size_t)s * a.C + 4 * q + j);
            m1[j] = (float)(sm[0] * (double)inv_n);
            m2[j] = (float)(sm[1] * (double)inv_n);
            gr[j] = a.gamma[4 * q + j] * rstd[j];
        }
    }
    if (APPLY && blockIdx.x == 0 && s == 0 && pl == 0 && dgamma != nullptr) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double g = 0, b = 0;
            for (int t = 0; t < S; ++t) {
                b += a.sums[2 * ((size_t)t * a.C + 4 * q + j)];
                g += a.sums[2 * ((size_t)t * a.C + 4 * q + j) + 1];
            }
            dgamma[4 * q + j] = (float)g;
            dbeta[4 * q + j] = (float)b;
        }
    }
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
    if (active) {
        float* dst_s = APPLY ? a.dst + (size_t)s * npix * a.dst_cstride + 4 * q : nullptr;
        // four pixels per iteration: their loads are issued together (a loop of dependent 3-load iterations made the small
        // late-block launches pure latency)
        for (int pb = p0 + pl; pb < p1; pb += 4 * PL) {
            float4 xq[4], dq4[4], oq[4];
            bool ok[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int p = pb + u * PL;
                ok[u] = p < p1;
                const int pc = ok[u] ? p : p0;
                xq[u] = *reinterpret_cast<const float4*>(x_s + (size_t)pc * a.x_cstride);
                if (a.da_pooled) {
                    const int y = pc / a.hw, x = pc - y * a.hw;
                    dq4[u] = *reinterpret_cast<const float4*>(da_s + ((size_t)(y >> 1) * hw_da + (x >> 1)) * a.da_cstride);
                } else {
                    dq4[u] = *reinterpret_cast<const float4*>(da_s + (size_t)pc * a.da_cstride);
                }
                if (APPLY && a.accumulate) oq[u] = *reinterpret_cast<const float4*>(dst_s + (size_t)pc * a.dst_cstride);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (!ok[u]) continue;
                const int p = pb + u * PL;
                const float sc_d = a.da_pooled ? 0.25f : 1.f;
                const float xv[4] = {xq[u].x, xq[u].y, xq[u].z, xq[u].w};
                float dz[4] = {dq4[u].x * sc_d, dq4[u].y * sc_d, dq4[u].z * sc_d, dq4[u].w * sc_d};
                float o[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (a.relu && !(fmaf(xv[j], sc[j], sh[j]) > 0.f)) dz[j] = 0.f;
                    const float xh = (xv[j] - mean[j]) * rstd[j];
                    o[j] = gr[j] * (dz[j] - m1[j] - xh * m2[j]);
                    s1[j] += dz[j];
                    s2[j] = fmaf(dz[j], xh, s2[j]);
                }
                if (APPLY) {
                    float4 r = make_float4(o[0], o[1], o[2], o[3]);
                    if (a.accumulate) { r.x += oq[u].x; r.y += oq[u].y; r.z += oq[u].z; r.w += oq[u].w; }
                    *reinterpret_cast<float4*>(dst_s + (size_t)p * a.dst_cstride) = r;
                }
            }
        }
    }
    if (!APPLY) {
        if (active) {
            red[pl * Q + q] = make_float4(s1[0], s1[1], s1[2], s1[3]);
            red[(PL + pl) * Q + q] = make_float4(s2[0], s2[1], s2[2], s2[3]);
        }
        __syncthreads();
        if (pl == 0) {
            double t1[4] = {0, 0, 0, 0}, t2[4] = {0, 0, 0, 0};
            for (int i = 0; i < PL; ++i) {
                const float4 u = red[i * Q + q], v = red[(PL + i) * Q + q];
                t1[0] += u.x; t1[1] += u.y; t1[2] += u.z; t1[3] += u.w;
                t2[0] += v.x; t2[1] += v.y; t2[2] += v.z; t2[3] += v.w;
            }
            double* sm = a.sums + 2 * ((size_t)s * a.C + 4 * q);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                atomicAdd(sm + 2 * j, t1[j]);
                atomicAdd(sm + 2 * j + 1, t2[j]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// weight gradient: dW[o][c][tap] += sum_p G[p][o] * A[p (+tap shift)][c],  A = pool?(relu?(x*sc+sh))
// ------------------------------------------------------------------------------------------------

template <int OT, int TAPS, int POOL>
__global__ void __launch_bounds__(256)
wgrad_kernel(Wgrad a) {
    __shared__ __align__(16) float Gs[32][OT + 4];
    __shared__ __align__(16) float As[32][36];
    __shared__ __align__(16) float s_sc[32];
    __shared__ __align__(16) float s_sh[32];
    constexpr int ON = OT / 16;
    const int tid = threadIdx.x;
    const int c0 = blockIdx.x * 32;
    const int ot = blockIdx.y / TAPS, tap = blockIdx.y - ot * TAPS;
    const int o0 = ot * OT;
    const int s = blockIdx.z / a.chunks_per_sample;
    const int chunk = blockIdx.z - s * a.chunks_per_sample;
    const int npix = a.hout * a.hout;
    const int p0 = chunk * a.pix_per_cta;
    const int p1 = min(p0 + a.pix_per_cta, npix);
    const int dy = TAPS == 9 ? tap / 3 - 1 : 0, dx = TAPS == 9 ? tap % 3 - 1 : 0;
    if (tid < 32) {
        const int c = c0 + tid;
        if (a.prologue_mode == 0) {
            const double cnt = (double)a.hin * a.hin;
            const double* st = a.stats + 2 * ((size_t)s * a.stats_stride + c);
            const double m = st[0] / cnt;
            double var = st[1] / cnt - m * m;
            if (var < 0) var = 0;
            const float sc = a.gamma[c] * (float)(1.0 / sqrt(var + (double)kBnEps));
            s_sc[tid] = sc;
            s_sh[tid] = a.beta[c] - (float)m * sc;
        } else {
            s_sc[tid] = a.scale[(size_t)s * a.cin + c];
            s_sh[tid] = a.shift[(size_t)s * a.cin + c];
        }
    }
    __syncthreads();
    const float* g_s = a.g + (size_t)s * npix * a.g_cstride + a.g_coff + o0;
    const float* x_s = a.x + (size_t)s * a.hin * a.hin * a.x_cstride + c0;
    const int ty = tid >> 4, tx = tid & 15;
    float acc[ON][2];
#pragma unroll
    for (int i = 0; i < ON; ++i) acc[i][0] = acc[i][1] = 0.f;
    // loader roles
    const int apx = tid >> 3, aq = tid & 7;
    const float4 sc4 = *reinterpret_cast<const float4*>(&s_sc[aq * 4]);
    const float4 sh4 = *reinterpret_cast<const float4*>(&s_sh[aq * 4]);
    for (int pb = p0; pb < p1; pb += 32) {
        // G tile [32][OT]
        for (int idx = tid; idx < 32 * (OT / 4); idx += 256) {
            const int px = idx / (OT / 4), q = idx - px * (OT / 4);
            const int p = pb + px;
            float4 v = make_float4(0, 0, 0, 0);
            if (p < p1) v = *reinterpret_cast<const float4*>(g_s + (size_t)p * a.g_cstride + q * 4);
            *reinterpret_cast<float4*>(&Gs[px][q * 4]) = v;
        }
        // A tile [32][32]
        {
            const int p = pb + apx;
            float4 v = make_float4(0, 0, 0, 0);
            if (p < p1) {
                const int y = p / a.hout + dy, x = p % a.hout + dx;
                if (y >= 0 && y < a.hout && x >= 0 && x < a.hout) {
                    if (POOL) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float4 xv = *reinterpret_cast<const float4*>(
                                x_s + ((size_t)(2 * y + (q >> 1)) * a.hin + 2 * x + (q & 1)) * a.x_cstride + aq * 4);
                            float t0 = fmaf(xv.x, sc4.x, sh4.x), t1 = fmaf(xv.y, sc4.y, sh4.y);
                            float t2 = fmaf(xv.z, sc4.z, sh4.z), t3 = fmaf(xv.w, sc4.w, sh4.w);
                            if (a.relu) { t0 = fmaxf(t0, 0.f); t1 = fmaxf(t1, 0.f); t2 = fmaxf(t2, 0.f); t3 = fmaxf(t3, 0.f); }
                            v.x += t0; v.y += t1; v.z += t2; v.w += t3;
                        }
                        v.x *= 0.25f; v.y *= 0.25f; v.z *= 0.25f; v.w *= 0.25f;
                    } else {
                        const float4 xv = *reinterpret_cast<const float4*>(x_s + ((size_t)y * a.hin + x) * a.x_cstride + aq * 4);
                        v.x = fmaf(xv.x, sc4.x, sh4.x); v.y = fmaf(xv.y, sc4.y, sh4.y);
                        v.z = fmaf(xv.z, sc4.z, sh4.z); v.w = fmaf(xv.w, sc4.w, sh4.w);
                        if (a.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                    }
                }
            }
            *reinterpret_cast<float4*>(&As[apx][aq * 4]) = v;
        }
        __syncthreads();
#pragma unroll 8
        for (int px = 0; px < 32; ++px) {
            const float a0 = As[px][tx * 2], a1 = As[px][tx * 2 + 1];
#pragma unroll
            for (int i = 0; i < ON; ++i) {
                const float gv = Gs[px][ty * ON + i];
                acc[i][0] = fmaf(gv, a0, acc[i][0]);
                acc[i][1] = fmaf(gv, a1, acc[i][1]);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < ON; ++i) {
        const int o = o0 + ty * ON + i;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int c = c0 + tx * 2 + j;
            atomicAdd(a.dw + ((size_t)o * a.k_total + a.k_off + c) * TAPS + tap, acc[i][j]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// stem backward
// ------------------------------------------------------------------------------------------------
// maxpool 3x3/2 pad 1 of relu(bn0(conv0)): route each pooled gradient to the first maximum of its window
__global__ void __launch_bounds__(256)
pool0_bwd_kernel(const float* __restrict__ g, int g_cstride, const float* __restrict__ conv0,
                 const double* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ beta,
                 float* __restrict__ da0, int Hc) {
    __shared__ float s_sc[64], s_sh[64];
    const int s = blockIdx.y;
    const int tid = threadIdx.x;
    const int Hp = Hc / 2;
    if (tid < 64) {
        const double cnt = (double)Hc * Hc;
        const double* st = stats + 2 * ((size_t)s * 64 + tid);
        const double m = st[0] / cnt;
        double var = st[1] / cnt - m * m;
        if (var < 0) var = 0;
        const float sc = gamma[tid] * (float)(1.0 / sqrt(var + (double)kBnEps));
        s_sc[tid] = sc;
        s_sh[tid] = beta[tid] - (float)m * sc;
    }
    __syncthreads();
    const int cg = tid & 15, pl = tid >> 4;
    const int p = blockIdx.x * 16 + pl;
    if (p >= Hp * Hp) return;
    const int py = p / Hp, px = p - py * Hp;
    const float* cin = conv0 + (size_t)s * Hc * Hc * 64;
    float* dout = da0 + (size_t)s * Hc * Hc * 64;
    const float4 gv = *reinterpret_cast<const float4*>(g + ((size_t)s * Hp * Hp + p) * g_cstride + cg * 4);
    const float gg[4] = {gv.x, gv.y, gv.z, gv.w};
    float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    int arg[4] = {-1, -1, -1, -1};
    for (int dy = 0; dy < 3; ++dy) {
        const int y = 2 * py - 1 + dy;
        if (y < 0 || y >= Hc) continue;
        for (int dx = 0; dx < 3; ++dx) {
            const int x = 2 * px - 1 + dx;
            if (x < 0 || x >= Hc) continue;
            const float4 v = *reinterpret_cast<const float4*>(cin + ((size_t)y * Hc + x) * 64 + cg * 4);
            const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float r = fmaxf(fmaf(vv[j], s_sc[cg * 4 + j], s_sh[cg * 4 + j]), 0.f);
                if (r > best[j]) { best[j] = r; arg[j] = y * Hc + x; }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (arg[j] >= 0 && best[j] > 0.f) atomicAdd(dout + (size_t)arg[j] * 64 + cg * 4 + j, gg[j]);
}

// dW0[o][c][kh][kw] += sum_p d[p][o] * in[c][2y+kh-3][2x+kw-3]; thread = one (c,kh,kw), 64 accumulators
__global__ void __launch_bounds__(160)
conv0_wgrad_kernel(const float* __restrict__ d, const float* __restrict__ in, float* __restrict__ dw, int H, int cin,
                   int pix_per_cta) {
    __shared__ __align__(16) float Ds[32][64];
    const int s = blockIdx.y;
    const int tid = threadIdx.x;
    const int Ho = H / 2;
    const int K = cin * 49;
    const int p0 = blockIdx.x * pix_per_cta, p1 = min(p0 + pix_per_cta, Ho * Ho);
    // K = cin * 49 taps; with one input channel three pixel lanes share the 147 busy threads
    const int nl = cin == 1 ? 3 : 1;
    const int lane_p = tid / K;
    const int k = lane_p < nl ? tid - lane_p * K : K;
    const int c = k / 49, kh = (k % 49) / 7, kw = k % 7;
    const float* d_s = d + (size_t)s * Ho * Ho * 64;
    const float* in_s = in + (size_t)s * cin * H * H;
    float acc[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) acc[i] = 0.f;
    for (int pb = p0; pb < p1; pb += 32) {
        for (int idx = tid; idx < 32 * 16; idx += 160) {
            const int px = idx >> 4, q = idx & 15;
            float4 v = make_float4(0, 0, 0, 0);
            if (pb + px < p1) v = *reinterpret_cast<const float4*>(d_s + (size_t)(pb + px) * 64 + q * 4);
            *reinterpret_cast<float4*>(&Ds[px][q * 4]) = v;
        }
        __syncthreads();
        if (k < K) {
            for (int px = lane_p; px < 32 && pb + px < p1; px += nl) {
                const int p = pb + px;
                const int y = 2 * (p / Ho) + kh - 3, x = 2 * (p % Ho) + kw - 3;
                if (y < 0 || y >= H || x < 0 || x >= H) continue;
                const float av = in_s[((size_t)c * H + y) * H + x];
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const float4 dv = *reinterpret_cast<const float4*>(&Ds[px][q * 4]);
                    acc[4 * q + 0] = fmaf(av, dv.x, acc[4 * q + 0]);
                    acc[4 * q + 1] = fmaf(av, dv.y, acc[4 * q + 1]);
                    acc[4 * q + 2] = fmaf(av, dv.z, acc[4 * q + 2]);
                    acc[4 * q + 3] = fmaf(av, dv.w, acc[4 * q + 3]);
                }
            }
        }
        __syncthreads();
    }
    if (k < K) {
#pragma unroll
        for (int o = 0; o < 64; ++o) atomicAdd(dw + (size_t)o * K + k, acc[o]);
    }
}

// ------------------------------------------------------------------------------------------------
// head backward
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float sum16(const float (*r)[64], int c) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) t += r[i][c];
    return t;
}

// one CTA: BN(64)+ReLU+20x20 conv backward for the (scene, mask) pair.  p: [2][npix][64] partial products.
__global__ void __launch_bounds__(1024)
head_tail_bwd_kernel(const float* __restrict__ p, int npix, const float* __restrict__ g1, const float* __restrict__ b1,
                     const float* __restrict__ w1, int n_out, const float* __restrict__ dq, float* __restrict__ dP,
                     float* __restrict__ dg1, float* __restrict__ db1, float* __restrict__ dw1) {
    __shared__ float red[2][16][64];   // 64 channels x 16 pixel groups
    __shared__ float s_dq[4];
    const int tid = threadIdx.x, c = tid & 63, g = tid >> 6;
    const float* ps = p;
    const float* pm = p + (size_t)npix * 64;
    if (tid < 4) s_dq[tid] = tid < n_out ? dq[tid] : 0.f;
    float su = 0.f;
    for (int px = g; px < npix; px += 16) su += ps[px * 64 + c] + pm[px * 64 + c];
    red[0][g][c] = su;
    __syncthreads();
    const float mean = sum16(red[0], c) / (float)npix;
    float sq = 0.f;
    for (int px = g; px < npix; px += 16) {
        const float d = ps[px * 64 + c] + pm[px * 64 + c] - mean;
        sq = fmaf(d, d, sq);
    }
    red[1][g][c] = sq;
    __syncthreads();
    const float var = sum16(red[1], c) / (float)npix;
    const float rstd = rsqrtf(var + kBnEps);
    const float gam = g1[c], bet = b1[c];
    __syncthreads();
    float s1 = 0.f, s2 = 0.f;
    for (int px = g; px < npix; px += 16) {
        const float yh = (ps[px * 64 + c] + pm[px * 64 + c] - mean) * rstd;
        const float act = fmaxf(fmaf(gam, yh, bet), 0.f);
        float da = 0.f;
        for (int o = 0; o < n_out; ++o) {
            da = fmaf(s_dq[o], w1[((size_t)o * npix + px) * 64 + c], da);
            dw1[((size_t)o * 64 + c) * npix + px] = s_dq[o] * act;   // torch layout [o][64][20][20]
        }
        const float dz = act > 0.f ? da : 0.f;
        s1 += dz;
        s2 = fmaf(dz, yh, s2);
    }
    red[0][g][c] = s1;
    red[1][g][c] = s2;
    __syncthreads();
    const float S1 = sum16(red[0], c);
    const float S2 = sum16(red[1], c);
    if (g == 0) { dg1[c] = S2; db1[c] = S1; }
    const float m1 = S1 / (float)npix, m2 = S2 / (float)npix;
    for (int px = g; px < npix; px += 16) {
        const float yh = (ps[px * 64 + c] + pm[px * 64 + c] - mean) * rstd;
        const float act = fmaxf(fmaf(gam, yh, bet), 0.f);
        float da = 0.f;
        for (int o = 0; o < n_out; ++o) da = fmaf(s_dq[o], w1[((size_t)o * npix + px) * 64 + c], da);
        const float dz = act > 0.f ? da : 0.f;
        dP[px * 64 + c] = gam * rstd * (dz - m1 - yh * m2);
    }
}

// backward through head norm0 (+ReLU) and the trunk's norm5 for the two samples; CTA = 32 channels x 8 pixel lanes
__device__ __forceinline__ void lane8_sum2(float (*red)[8][32], int pl, int cl, float& a, float& b) {
    red[0][pl][cl] = a;
    red[1][pl][cl] = b;
    __syncthreads();
    float ta = 0.f, tb = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { ta += red[0][i][cl]; tb += red[1][i][cl]; }
    __syncthreads();
    a = ta;
    b = tb;
}

__global__ void __launch_bounds__(256)
head_norm_bwd_kernel(const float* __restrict__ da0, const float* __restrict__ x4, const double* __restrict__ stats,
                     int stats_stride, int npix, const float* __restrict__ g5, const float* __restrict__ b5,
                     const float* __restrict__ gh, const float* __restrict__ bh, float* __restrict__ dx4,
                     float* __restrict__ dg5, float* __restrict__ db5, float* __restrict__ dgh, float* __restrict__ dbh) {
    __shared__ float red[2][8][32];
    const int cl = threadIdx.x & 31, pl = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl;
    float acc_g5 = 0.f, acc_b5 = 0.f;
    const float n = (float)npix;
    for (int s = 0; s < 2; ++s) {
        const double* st = stats + 2 * ((size_t)s * stats_stride + c);
        const double md = st[0] / (double)npix;
        double var = st[1] / (double)npix - md * md;
        if (var < 0) var = 0;
        const double r5d = 1.0 / sqrt(var + (double)kBnEps);
        const double varz = (double)g5[c] * g5[c] * var * r5d * r5d;
        const float mean = (float)md, r5 = (float)r5d, rh = (float)(1.0 / sqrt(varz + (double)kBnEps));
        const float gam5 = g5[c], gamh = gh[s * kFeatC + c], beth = bh[s * kFeatC + c];
        const float kap = gam5 * rh;  // zhat = kap * xhat
        const float* xs = x4 + (size_t)s * npix * kFeatC + c;
        const float* ds = da0 + (size_t)s * npix * kFeatC + c;
        float* os = dx4 + (size_t)s * npix * kFeatC + c;
        float A1 = 0.f, A2 = 0.f;
        for (int p = pl; p < npix; p += 8) {
            const float xh = (xs[(size_t)p * kFeatC] - mean) * r5;
            const float zh = kap * xh;
            const float u = fmaf(gamh, zh, beth);
            const float du = u > 0.f ? ds[(size_t)p * kFeatC] : 0.f;
            A1 += du;
            A2 = fmaf(du, zh, A2);
        }
        lane8_sum2(red, pl, cl, A1, A2);
        if (pl == 0) {
            dgh[s * kFeatC + c] = A2;
            dbh[s * kFeatC + c] = A1;
        }
        const float a1 = A1 / n, a2 = A2 / n;
        float B1 = 0.f, B2 = 0.f;
        for (int p = pl; p < npix; p += 8) {
            const float xh = (xs[(size_t)p * kFeatC] - mean) * r5;
            const float zh = kap * xh;
            const float u = fmaf(gamh, zh, beth);
            const float du = u > 0.f ? ds[(size_t)p * kFeatC] : 0.f;
            const float dz = gamh * rh * (du - a1 - zh * a2);
            B1 += dz;
            B2 = fmaf(dz, xh, B2);
        }
        lane8_sum2(red, pl, cl, B1, B2);
        acc_g5 += B2;
        acc_b5 += B1;
        const float b1 = B1 / n, b2 = B2 / n;
        for (int p = pl; p < npix; p += 8) {
            const float xh = (xs[(size_t)p * kFeatC] - mean) * r5;
            const float zh = kap * xh;
            const float u = fmaf(gamh, zh, beth);
            const float du = u > 0.f ? ds[(size_t)p * kFeatC] : 0.f;
            const float dz = gamh * rh * (du - a1 - zh * a2);
            os[(size_t)p * kFeatC] = gam5 * r5 * (dz - b1 - xh * b2);
        }
    }
    if (pl == 0) {
        dg5[c] = acc_g5;
        db5[c] = acc_b5;
    }
}

// ------------------------------------------------------------------------------------------------
// launch helpers
// ------------------------------------------------------------------------------------------------
int launch_bn_bwd(smg_handle* h, BnBwd a, int S, bool apply, float* dgamma, float* dbeta, cudaStream_t st) {
    SMG_CHECK(a.C % 4 == 0 && a.C >= 4 && a.C <= 1024 && a.da_cstride % 4 == 0 && a.x_cstride % 4 == 0 &&
                  (!apply || a.dst_cstride % 4 == 0),
              SMG_ERR_INVALID, "bn_bwd: C %d / strides must be multiples of 4 (C <= 1024)", a.C);
    const int npix = a.hw * a.hw;
    const int Q = a.C / 4, PL = 256 / Q;
    // pixels per CTA: at least 4 per pixel lane (one batch of loads), and enough CTAs to fill the GPU twice
    int ppc = 2048;
    while (ppc > 4 * PL && (npix + ppc - 1) / ppc * S < 2 * h->num_sms) ppc >>= 1;
    a.pix_per_cta = ppc;
    dim3 grid((npix + ppc - 1) / ppc, S);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(256);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = h->use_pdl ? 1 : 0;
    float* nullp = nullptr;
    if (apply) {
        SMG_CUDA(cudaLaunchKernelEx(&cfg, bn_bwd_kernel<1>, a, S, dgamma, dbeta));
    } else {
        cfg.dynamicSmemBytes = 2 * PL * Q * sizeof(float4);
        SMG_CUDA(cudaLaunchKernelEx(&cfg, bn_bwd_kernel<0>, a, S, nullp, nullp));
    }
    h->launches++;
    SMG_CUDA(cudaGetLastError());
    return SMG_OK;
}

int launch_wgrad(smg_handle* h, Wgrad a, int S, int taps, int pool, cudaStream_t st) {
    SMG_CHECK(a.cin % 32 == 0 && a.cout % 32 == 0, SMG_ERR_INVALID, "wgrad: cin %d / cout %d", a.cin, a.cout);
    const int npix = a.hout * a.hout;
    int ppc = 1024;
    const int ot = a.cout % 64 == 0 ? 64 : 32;
    const int tiles = (a.cin / 32) * (a.cout / ot) * taps;
    while (ppc > 64 && (npix + ppc - 1) / ppc * S * tiles < 2 * h->num_sms) ppc >>= 1;
    a.pix_per_cta = ppc;
    a.chunks_per_sample = (npix + ppc - 1) / ppc;
    dim3 grid(a.cin / 32, (a.cout / ot) * taps, S * a.chunks_per_sample);
    if (taps == 9) {
        SMG_CHECK(!pool, SMG_ERR_INVALID, "wgrad: pooled 3x3 unsupported");
        if (ot == 64) wgrad_kernel<64, 9, 0><<<grid, 256, 0, st>>>(a);
        else wgrad_kernel<32, 9, 0><<<grid, 256, 0, st>>>(a);
    } else if (pool) {
        SMG_CHECK(ot == 64, SMG_ERR_INVALID, "wgrad: pooled conv needs cout %% 64 == 0");
        wgrad_kernel<64, 1, 1><<<grid, 256, 0, st>>>(a);
    } else {
        if (ot == 64) wgrad_kernel<64, 1, 0><<<grid, 256, 0, st>>>(a);
        else wgrad_kernel<32, 1, 0><<<grid, 256, 0, st>>>(a);
    }
    h->launches++;
    SMG_CUDA(cudaGetLastError());
    return SMG_OK;
}

int launch_pool0_bwd(smg_handle* h, int S, const float* g, int g_cstride, const float* conv0, const double* stats,
                     const float* gamma, const float* beta, float* da0, cudaStream_t st) {
    const int Hc = h->H / 2, Hp = Hc / 2;
    dim3 grid((Hp * Hp + 15) / 16, S);
    pool0_bwd_kernel<<<grid, 256, 0, st>>>(g, g_cstride, conv0, stats, gamma, beta, da0, Hc);
    h->launches++;
    SMG_CUDA(cudaGetLastError());
    return SMG_OK;
}

__global__ void replicate_conv0_grad_kernel(const float* __restrict__ g1, float* __restrict__ g3) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;   // over [64][3][49]
    if (i < 64 * 147) g3[i] = g1[(i / 147) * 49 + i % 49];
}

int launch_replicate_conv0_grad(smg_handle* h, const float* g1, float* g3, cudaStream_t st) {
    replicate_conv0_grad_kernel<<<(64 * 147 + 255) / 256, 256, 0, st>>>(g1, g3);
    h->launches++;
    SMG_CUDA(cudaGetLastError());
    return SMG_OK;
}

int launch_conv0_wgrad(smg_handle* h, int S, const float* d, const float* in, int cin, float* dw, cudaStream_t st) {
    const int Ho = h->H / 2;
    const int ppc = 512;
    dim3 grid((Ho * Ho + ppc - 1) / ppc, S);
    conv0_wgrad_kernel<<<grid, 160, 0, st>>>(d, in, dw, h->H, cin, ppc);
    h->launches++;
    SMG_CUDA(cudaGetLastError());
    return SMG_OK;
}

int launch_head_tail_bwd(smg_handle* h, const float* p, const HeadW& hw, const float* dq, float* dP, float* dg1,
                         float* db1, float* dw1, cudaStream_t st) {
    const int npix = h->geom[3].hw * h->geom[3].hw;
    head_tail_bwd_kernel<<<1, 1024, 0, st>>>(p, npix, hw.norm1.gamma, hw.norm1.beta, hw.conv1, hw.n_out, dq, dP, dg1, db1, dw1);
    h->launches++;
    SMG_CUDA(cudaGetLastError());
    return SMG_OK;
}

int launch_head_norm_bwd(smg_handle* h, const float* da0, const float* x4, const double* stats, int stats_stride,
                         const BnP& norm5, const BnP& hnorm0, float* dx4, float* dg5, float* db5, float* dgh, float* dbh,
                         cudaStream_t st) {
    const int npix = h->geom[3].hw * h->geom[3].hw;
    head_norm_bwd_kernel<<<kFeatC / 32, 256, 0, st>>>(da0, x4, stats, stats_stride, npix, norm5.gamma, norm5.beta,
                                                      hnorm0.gamma, hnorm0.beta, dx4, dg5, db5, dgh, dbh);
    h->launches++;
    SMG_CUDA(cudaGetLastError());
    return SMG_OK;
}

}  // namespace smg
