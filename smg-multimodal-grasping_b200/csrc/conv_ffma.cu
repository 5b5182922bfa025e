// conv_ffma.cu - fp32 CUDA-core implicit-GEMM convolution with fused train-mode BN-ReLU
// prologue and per-channel statistics epilogue.  This is the SMG_PREC_FP32 arithmetic
// mode (parity <= 1e-4 against the fp32 reference needs fp32 operands; the tensor-core
// kernels in conv_umma.cu round operands to tf32 / bf16).
//
// One kernel serves every convolution after the stem (torchvision densenet `_DenseLayer`
// conv1/conv2, `_Transition` conv, and the head's 1x1, /root/reference/code/models.py:319):
//   prologue : a = relu(x*scale + shift), scale/shift from the PRODUCER's (sum, sumsq) and this
//              layer's gamma/beta  (BatchNorm2d in train mode, batch 1, eps 1e-5)
//   POOL     : a = mean of the 2x2 window of prologue outputs (transition avg-pool commuted
//              in front of the 1x1 convolution)
//   TAPS=9   : 3x3, zero padding 1 applied AFTER the prologue (as nn.Conv2d pads its input)
//   epilogue : raw output into its channel slice of the dense-block buffer + (sum, sumsq)
#include "smg_internal.cuh"

namespace smg {

struct ConvDev {
    const float* in;
    int in_cstride, cin, hin;
    int prologue_mode;
    const double* in_stats;
    int in_stats_stride;
    const float* gamma;
    const float* beta;
    const float* scale;
    const float* shift;
    int relu;
    const float* w;  // [taps][cin][cout]
    float* out;
    int out_cstride, out_coff, cout;
    double* out_stats;
    int out_stats_stride;
    int hout;
};

constexpr int FM = 64;   // pixels per CTA
constexpr int FK = 16;   // channels per step

template <int NT, int TAPS, int POOL>
__global__ void __launch_bounds__(256)
conv_ffma_kernel(ConvDev a) {
    extern __shared__ float sm[];
    float* s_sc = sm;                    // [cin]
    float* s_sh = s_sc + a.cin;          // [cin]
    float* As = s_sh + a.cin;            // [FK][FM+4]
    float* Bs = As + FK * (FM + 4);      // [FK][NT]
    float* red = Bs + FK * NT;           // [2][16][NT]
    constexpr int CN = NT / 16;          // columns per thread

    const int tid = threadIdx.x;
    const int s = blockIdx.z;
    const int n0 = blockIdx.y * NT;
    const int m0 = blockIdx.x * FM;
    const int hout = a.hout, hin = a.hin;
    const int hw_out = hout * hout;

    // ---- prologue parameters for this sample
    if (a.prologue_mode == 0) {
        const double cnt = (double)hin * hin;
        for (int c = tid; c < a.cin; c += 256) {
            const double* st = a.in_stats + 2 * ((size_t)s * a.in_stats_stride + c);
            const double m = st[0] / cnt;
            double var = st[1] / cnt - m * m;
            if (var < 0) var = 0;
            const float sc = a.gamma[c] * (float)(1.0 / sqrt(var + (double)kBnEps));
            s_sc[c] = sc;
            s_sh[c] = a.beta[c] - (float)m * sc;
        }
    } else if (a.prologue_mode == 1) {
        for (int c = tid; c < a.cin; c += 256) {
            s_sc[c] = a.scale[(size_t)s * a.cin + c];
            s_sh[c] = a.shift[(size_t)s * a.cin + c];
        }
    } else {
        for (int c = tid; c < a.cin; c += 256) {
            s_sc[c] = 1.f;
            s_sh[c] = 0.f;
        }
    }
    __syncthreads();

    const float* inp = a.in + (size_t)s * hin * hin * a.in_cstride;
    // loader role: row lr, channel quad lq
    const int lr = tid >> 2, lq = tid & 3;
    const int lm = m0 + lr;
    const bool lvalid = lm < hw_out;
    const int ly = lvalid ? lm / hout : 0, lx = lvalid ? lm - ly * hout : 0;
    // compute role
    const int tm = tid >> 4, tn = tid & 15;

    float acc[4][CN];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < CN; ++j) acc[i][j] = 0.f;

    for (int tap = 0; tap < TAPS; ++tap) {
        const int dy = TAPS == 9 ? tap / 3 - 1 : 0;
        const int dx = TAPS == 9 ? tap % 3 - 1 : 0;
        const int iy = ly + dy, ix = lx + dx;
        const bool inb = lvalid && iy >= 0 && iy < hout && ix >= 0 && ix < hout;
        for (int k0 = 0; k0 < a.cin; k0 += FK) {
            // ---- A tile
            {
                const int c = k0 + lq * 4;
                float4 v = make_float4(0, 0, 0, 0);
                if (inb) {
                    const float4 sc = *reinterpret_cast<const float4*>(s_sc + c);
                    const float4 sh = *reinterpret_cast<const float4*>(s_sh + c);
                    if (POOL) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int py = 2 * iy + (q >> 1), px = 2 * ix + (q & 1);
                            const float4 x = *reinterpret_cast<const float4*>(
                                inp + ((size_t)py * hin + px) * a.in_cstride + c);
                            float t0 = fmaf(x.x, sc.x, sh.x), t1 = fmaf(x.y, sc.y, sh.y);
                            float t2 = fmaf(x.z, sc.z, sh.z), t3 = fmaf(x.w, sc.w, sh.w);
                            if (a.relu) { t0 = fmaxf(t0, 0.f); t1 = fmaxf(t1, 0.f); t2 = fmaxf(t2, 0.f); t3 = fmaxf(t3, 0.f); }
                            v.x += t0; v.y += t1; v.z += t2; v.w += t3;
                        }
                        v.x *= 0.25f; v.y *= 0.25f; v.z *= 0.25f; v.w *= 0.25f;
                    } else {
                        const float4 x = *reinterpret_cast<const float4*>(
                            inp + ((size_t)iy * hin + ix) * a.in_cstride + c);
                        v.x = fmaf(x.x, sc.x, sh.x); v.y = fmaf(x.y, sc.y, sh.y);
                        v.z = fmaf(x.z, sc.z, sh.z); v.w = fmaf(x.w, sc.w, sh.w);
                        if (a.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                    }
                }
                As[(lq * 4 + 0) * (FM + 4) + lr] = v.x;
                As[(lq * 4 + 1) * (FM + 4) + lr] = v.y;
                As[(lq * 4 + 2) * (FM + 4) + lr] = v.z;
                As[(lq * 4 + 3) * (FM + 4) + lr] = v.w;
            }
            // ---- B tile  [FK][NT] from w[tap][k0+k][n0+n]
            {
                constexpr int QN = NT / 4;  // float4 per row
                if (tid < FK * QN) {
                    const int k = tid / QN, q = tid - k * QN;
                    const float4 wv = *reinterpret_cast<const float4*>(
                        a.w + ((size_t)tap * a.cin + k0 + k) * a.cout + n0 + q * 4);
                    *reinterpret_cast<float4*>(Bs + k * NT + q * 4) = wv;
                }
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < FK; ++k) {
                const float4 av = *reinterpret_cast<const float4*>(As + k * (FM + 4) + tm * 4);
                float bv[CN];
#pragma unroll
                for (int j = 0; j < CN; ++j) bv[j] = Bs[k * NT + tn * CN + j];
#pragma unroll
                for (int j = 0; j < CN; ++j) {
                    acc[0][j] = fmaf(av.x, bv[j], acc[0][j]);
                    acc[1][j] = fmaf(av.y, bv[j], acc[1][j]);
                    acc[2][j] = fmaf(av.z, bv[j], acc[2][j]);
                    acc[3][j] = fmaf(av.w, bv[j], acc[3][j]);
                }
            }
            __syncthreads();
        }
    }

    // ---- epilogue: store + statistics
    float* outp = a.out + (size_t)s * hw_out * a.out_cstride + a.out_coff + n0 + tn * CN;
    float su[CN], sq[CN];
#pragma unroll
    for (int j = 0; j < CN; ++j) su[j] = sq[j] = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + tm * 4 + i;
        if (m < hw_out) {
#pragma unroll
            for (int j = 0; j < CN; ++j) {
                outp[(size_t)m * a.out_cstride + j] = acc[i][j];
                su[j] += acc[i][j];
                sq[j] = fmaf(acc[i][j], acc[i][j], sq[j]);
            }
        }
    }
    if (a.out_stats != nullptr) {
#pragma unroll
        for (int j = 0; j < CN; ++j) {
            red[tm * NT + tn * CN + j] = su[j];
            red[16 * NT + tm * NT + tn * CN + j] = sq[j];
        }
        __syncthreads();
        if (tid < NT) {
            double x = 0, y = 0;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                x += (double)red[i * NT + tid];
                y += (double)red[16 * NT + i * NT + tid];
            }
            double* st = a.out_stats + 2 * ((size_t)s * a.out_stats_stride + a.out_coff + n0 + tid);
            atomicAdd(st, x);
            atomicAdd(st + 1, y);
        }
    }
}

template <int NT, int TAPS, int POOL>
static int launch_one(smg_handle* h, const ConvDev& d, int n, cudaStream_t st) {
    const size_t smem = (size_t)(2 * d.cin + FK * (FM + 4) + FK * NT + 2 * 16 * NT) * sizeof(float);
    dim3 grid((d.hout * d.hout + FM - 1) / FM, d.cout / NT, n);
    conv_ffma_kernel<NT, TAPS, POOL><<<grid, 256, smem, st>>>(d);
    h->launches++;
    SMG_CUDA(cudaGetLastError());
    return SMG_OK;
}

int launch_conv_ffma(smg_handle* h, const ConvArgs& a, cudaStream_t st) {
    SMG_CHECK(a.w_raw != nullptr || (a.w != nullptr && a.w->w_ffma != nullptr), SMG_ERR_STATE,
              "conv_ffma: weights not packed");
    SMG_CHECK(a.cin % FK == 0 && a.cin <= 2048, SMG_ERR_INVALID, "conv_ffma: cin %d unsupported", a.cin);
    ConvDev d;
    d.in = a.in; d.in_cstride = a.in_cstride; d.cin = a.cin; d.hin = a.hin;
    d.prologue_mode = a.prologue_mode; d.in_stats = a.in_stats; d.in_stats_stride = a.in_stats_stride;
    d.gamma = a.gamma; d.beta = a.beta; d.scale = a.scale; d.shift = a.shift; d.relu = a.relu;
    d.w = a.w_raw ? a.w_raw : a.w->w_ffma; d.out = a.out; d.out_cstride = a.out_cstride; d.out_coff = a.out_coff; d.cout = a.cout;
    d.out_stats = a.out_stats; d.out_stats_stride = a.out_stats_stride;
    d.hout = a.pool ? a.hin / 2 : a.hin;
    if (a.taps == 9) {
        SMG_CHECK(a.cout % 32 == 0 && !a.pool, SMG_ERR_INVALID, "conv_ffma: bad 3x3 config");
        return launch_one<32, 9, 0>(h, d, a.n, st);
    }
    SMG_CHECK(a.cout % 32 == 0, SMG_ERR_INVALID, "conv_ffma: cout %d must be a multiple of 32", a.cout);
    if (a.cout % 64 != 0) {
        SMG_CHECK(!a.pool, SMG_ERR_INVALID, "conv_ffma: pooled conv needs cout %% 64 == 0");
        return launch_one<32, 1, 0>(h, d, a.n, st);
    }
    if (a.pool) return launch_one<64, 1, 1>(h, d, a.n, st);
    return launch_one<64, 1, 0>(h, d, a.n, st);
}

}  // namespace smg
