// wgrad_umma.cu - weight gradients of the dense layers' convolutions on the tensor cores (tf32, tcgen05 + TMEM + TMA).
//
// Trainer.backprop (/root/reference/code/trainer.py:350-351) differentiates every convolution of the trunk w.r.t. its
// weights: dW[co][ci][tap] = sum over samples and pixels of  G[p][co] * A[p + tap shift][ci],  A = relu(bn(x)) recomputed
// from the raw activation and the saved (sum, sumsq) statistics.  As a GEMM the contraction index K is the PIXEL, and both
// operands sit in memory pixel-major with the channel contiguous (NHWC), i.e. "MN-major" in UMMA terms.  tcgen05 reads
// MN-major tf32 operands directly (instruction-descriptor bits 15/16) from the shared-memory image a tensor-map TMA box
// {32 channels x P pixels} produces with the 128-byte swizzle of 32-byte atomicity (the only layout 32-bit MN-major operands
// may use: descriptor layout type SWIZZLE_128B_BASE32B; TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): one 128-byte row per pixel
// whose 32-byte quarters are XOR-ed with (row mod 4), 4-row atoms of 512 bytes, 32-channel column blocks side by side
// (descriptor: leading byte offset = block stride, stride byte offset = 512).  So the kernel needs no transposition at all:
//
//   1x1 (conv1, cin -> 128):  D[co 128][ci tile <= 256] += G^T [co][pixel] * A[pixel][ci]
//        M operand = 4 blocks of the output gradient (raw), N operand = up to 8 blocks of the activation (BN-ReLU applied
//        in place by the transform warps), accumulator = the torch [cout][cin] layout as it is.
//   3x3 (conv2, 128 -> 32):   D[ci 128][(tap, co) 288] += A^T[ci][pixel] * G[pixel - shift(tap)][co]
//        M operand = 4 blocks of the activation (transformed in place), N operand = nine 32-channel blocks of the output
//        gradient, block t fetched by a 4-D TMA box displaced by tap t (the copy engine zero-fills outside the image, which
//        is the convolution's zero padding seen from the gradient side); two MMAs per K step (N = 160 + 128).
//
// K is split over CTAs (every CTA owns a contiguous range of 40-pixel units); partial accumulators are added to global
// memory with red.global.add (1x1: straight into the caller's gradient tensor; 3x3: into a [tap][co][ci] scratch that
// wgrad3_finish_kernel transposes into torch's [co][ci][3][3]).
// Warps (320 threads): 0 TMA loader, 1 MMA issuer, 2-5 in-place transform, 6-9 epilogue (TMEM lane quadrant = warp & 3).
#include "tma_common.cuh"

namespace smg {

namespace {

constexpr int G_P = 40;                       // pixels per K unit (5 MMA K steps of 8)
constexpr int G_BLK = G_P * 128;              // one 32-channel block of a unit: 40 rows x 128 B (5 swizzle atoms)
constexpr int G_MAXB = 13;                    // blocks per stage: 4 (M operand) + 9 (N operand, 3x3) / 8 (1x1)
constexpr int G_STAGE = G_MAXB * G_BLK;       // 66 560 B
constexpr int G_NSTAGE = 3;
constexpr int G_OFF_SC = G_NSTAGE * G_STAGE;  // scale[1024], shift[1024]
constexpr int G_OFF_BAR = G_OFF_SC + 8192;
constexpr int G_TOTAL = G_OFF_BAR + 256;
constexpr int G_THREADS = 320;
static_assert(G_TOTAL <= 232448, "shared-memory plan exceeds the 227 KB of one SM");
static_assert(G_BLK % 1024 == 0, "blocks must keep the 1024-byte swizzle phase");

struct WgradDev {
    int hw, S, units_per_sample, total_units, units_per_cta;
    int tw, th;                 // 3x3: the 40 pixels of a unit as a tw x th rectangle (tw * th = 40)
    int cin;                    // 1x1: channels of the activation operand (N); tiles of 256
    const double* stats;        // (sum, sumsq) of the raw activation: [S][stats_stride] double2
    int stats_stride;
    const float* gamma;
    const float* beta;
    float* dw;                  // 1x1: [128][ld] gradient tensor; 3x3: scratch [288][128]
    int ld;
};

// MN-major 32-bit operand, 128-byte swizzle with 32-byte atomicity (layout type 1): 32-channel blocks `lbo` bytes apart,
// 4-pixel atoms 512 bytes apart
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t saddr, uint32_t lbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;
    return d;
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int TAPS>
__global__ void __launch_bounds__(G_THREADS, 1)
wgrad_umma_kernel(const __grid_constant__ CUtensorMap tmM, const __grid_constant__ CUtensorMap tmN, WgradDev a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + G_OFF_BAR);
    uint64_t* raw_full = bars;           // [3] all blocks of the stage landed
    uint64_t* ready = bars + 3;          // [3] activation blocks normalised (128 transform threads)
    uint64_t* empty = bars + 6;          // [3] the MMAs reading the stage retired
    uint64_t* acc_full = bars + 9;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 10);
    float* s_sc = reinterpret_cast<float*>(smem + G_OFF_SC);
    float* s_sh = s_sc + 1024;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int n0 = TAPS == 9 ? 0 : blockIdx.y * 256;                         // first activation channel of this N tile
    const int ntile = TAPS == 9 ? 288 : min(256, a.cin - n0);               // accumulator columns
    const int nb_n = TAPS == 9 ? 9 : ntile / 32;                             // blocks of the N operand per stage
    const int u0 = blockIdx.x * a.units_per_cta;
    const int u1 = min(u0 + a.units_per_cta, a.total_units);
    const int nunits = u1 - u0;
    constexpr uint32_t TCOLS = TAPS == 9 ? 512 : 256;

    if (warp == 1 && lane == 0) {
        if (smem_u32(smem) & 1023u) __trap();
        for (int i = 0; i < G_NSTAGE; ++i) { mbar_init(&raw_full[i], 1); mbar_init(&ready[i], 128); mbar_init(&empty[i], 1); }
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 6) tmem_alloc(tmem_ptr, TCOLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = *tmem_ptr;

    if (warp == 0) {
        // =============================== loader ===============================
        if (lane == 0) {
            for (int i = 0; i < nunits; ++i) {
                const int u = u0 + i;
                const int s = u / a.units_per_sample;
                const int p0 = (u - s * a.units_per_sample) * G_P;
                const int slot = i % G_NSTAGE;
                uint8_t* st = smem + slot * G_STAGE;
                mbar_wait_sleep(&empty[slot], ((i / G_NSTAGE) & 1) ^ 1, 32);
                mbar_arrive_expect_tx(&raw_full[slot], (uint32_t)(4 + nb_n) * G_BLK);
#pragma unroll
                for (int b = 0; b < 4; ++b) tma_tile_3d(st + b * G_BLK, &tmM, b * 32, p0, s, &raw_full[slot]);
                if (TAPS == 9) {
                    const int y0 = p0 / a.hw, x0 = p0 - y0 * a.hw;
#pragma unroll
                    for (int t = 0; t < 9; ++t)   // block t = the gradient displaced by tap t (zero outside the image)
                        tma_tile_4d(st + (4 + t) * G_BLK, &tmN, 0, x0 - (t % 3 - 1), y0 - (t / 3 - 1), s, &raw_full[slot]);
                } else {
                    for (int b = 0; b < nb_n; ++b) tma_tile_3d(st + (4 + b) * G_BLK, &tmN, n0 + b * 32, p0, s, &raw_full[slot]);
                }
            }
        }
    } else if (warp == 1) {
        // =============================== MMA issuer ===============================
        if (lane == 0) {
            // D[128 x N] += M-operand[128 x 8]^T * N-operand[N x 8]^T, both MN-major (bits 15, 16), tf32, fp32 accumulate
            const uint32_t ibase = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t smem_u = smem_u32(smem);
            uint32_t accum = 0;
            for (int i = 0; i < nunits; ++i) {
                const int slot = i % G_NSTAGE;
                mbar_wait(&ready[slot], (i / G_NSTAGE) & 1);
                tc_fence_after();
                const uint32_t st = smem_u + slot * G_STAGE;
#pragma unroll
                for (int ks = 0; ks < G_P / 8; ++ks) {
                    const uint64_t md = make_desc_mn_sw128(st + ks * 1024, G_BLK);
                    if (TAPS == 9) {
                        umma<4>(tmem_d, md, make_desc_mn_sw128(st + 4 * G_BLK + ks * 1024, G_BLK),
                                ibase | ((uint32_t)(160 >> 3) << 17), accum);
                        umma<4>(tmem_d + 160, md, make_desc_mn_sw128(st + 9 * G_BLK + ks * 1024, G_BLK),
                                ibase | ((uint32_t)(128 >> 3) << 17), accum);
                    } else {
                        umma<4>(tmem_d, md, make_desc_mn_sw128(st + 4 * G_BLK + ks * 1024, G_BLK),
                                ibase | ((uint32_t)(ntile >> 3) << 17), accum);
                    }
                    accum = 1;
                }
                umma_commit(&empty[slot]);
            }
            umma_commit(acc_full);
        }
    } else if (warp < 6) {
        // =============================== in-place transform of the activation blocks ===============================
        // 1x1: the N operand's blocks (channels n0 + 32 b + ..); 3x3: the M operand's four blocks (channels 32 b + ..)
        const int t = tid - 64;                   // 0..127
        const int j = t & 7;                      // physical 16-byte piece of the 128-byte row
        const int r0 = t >> 3;                    // rows r0, r0 + 16, r0 + 32 (< 40)
        const int tb0 = TAPS == 9 ? 0 : 4, ntb = TAPS == 9 ? 4 : nb_n;
        const int cbase = TAPS == 9 ? 0 : n0;
        const int nch = ntb * 32;
        int cur_s = -1;
        for (int i = 0; i < nunits; ++i) {
            const int u = u0 + i;
            const int s = u / a.units_per_sample;
            const int slot = i % G_NSTAGE;
            if (s != cur_s) {
                asm volatile("bar.sync 2, 128;" ::: "memory");
                const double inv = 1.0 / ((double)a.hw * a.hw);
                for (int c = t; c < nch; c += 128) {
                    const double2 sv = *reinterpret_cast<const double2*>(a.stats + 2 * ((size_t)s * a.stats_stride + cbase + c));
                    const double m = sv.x * inv;
                    double var = sv.y * inv - m * m;
                    if (var < 0) var = 0;
                    const float sc = a.gamma[cbase + c] * (float)(1.0 / sqrt(var + (double)kBnEps));
                    s_sc[c] = sc;
                    s_sh[c] = a.beta[cbase + c] - (float)m * sc;
                }
                asm volatile("bar.sync 2, 128;" ::: "memory");
                cur_s = s;
            }
            mbar_wait_sleep(&raw_full[slot], (i / G_NSTAGE) & 1, 32);
            uint8_t* st = smem + slot * G_STAGE + tb0 * G_BLK;
            for (int b = 0; b < ntb; ++b) {
#pragma unroll
                for (int rr = 0; rr < 3; ++rr) {
                    const int r = r0 + 16 * rr;
                    if (r < G_P) {
                        // 32-byte quarters are XOR-ed with (row mod 4): logical 4-channel chunk held by this 16-byte piece
                        const int chunk = ((((j >> 1) ^ (r & 3)) << 1) | (j & 1));
                        const float4 sc = *reinterpret_cast<const float4*>(s_sc + b * 32 + chunk * 4);
                        const float4 sh = *reinterpret_cast<const float4*>(s_sh + b * 32 + chunk * 4);
                        float4* p = reinterpret_cast<float4*>(st + b * G_BLK + r * 128 + j * 16);
                        float4 x = *p;
                        x.x = fmaxf(fmaf(x.x, sc.x, sh.x), 0.f); x.y = fmaxf(fmaf(x.y, sc.y, sh.y), 0.f);
                        x.z = fmaxf(fmaf(x.z, sc.z, sh.z), 0.f); x.w = fmaxf(fmaf(x.w, sc.w, sh.w), 0.f);
                        *p = x;
                    }
                }
            }
            fence_proxy_async();
            mbar_arrive(&ready[slot]);
        }
    } else {
        // =============================== epilogue ===============================
        const int q4 = warp & 3;                  // TMEM lane quadrant: rows 32 q4 + lane of the accumulator
        const int row = q4 * 32 + lane;
        if (nunits > 0) {
            mbar_wait_sleep(acc_full, 0, 64);
            tc_fence_after();
            for (int cb = 0; cb < ntile; cb += 32) {
                float v[32];
                tmem_ld32(tmem_d + ((uint32_t)(q4 * 32) << 16) + (uint32_t)cb, v);
                if (TAPS == 9) {
                    // scratch [(tap, co)][ci]: lanes = ci are contiguous -> one coalesced 128-byte reduction per column
#pragma unroll
                    for (int i = 0; i < 32; ++i) atomicAdd(a.dw + (size_t)(cb + i) * 128 + row, v[i]);
                } else {
                    float* o = a.dw + (size_t)row * a.ld + n0 + cb;     // torch [cout][cin]: this thread's 32 consecutive inputs
#pragma unroll
                    for (int i = 0; i < 32; i += 4) red_add_v4(o + i, v[i], v[i + 1], v[i + 2], v[i + 3]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 6) {
        tc_fence_after();
        tmem_dealloc(tmem_d, TCOLS);
    }
}

// scratch [58][(tap, co) 288][ci 128] -> torch [co 32][ci 128][3][3] of every dense layer, one launch
struct FinishJob {
    const float* src;
    float* dst;
};
__global__ void wgrad3_finish_kernel(const FinishJob* __restrict__ jobs) {
    const FinishJob j = jobs[blockIdx.y];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 32 * 128 * 9; i += gridDim.x * blockDim.x) {
        const int tap = i % 9, ci = (i / 9) % 128, co = i / (9 * 128);
        j.dst[i] = j.src[(size_t)(tap * 32 + co) * 128 + ci];
    }
}

}  // namespace

// dW of a dense layer's 1x1 convolution: g = gradient w.r.t. its output [S,hw,hw,128] (dense), x = raw block buffer
// [S,hw,hw,x_cstride] (channels [0,cin) used), BN-ReLU prologue from (stats, gamma, beta); dw [128][cin] is ACCUMULATED into.
int launch_wgrad1_umma(smg_handle* h, const float* g, const float* x, int x_cstride, int cin, int hw, int S, const double* stats,
                       int stats_stride, const float* gamma, const float* beta, float* dw, cudaStream_t st) {
    const int npix = hw * hw;
    if (npix % G_P != 0 || cin % 32 != 0 || cin > 1024 || x_cstride % 4 != 0) return SMG_ERR_UNSUPPORTED;
    CUtensorMap tmM, tmN;
    {
        const cuuint64_t dims[3] = {128, (cuuint64_t)npix, (cuuint64_t)S};
        const cuuint64_t strides[2] = {128 * 4, (cuuint64_t)npix * 128 * 4};
        const cuuint32_t box[3] = {32, G_P, 1};
        SMG_TRY(make_tensor_map_f32(&tmM, g, 3, dims, strides, box, kSwizzle128Atom32));
    }
    {
        const cuuint64_t dims[3] = {(cuuint64_t)x_cstride, (cuuint64_t)npix, (cuuint64_t)S};
        const cuuint64_t strides[2] = {(cuuint64_t)x_cstride * 4, (cuuint64_t)npix * x_cstride * 4};
        const cuuint32_t box[3] = {32, G_P, 1};
        SMG_TRY(make_tensor_map_f32(&tmN, x, 3, dims, strides, box, kSwizzle128Atom32));
    }
    WgradDev d{};
    d.hw = hw; d.S = S; d.units_per_sample = npix / G_P; d.total_units = S * d.units_per_sample;
    d.cin = cin; d.stats = stats; d.stats_stride = stats_stride; d.gamma = gamma; d.beta = beta; d.dw = dw; d.ld = cin;
    const int ntiles = (cin + 255) / 256;
    int splits = (h->wgrad_cta_cap > 0 ? h->wgrad_cta_cap : h->num_sms) / ntiles;
    if (splits > d.total_units / 2) splits = d.total_units / 2;      // at least two units per CTA
    if (splits < 1) splits = 1;
    d.units_per_cta = (d.total_units + splits - 1) / splits;
    splits = (d.total_units + d.units_per_cta - 1) / d.units_per_cta;
    SMG_TRY(ensure_dyn_smem(h, (const void*)wgrad_umma_kernel<1>, G_TOTAL));
    wgrad_umma_kernel<1><<<dim3(splits, ntiles), G_THREADS, G_TOTAL, st>>>(tmM, tmN, d);
    h->launches++;
    SMG_CUDA(cudaGetLastError());
    return SMG_OK;
}

// dW of a dense layer's 3x3 convolution (128 -> 32): g = block gradient buffer [S,hw,hw,g_cstride], channels
// [g_coff, g_coff + 32); y = raw bottleneck activation [S,hw,hw,128]; scratch [288][128] is ACCUMULATED into.
int launch_wgrad3_umma(smg_handle* h, const float* g, int g_cstride, int g_coff, const float* y, int hw, int S,
                       const double* stats, int stats_stride, const float* gamma, const float* beta, float* scratch,
                       cudaStream_t st) {
    const int npix = hw * hw;
    if (npix % G_P != 0 || g_cstride % 4 != 0 || g_coff % 4 != 0 || (hw % G_P != 0 && G_P % hw != 0)) return SMG_ERR_UNSUPPORTED;
    CUtensorMap tmM, tmN;
    {
        const cuuint64_t dims[3] = {128, (cuuint64_t)npix, (cuuint64_t)S};
        const cuuint64_t strides[2] = {128 * 4, (cuuint64_t)npix * 128 * 4};
        const cuuint32_t box[3] = {32, G_P, 1};
        SMG_TRY(make_tensor_map_f32(&tmM, y, 3, dims, strides, box, kSwizzle128Atom32));
    }
    WgradDev d{};
    d.hw = hw; d.S = S; d.units_per_sample = npix / G_P; d.total_units = S * d.units_per_sample;
    d.tw = hw >= G_P ? G_P : hw;
    d.th = G_P / d.tw;
    {
        const cuuint64_t dims[4] = {32, (cuuint64_t)hw, (cuuint64_t)hw, (cuuint64_t)S};
        const cuuint64_t strides[3] = {(cuuint64_t)g_cstride * 4, (cuuint64_t)hw * g_cstride * 4, (cuuint64_t)npix * g_cstride * 4};
        const cuuint32_t box[4] = {32, (cuuint32_t)d.tw, (cuuint32_t)d.th, 1};
        SMG_TRY(make_tensor_map_f32(&tmN, g + g_coff, 4, dims, strides, box, kSwizzle128Atom32));
    }
    d.cin = 128; d.stats = stats; d.stats_stride = stats_stride; d.gamma = gamma; d.beta = beta; d.dw = scratch; d.ld = 128;
    int splits = h->wgrad_cta_cap > 0 ? h->wgrad_cta_cap : h->num_sms;
    if (splits > d.total_units / 2) splits = d.total_units / 2;
    if (splits < 1) splits = 1;
    d.units_per_cta = (d.total_units + splits - 1) / splits;
    splits = (d.total_units + d.units_per_cta - 1) / d.units_per_cta;
    SMG_TRY(ensure_dyn_smem(h, (const void*)wgrad_umma_kernel<9>, G_TOTAL));
    wgrad_umma_kernel<9><<<dim3(splits, 1), G_THREADS, G_TOTAL, st>>>(tmM, tmN, d);
    h->launches++;
    SMG_CUDA(cudaGetLastError());
    return SMG_OK;
}

// transposes every layer's [tap][co][ci] scratch into the caller's [co][ci][3][3] tensors; `jobs` is a device table
int launch_wgrad3_finish(smg_handle* h, const void* dev_jobs, int n_jobs, cudaStream_t st) {
    wgrad3_finish_kernel<<<dim3(8, n_jobs), 256, 0, st>>>(reinterpret_cast<const FinishJob*>(dev_jobs));
    h->launches++;
    SMG_CUDA(cudaGetLastError());
    return SMG_OK;
}

}  // namespace smg
