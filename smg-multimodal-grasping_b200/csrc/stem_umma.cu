// stem_umma.cu - densenet `features.conv0` (7x7 stride 2 pad 3) on the tensor cores, tf32 mode, identical input channels.
//
// torchvision densenet.py `features.conv0` as called at /root/reference/code/models.py:384-385; Trainer.forward replicates the
// depth map over the three input channels (code/trainer.py:178-181), so the weights are summed over the input channel at
// pack time and the layer is a K = 49 GEMM per output pixel.  On the CUDA cores (stem.cu) that is 21.8 G FMA per 4-unit step
// and FMA-bound (1.8 ms); here it is an implicit GEMM M = 128 output pixels (8 rows x 16 columns) x N = 64 x K = 56:
//   k = kh * 8 + kw with the eighth tap of every kernel row multiplied by a zero weight, so that a 16-byte K chunk is four
//   CONSECUTIVE input floats x[2 oy + kh][2 ox - 3 + 4 half .. + 3] - the im2col operand is built from a 21 x 40 input patch
//   with 4 scalar LDS + 1 STS.128 per chunk.
// Accuracy: the stem feeds 120 BatchNorm layers, so it is kept at fp32 accuracy with the 3xTF32 split: x = x_hi + x_lo
// (x_hi = x with the 13 low mantissa bits cleared, exactly representable in tf32; x_lo = x - x_hi, exact in fp32), likewise
// for the weights at pack time, and D += x_hi w_hi + x_lo w_hi + x_hi w_lo (the dropped x_lo w_lo term is 2^-22 relative).
// 21 MMAs per tile instead of 7 - the tensor pipe is still 80 % idle in this kernel.
// Persistent CTA per SM, 14 warps:
//   warp 9 lane 0    TMA: 3-D boxes (x, y, sample) of the single-channel input, out-of-image taps zero-filled by the copy
//                    engine (= the convolution's zero padding); the 14 KB packed weights once;
//   warps 0-3,10-13  build the UMMA no-swizzle K-major A tile [14 chunks][128 rows][16 B] from the patch;
//   warp 8 lane 0    7 x tcgen05.mma kind::tf32 (M = 128, N = 64) per tile into a double-buffered TMEM accumulator;
//   warps 4-7        epilogue: TMEM -> registers -> raw NHWC output row (256 B per pixel) + (sum, sumsq) by warp
//                    transpose-reduction into per-warp double registers, flushed once per (warp, sample).
#include "tma_common.cuh"

namespace smg {

namespace {

constexpr int S_TR = 8, S_TC = 16;                 // output tile: 8 rows x 16 columns = 128 pixels
constexpr int S_PH = 2 * S_TR + 5;                 // 21 input rows
constexpr int S_PW = 40;                           // input columns 32 tx - 4 .. + 35 (the taps need - 3 .. + 34), box row 160 B
constexpr int S_PATCH = S_PH * S_PW * 4;           // 3360 B
constexpr int S_PSLOT = 3456;                      // patch slot stride (128-byte multiple)
constexpr int S_NP = 4;                            // patch slots
constexpr int S_CH = 14;                           // 16-byte K chunks (K = 56)
constexpr int S_LBO = 129 * 16;                    // padded rows: conflict-free 16-byte stores
constexpr int S_AHALF = 226 * 128;                 // one operand image: >= 14 chunks x 2064 B = 28 896 B, 128-byte multiple
constexpr int S_ASLOT = 2 * S_AHALF;               // hi image + lo image
constexpr int S_NA = 3;                            // A tiles in flight
constexpr int S_WHALF = S_CH * 64 * 16;            // 14 336 B
constexpr int S_WBYTES = 2 * S_WHALF;              // hi image + lo image
constexpr int S_OFF_P = 0;
constexpr int S_OFF_A = S_OFF_P + S_NP * S_PSLOT;
constexpr int S_OFF_W = S_OFF_A + S_NA * S_ASLOT;
constexpr int S_OFF_BAR = S_OFF_W + S_WBYTES;
constexpr int S_TOTAL = S_OFF_BAR + 256;
static_assert(S_OFF_A % 128 == 0 && S_OFF_W % 128 == 0 && S_OFF_BAR % 8 == 0, "alignment");

__global__ void __launch_bounds__(448, 1)
conv0_umma_kernel(const __grid_constant__ CUtensorMap tmX, const uint8_t* __restrict__ w, float* __restrict__ out,
                  double* __restrict__ stats, int stats_stride, int Ho, int total_tiles) {
    constexpr int BN = 64;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S_OFF_BAR);
    uint64_t* p_full = bars;            // [4] patch landed
    uint64_t* p_empty = bars + 4;       // [4] patch consumed by the 256 builders
    uint64_t* a_ready = bars + 8;       // [3] A tile built (256 builders)
    uint64_t* a_empty = bars + 11;      // [3] MMAs retired
    uint64_t* t_full = bars + 14;       // [2]
    uint64_t* t_empty = bars + 16;      // [2] 128 epilogue threads
    uint64_t* w_full = bars + 18;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 19);
    uint8_t* sP = smem + S_OFF_P;
    uint8_t* sA = smem + S_OFF_A;
    uint8_t* sW = smem + S_OFF_W;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int tiles_x = Ho / S_TC, tps = tiles_x * (Ho / S_TR);
    const int tile_begin = (int)(((long long)blockIdx.x * total_tiles) / gridDim.x);
    const int tile_end = (int)(((long long)(blockIdx.x + 1) * total_tiles) / gridDim.x);
    const int ntiles = tile_end - tile_begin;

    if (warp == 8 && lane == 0) {
        for (int i = 0; i < S_NP; ++i) { mbar_init(&p_full[i], 1); mbar_init(&p_empty[i], 256); }
        for (int i = 0; i < S_NA; ++i) { mbar_init(&a_ready[i], 256); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 128); }
        mbar_init(w_full, 1);
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc(tmem_ptr, 2 * BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 9) {
        // =============================== loader ===============================
        if (lane == 0) {
            mbar_arrive_expect_tx(w_full, S_WBYTES);
            tma_bulk_load(sW, w, S_WBYTES, w_full);
            for (int it = 0; it < ntiles; ++it) {
                const int tile = tile_begin + it;
                const int s = tile / tps, rem = tile - s * tps;
                const int ty = rem / tiles_x, tx = rem - ty * tiles_x;
                const int slot = it % S_NP;
                mbar_wait_sleep(&p_empty[slot], ((it / S_NP) & 1) ^ 1, 32);
                mbar_arrive_expect_tx(&p_full[slot], S_PATCH);
                // the innermost box coordinate must keep every box row 16-byte aligned in global memory: start one column early
                tma_tile_3d(sP + slot * S_PSLOT, &tmX, 2 * tx * S_TC - 4, 2 * ty * S_TR - 3, s, &p_full[slot]);
            }
        }
    } else if (warp < 4 || warp >= 10) {
        // =============================== im2col builders ===============================
        const int ptid = warp < 4 ? tid : tid - 192;          // 0..255
        const int m = ptid & 127;                             // tile row (pixel)
        const int r = m >> 4, c = m & 15;
        const int half0 = ptid >> 7;                          // chunk = 2 i + half0, i < 7  ->  kh = i, kw = 4 half0 .. + 3
        for (int it = 0; it < ntiles; ++it) {
            const int ps = it % S_NP, as = it % S_NA;
            mbar_wait_sleep(&p_full[ps], (it / S_NP) & 1, 32);
            mbar_wait_sleep(&a_empty[as], ((it / S_NA) & 1) ^ 1, 32);
            const float* patch = reinterpret_cast<const float*>(sP + ps * S_PSLOT) + (2 * r) * S_PW + 2 * c + 4 * half0 + 1;
            uint8_t* dst = sA + as * S_ASLOT + half0 * S_LBO + m * 16;
#pragma unroll
            for (int kh = 0; kh < 7; ++kh) {
                const float* p = patch + kh * S_PW;
                const float4 v = make_float4(p[0], p[1], p[2], p[3]);
                float4 hi, lo;
                hi.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); lo.x = v.x - hi.x;
                hi.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); lo.y = v.y - hi.y;
                hi.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); lo.z = v.z - hi.z;
                hi.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); lo.w = v.w - hi.w;
                *reinterpret_cast<float4*>(dst + 2 * kh * S_LBO) = hi;
                *reinterpret_cast<float4*>(dst + S_AHALF + 2 * kh * S_LBO) = lo;
            }
            fence_proxy_async();
            mbar_arrive(&a_ready[as]);
            mbar_arrive(&p_empty[ps]);
        }
    } else if (warp == 8) {
        // =============================== MMA issuer ===============================
        if (lane == 0) {
            constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                       ((uint32_t)(UM >> 4) << 24);
            const uint32_t sA_u = smem_u32(sA), sW_u = smem_u32(sW);
            mbar_wait(w_full, 0);
            for (int it = 0; it < ntiles; ++it) {
                const int buf = it & 1, as = it % S_NA;
                mbar_wait(&t_empty[buf], ((it >> 1) & 1) ^ 1);
                mbar_wait(&a_ready[as], (it / S_NA) & 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
#pragma unroll
                for (int k = 0; k < S_CH / 2; ++k) {
                    const uint64_t a_hi = make_desc(sA_u + as * S_ASLOT + 2 * k * S_LBO, S_LBO, 128);
                    const uint64_t a_lo = make_desc(sA_u + as * S_ASLOT + S_AHALF + 2 * k * S_LBO, S_LBO, 128);
                    const uint64_t b_hi = make_desc(sW_u + 2 * k * BN * 16, BN * 16, 128);
                    const uint64_t b_lo = make_desc(sW_u + S_WHALF + 2 * k * BN * 16, BN * 16, 128);
                    umma<4>(d_tmem, a_lo, b_hi, idesc, k > 0 ? 1u : 0u);   // small terms first
                    umma<4>(d_tmem, a_hi, b_lo, idesc, 1u);
                    umma<4>(d_tmem, a_hi, b_hi, idesc, 1u);
                }
                umma_commit(&a_empty[as]);
                umma_commit(&t_full[buf]);
            }
        }
    } else {
        // =============================== epilogue (warps 4-7) ===============================
        const int e = warp - 4;
        const int row = e * 32 + lane;
        const int r = row >> 4, c = row & 15;
        double acc_su[2] = {0.0, 0.0}, acc_ss[2] = {0.0, 0.0};
        int cur_s = -1;
        auto flush = [&](int s_done) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                double* st = stats + 2 * ((size_t)s_done * stats_stride + k * 32 + lane);
                atomicAdd(st, acc_su[k]);
                atomicAdd(st + 1, acc_ss[k]);
                acc_su[k] = 0.0;
                acc_ss[k] = 0.0;
            }
        };
        for (int it = 0; it < ntiles; ++it) {
            const int tile = tile_begin + it;
            const int s = tile / tps, rem = tile - s * tps;
            const int ty = rem / tiles_x, tx = rem - ty * tiles_x;
            const int buf = it & 1;
            if (s != cur_s) {
                if (cur_s >= 0) flush(cur_s);
                cur_s = s;
            }
            float* orow = out + (((size_t)s * Ho + ty * S_TR + r) * Ho + tx * S_TC + c) * 64;
            mbar_wait_sleep(&t_full[buf], (it >> 1) & 1, 64);
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(e * 32) << 16) + (uint32_t)(buf * BN + k * 32), v);
                if (k == 1) {
                    tc_fence_before();
                    mbar_arrive(&t_empty[buf]);
                }
                float4* o = reinterpret_cast<float4*>(orow + k * 32);
#pragma unroll
                for (int q4 = 0; q4 < 8; ++q4) o[q4] = make_float4(v[4 * q4], v[4 * q4 + 1], v[4 * q4 + 2], v[4 * q4 + 3]);
                float sq[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) sq[i] = v[i] * v[i];
                acc_su[k] += (double)warp_transpose_sum(v, lane);
                acc_ss[k] += (double)warp_transpose_sum(sq, lane);
            }
        }
        if (cur_s >= 0) flush(cur_s);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 2 * BN);
    }
}

// [49][64] folded weights (k = kh*7 + kw) -> two UMMA no-swizzle B images (hi, lo) [chunk 14][n 64][4], k' = kh*8 + kw,
// zero for kw = 7
__global__ void pack_conv0_umma_kernel(const float* __restrict__ folded, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < S_CH * 64 * 4) {
        const int e = i & 3, n = (i >> 2) & 63, ch = i >> 8;
        const int kp = ch * 4 + e, kh = kp >> 3, kw = kp & 7;
        const float v = kw < 7 ? folded[(kh * 7 + kw) * 64 + n] : 0.f;
        const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
        out[i] = hi;
        out[S_CH * 64 * 4 + i] = v - hi;
    }
}

}  // namespace

int pack_conv0_umma(smg_handle* h, const float* folded, float* out, cudaStream_t st) {
    pack_conv0_umma_kernel<<<(S_CH * 64 * 4 + 255) / 256, 256, 0, st>>>(folded, out);
    h->launches++;
    SMG_CUDA(cudaGetLastError());
    return SMG_OK;
}

// single-channel input [n][H][H]; returns SMG_ERR_UNSUPPORTED when the tensor-core stem does not apply
int launch_conv0_umma(smg_handle* h, const float* in, int n, const float* w_umma, float* out, double* stats, cudaStream_t st) {
    const int H = h->H, Ho = H / 2;
    if (Ho % S_TC != 0 || Ho % S_TR != 0 || w_umma == nullptr || (reinterpret_cast<uintptr_t>(in) & 15) != 0 || H % 4 != 0)
        return SMG_ERR_UNSUPPORTED;
    EncodeTiledFn enc = encode_tiled_fn();
    SMG_CHECK(enc != nullptr, SMG_ERR_CUDA, "conv0_umma: cuTensorMapEncodeTiled not available from the driver");
    CUtensorMap tm;
    const cuuint64_t dims[3] = {(cuuint64_t)H, (cuuint64_t)H, (cuuint64_t)n};
    const cuuint64_t strides[2] = {(cuuint64_t)H * 4, (cuuint64_t)H * H * 4};
    const cuuint32_t box[3] = {S_PW, S_PH, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(in), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SMG_CHECK(r == CUDA_SUCCESS, SMG_ERR_CUDA, "conv0_umma: cuTensorMapEncodeTiled failed (%d)", (int)r);
    static bool attr = false;
    if (!attr) {
        SMG_CUDA(cudaFuncSetAttribute(conv0_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, S_TOTAL));
        attr = true;
    }
    const int total = (Ho / S_TC) * (Ho / S_TR) * n;
    const int grid = total < h->num_sms ? total : h->num_sms;
    conv0_umma_kernel<<<grid, 448, S_TOTAL, st>>>(tm, reinterpret_cast<const uint8_t*>(w_umma), out, stats, 64, Ho, total);
    h->launches++;
    SMG_CUDA(cudaGetLastError());
    return SMG_OK;
}

}  // namespace smg
