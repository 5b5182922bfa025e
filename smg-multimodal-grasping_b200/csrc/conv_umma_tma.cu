// conv_umma_tma.cu - 1x1 tf32 convolution whose activation operand is fetched by tensor-map TMA.
//
// Same contract and results as conv_umma_kernel<4, 128, 1, 0> (conv_umma.cu): BN-ReLU prologue, raw NHWC output,
// (sum, sumsq) epilogue.  Serves torchvision densenet `_DenseLayer.conv1` (/root/reference/code/models.py:319 builds
// the trunks; the layer itself lives in torchvision).  What changes is how the activation tile reaches shared memory:
//
//   warp 9 lane 0   cp.async.bulk.tensor.3d (tensor map over the block buffer [sample][pixel][channel], box
//                   32 channels x 128 pixels, SWIZZLE_128B, rows beyond the sample zero-filled) straight into the
//                   operand stage, and the bulk copy of the packed weight stage.  No registers, no LSU issue slots and
//                   NA stages (NA x 16 KB) in flight per CTA whatever the other warps are doing.
//   warps 0-3,10-13 wait for the stage's transaction barrier and apply relu(x*scale+shift) IN PLACE: thread t owns
//                   16-byte piece (t mod 8) of rows (t/8 + 32 i); under the 128-byte swizzle that piece holds logical
//                   chunk (t mod 8) xor (row mod 8), which is constant per thread, so scale/shift are read once per stage.
//   warp 8 lane 0   tcgen05.mma kind::tf32, A descriptor SWIZZLE_128B (SBO = 1024 B, K step = +32 B on the start
//                   address), B descriptor unchanged (no-swizzle stage images written by pack.cu).
//   warps 4-7       epilogue as in conv_umma.cu; all 14 warps share the stores and statistics.
#include "tma_common.cuh"

namespace smg {

namespace {

constexpr int NA_T = 4;   // activation stages in flight (16 KB each)
constexpr int NB_T = 2;   // weight stages (16 KB each at N = 128)

template <int BN>
struct TmaPlan {
    static constexpr int A_STAGE = UM * 128;                // 128 rows x 128 B
    static constexpr int B_STAGE = 8 * BN * 16;
    static constexpr int OFF_BAR = 0;
    static constexpr int OFF_SC = 256;
    static constexpr int OFF_A = 9216;                      // 1024-aligned, after 2 x 1024 floats of scale/shift
    static constexpr int OFF_B = OFF_A + NA_T * A_STAGE;
    static constexpr int STAGING = UM * (BN + 1) * 4;
    static constexpr int END_AB = OFF_B + NB_T * B_STAGE;
    static constexpr int USED = (OFF_A + STAGING > END_AB ? OFF_A + STAGING : END_AB);
    static constexpr int TOTAL = USED + 1024;               // slack to align the dynamic window to 1024 B
};

template <int BN>
__global__ void __launch_bounds__(448, 2)
conv_umma_tma_kernel(const __grid_constant__ CUtensorMap tmA, UmmaDev a) {
    using P = TmaPlan<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + P::OFF_BAR);
    uint64_t* raw_full = bars;          // [4] TMA transaction barriers (activations landed, still raw)
    uint64_t* a_ready = bars + 4;       // [4] 256 transform threads
    uint64_t* a_empty = bars + 8;       // [4] MMA commit
    uint64_t* b_full = bars + 12;       // [4]
    uint64_t* b_empty = bars + 16;      // [4]
    uint64_t* tmem_full = bars + 20;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 21);
    float* s_sc = reinterpret_cast<float*>(smem + P::OFF_SC);
    float* s_sh = s_sc + 1024;
    uint8_t* sA = smem + P::OFF_A;
    uint8_t* sB = smem + P::OFF_B;
    float* s_out = reinterpret_cast<float*>(smem + P::OFF_A);

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int s = blockIdx.z;
    const int ntile = blockIdx.y;
    const int hw_out = a.hout * a.hout;
    const int KG = a.cin / KC;
    const int m0 = blockIdx.x * UM;

    if (warp == 8 && lane == 0) {
        for (int i = 0; i < 4; ++i) {
            mbar_init(&raw_full[i], 1);
            mbar_init(&a_ready[i], 256);
            mbar_init(&a_empty[i], 1);
            mbar_init(&b_full[i], 1);
            mbar_init(&b_empty[i], 1);
        }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc(tmem_ptr, BN);
    __syncthreads();   // barriers initialised: the loader starts before the scale/shift tables exist

    const uint8_t* wsrc = a.w + (size_t)ntile * KG * P::B_STAGE;
    auto load_a = [&](int kg) {
        const int sa = kg % NA_T;
        mbar_arrive_expect_tx(&raw_full[sa], P::A_STAGE);
        tma_tile_3d(sA + sa * P::A_STAGE, &tmA, kg * KC, m0, s, &raw_full[sa]);
    };
    auto load_b = [&](int kg) {
        const int sb = kg % NB_T;
        mbar_arrive_expect_tx(&b_full[sb], P::B_STAGE);
        tma_bulk_load(sB + sb * P::B_STAGE, wsrc + (size_t)kg * P::B_STAGE, P::B_STAGE, &b_full[sb]);
    };
    if (warp == 9 && lane == 0) {
        // fill every stage before joining the CTA-wide barrier below
        for (int kg = 0; kg < NA_T && kg < KG; ++kg) load_a(kg);
        for (int kg = 0; kg < NB_T && kg < KG; ++kg) load_b(kg);
    }

    // BN prologue parameters of this sample (overlaps the first TMA loads)
    if (a.prologue_mode == 0) {
        // mean / variance from the double sums (the cancellation in E[x^2] - E[x]^2 needs double); the reciprocal square
        // root itself in fp32 (rsqrtf + one Newton step, <= 1 ulp): this table is on the critical path of every CTA and
        // the double-precision sqrt + divide sequence it replaces cost about a microsecond of it
        const double inv = 1.0 / ((double)a.hin * a.hin);
        for (int c = tid; c < a.cin; c += 448) {
            const double2 st = *reinterpret_cast<const double2*>(a.in_stats + 2 * ((size_t)s * a.in_stats_stride + c));
            const double m = st.x * inv;
            double var = st.y * inv - m * m;
            if (var < 0) var = 0;
            const float ve = (float)(var + (double)kBnEps);
            float r = rsqrtf(ve);
            r = r * (1.5f - 0.5f * ve * r * r);
            const float sc = a.gamma[c] * r;
            s_sc[c] = sc;
            s_sh[c] = a.beta[c] - (float)m * sc;
        }
    } else {
        for (int c = tid; c < a.cin; c += 448) {
            s_sc[c] = a.scale[(size_t)s * a.cin + c];
            s_sh[c] = a.shift[(size_t)s * a.cin + c];
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 9) {
        // =============================== TMA loader (activations + weights) ===============================
        if (lane == 0) {
            for (int kg = 0; kg < KG; ++kg) {   // the MMAs of stage kg free one activation and one weight slot
                const int ka = kg + NA_T, kb = kg + NB_T;
                if (ka < KG) {
                    mbar_wait_sleep(&a_empty[ka % NA_T], ((ka / NA_T) & 1) ^ 1, 64);
                    load_a(ka);
                }
                if (kb < KG) {
                    mbar_wait_sleep(&b_empty[kb % NB_T], ((kb / NB_T) & 1) ^ 1, 64);
                    load_b(kb);
                }
            }
        }
    } else if (warp < 4 || warp >= 10) {
        // =============================== in-place transform ===============================
        const int ptid = warp < 4 ? tid : tid - 192;          // 0..255
        const int j = ptid & 7;                               // physical 16-byte piece of the 128-byte row
        const int rbase = ptid >> 3;                          // rows rbase + 32 i
        const int chunk = j ^ (rbase & 7);                    // logical 4-channel chunk held by that piece
        int nvalid = hw_out - m0 - rbase;                     // rows with index < nvalid (in steps of 32) exist
        for (int kg = 0; kg < KG; ++kg) {
            const int slot = kg % NA_T;
            const float4 sc = *reinterpret_cast<const float4*>(s_sc + kg * KC + chunk * 4);
            const float4 sh = *reinterpret_cast<const float4*>(s_sh + kg * KC + chunk * 4);
            mbar_wait_sleep(&raw_full[slot], (kg / NA_T) & 1, 64);
            uint8_t* base = sA + slot * P::A_STAGE + rbase * 128 + j * 16;
            float4 x[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) x[i] = *reinterpret_cast<const float4*>(base + i * 32 * 128);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float4 y;
                y.x = fmaf(x[i].x, sc.x, sh.x); y.y = fmaf(x[i].y, sc.y, sh.y);
                y.z = fmaf(x[i].z, sc.z, sh.z); y.w = fmaf(x[i].w, sc.w, sh.w);
                if (a.relu) { y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f); }
                if (i * 32 >= nvalid) y = make_float4(0.f, 0.f, 0.f, 0.f);   // rows beyond the sample contribute nothing
                *reinterpret_cast<float4*>(base + i * 32 * 128) = y;
            }
            fence_proxy_async();
            mbar_arrive(&a_ready[slot]);
        }
    } else if (warp == 8) {
        // =============================== MMA issuer ===============================
        if (lane == 0) {
            constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                       ((uint32_t)(UM >> 4) << 24);
            const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB);
            uint32_t accum = 0;
            for (int kg = 0; kg < KG; ++kg) {
                const int sa = kg % NA_T, sb = kg % NB_T;
                mbar_wait(&a_ready[sa], (kg / NA_T) & 1);
                mbar_wait(&b_full[sb], (kg / NB_T) & 1);
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t ad = make_desc_sw128(sA_u + sa * P::A_STAGE + k * 32);
                    const uint64_t bd = make_desc(sB_u + sb * P::B_STAGE + 2 * k * BN * 16, BN * 16, 128);
                    umma<4>(tmem_base, ad, bd, idesc, accum);
                    accum = 1;
                }
                umma_commit(&a_empty[sa]);
                umma_commit(&b_empty[sb]);
            }
            umma_commit(tmem_full);
        }
    } else if (warp >= 4 && warp < 8) {
        // =============================== epilogue (warps 4-7) ===============================
        const int e = warp - 4;
        const int row = e * 32 + lane;
        mbar_wait_sleep(tmem_full, 0, 128);
        tc_fence_after();
        const bool valid = m0 + row < hw_out;
#pragma unroll
        for (int cb = 0; cb < BN; cb += 32) {
            float v[32];
            tmem_ld32(tmem_base + ((uint32_t)(e * 32) << 16) + cb, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) s_out[row * (BN + 1) + cb + i] = valid ? v[i] : 0.f;
        }
    }
    tc_fence_before();
    __syncthreads();
    {
        constexpr int NW = 14;
        // rows warp, warp + 14, ...: the loads of five rows are issued before their stores so that the shared-memory
        // latency is paid once per group, not once per row
        float* obase = a.out + ((size_t)s * hw_out + m0) * a.out_cstride + a.out_coff + ntile * BN + lane;
        const int nrows = min(UM, hw_out - m0);
        // the statistics ride on the same pass: every lane sums the columns it stores (rows staged as zeros beyond the
        // sample add nothing), the 14 warp partials are combined in a fixed order -> one double atomic pair per channel
        float su[BN / 32], sq[BN / 32];
#pragma unroll
        for (int cb = 0; cb < BN / 32; ++cb) { su[cb] = 0.f; sq[cb] = 0.f; }
#pragma unroll
        for (int i0 = 0; i0 < 10; i0 += 5) {
            float x[5][BN / 32];
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                const int r = warp + NW * (i0 + i);
#pragma unroll
                for (int cb = 0; cb < BN / 32; ++cb) x[i][cb] = r < UM ? s_out[r * (BN + 1) + cb * 32 + lane] : 0.f;
            }
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                const int r = warp + NW * (i0 + i);
#pragma unroll
                for (int cb = 0; cb < BN / 32; ++cb) {
                    su[cb] += x[i][cb];
                    sq[cb] = fmaf(x[i][cb], x[i][cb], sq[cb]);
                }
                if (r < nrows) {
#pragma unroll
                    for (int cb = 0; cb < BN / 32; ++cb) obase[(size_t)r * a.out_cstride + cb * 32] = x[i][cb];
                }
            }
        }
        if (a.out_stats != nullptr) {
            // [2][NW][BN] floats behind the staging tile (the weight stages there are dead by now)
            float* red = reinterpret_cast<float*>(smem + P::OFF_A + ((P::STAGING + 15) & ~15));
            static_assert(P::OFF_A + ((P::STAGING + 15) & ~15) + 2 * NW * BN * 4 <= P::USED, "no room for the warp partials");
#pragma unroll
            for (int cb = 0; cb < BN / 32; ++cb) {
                red[warp * BN + cb * 32 + lane] = su[cb];
                red[(NW + warp) * BN + cb * 32 + lane] = sq[cb];
            }
            __syncthreads();
            if (tid < BN) {
                double dsu = 0.0, dsq = 0.0;
#pragma unroll
                for (int w2 = 0; w2 < NW; ++w2) {
                    dsu += (double)red[w2 * BN + tid];
                    dsq += (double)red[(NW + w2) * BN + tid];
                }
                double* st = a.out_stats + 2 * ((size_t)s * a.out_stats_stride + a.out_coff + ntile * BN + tid);
                atomicAdd(st, dsu);
                atomicAdd(st + 1, dsq);
            }
        }
    }
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, BN);
    }
}

// ------------------------------------------------------------------------------------------
// 3x3 (128 -> 32): the (ht+2) x wp activation patch of one 32-channel group is ONE 4-D TMA box (channels, x, y, sample)
// whose out-of-image halo is zero-filled by the copy engine; it lands as patch rows of 128 B under the 128-byte
// swizzle.  The nine taps stay nine descriptors over the same patch: a shift of s patch rows is +128 s bytes on the
// descriptor start address - the swizzle is a function of the absolute shared-memory address, so the rows the tensor
// core reads are de-swizzled exactly as the copy engine swizzled them.
// ------------------------------------------------------------------------------------------
struct Tma3Plan {
    static constexpr int A_SLOT = 28 * 1024;                // >= 215 rows x 128 B, multiple of the swizzle period
    static constexpr int A_SLOTS = 2;
    static constexpr int B_STAGE = 8 * 32 * 16;
    static constexpr int OFF_SC = 256;
    static constexpr int OFF_A = 9216;
    static constexpr int OFF_B = OFF_A + A_SLOTS * A_SLOT;
    static constexpr int END_AB = OFF_B + NB9 * B_STAGE;
    static constexpr int TOTAL = END_AB + 1024;
};

__global__ void __launch_bounds__(448, 2)
conv3_umma_tma_kernel(const __grid_constant__ CUtensorMap tmA, UmmaDev a) {
    using P = Tma3Plan;
    constexpr int BN = 32;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    uint64_t* raw_full = bars;          // [2]
    uint64_t* a_ready = bars + 2;       // [2]
    uint64_t* a_empty = bars + 4;       // [2]
    uint64_t* b_full = bars + 6;        // [6]
    uint64_t* b_empty = bars + 12;      // [6]
    uint64_t* tmem_full = bars + 18;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 19);
    float* s_sc = reinterpret_cast<float*>(smem + P::OFF_SC);
    float* s_sh = s_sc + 1024;
    uint8_t* sA = smem + P::OFF_A;
    uint8_t* sB = smem + P::OFF_B;
    float* s_out = reinterpret_cast<float*>(smem + P::OFF_A);

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int s = blockIdx.z;
    const int hout = a.hout, hin = a.hin;
    const int hw_out = hout * hout;
    const int wp = a.wp;
    const int ty = blockIdx.x / a.tiles_x, tx = blockIdx.x - ty * a.tiles_x;
    const int h0 = ty * a.ht, w0 = tx * (wp - 2);
    const int pfill = (a.ht + 2) * wp;

    if (warp == 8 && lane == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(&raw_full[i], 1); mbar_init(&a_ready[i], 256); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < NB9; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc(tmem_ptr, BN);
    __syncthreads();

    auto load_a = [&](int g) {
        const int slot = g & 1;
        mbar_arrive_expect_tx(&raw_full[slot], (uint32_t)pfill * 128u);
        tma_tile_4d(sA + slot * P::A_SLOT, &tmA, g * KC, w0 - 1, h0 - 1, s, &raw_full[slot]);
    };
    auto load_b = [&](int j) {
        const int sb = j % NB9;
        mbar_arrive_expect_tx(&b_full[sb], P::B_STAGE);
        tma_bulk_load(sB + sb * P::B_STAGE, a.w + (size_t)j * P::B_STAGE, P::B_STAGE, &b_full[sb]);
    };
    if (warp == 9 && lane == 0) {
        load_a(0);
        load_a(1);
        for (int j = 0; j < NB9; ++j) load_b(j);
    }

    if (a.prologue_mode == 0) {
        const double cnt = (double)hin * hin;
        for (int c = tid; c < a.cin; c += 448) {
            const double* st = a.in_stats + 2 * ((size_t)s * a.in_stats_stride + c);
            const double m = st[0] / cnt;
            double var = st[1] / cnt - m * m;
            if (var < 0) var = 0;
            const float sc = a.gamma[c] * (float)(1.0 / sqrt(var + (double)kBnEps));
            s_sc[c] = sc;
            s_sh[c] = a.beta[c] - (float)m * sc;
        }
    } else {
        for (int c = tid; c < a.cin; c += 448) {
            s_sc[c] = a.scale[(size_t)s * a.cin + c];
            s_sh[c] = a.shift[(size_t)s * a.cin + c];
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 9) {
        if (lane == 0) {
            for (int j = 0; j < 36; ++j) {   // the MMAs of stage j free weight slot j % NB9; tap 8 of a group frees its patch slot
                const int kb = j + NB9;
                if (kb < 36) {
                    mbar_wait_sleep(&b_empty[kb % NB9], ((kb / NB9) & 1) ^ 1, 64);
                    load_b(kb);
                }
                const int g = j / 9;
                if (j - g * 9 == 8 && g + 2 < 4) {
                    mbar_wait_sleep(&a_empty[g & 1], 0, 64);
                    load_a(g + 2);
                }
            }
        }
    } else if (warp < 4 || warp >= 10) {
        // in-place transform: thread owns 16-byte piece j of patch rows rbase + 32 i
        const int ptid = warp < 4 ? tid : tid - 192;
        const int j = ptid & 7;
        const int rbase = ptid >> 3;
        const int chunk = j ^ (rbase & 7);
        constexpr int NI = 7;                    // 7 x 32 = 224 >= 5 * MAX_WP + ... patch rows
        uint32_t inside = 0, filled = 0;         // bit i: pixel inside the image / row part of the patch
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            const int q = rbase + 32 * i;
            const int py = q / wp, px = q - py * wp;
            const int y = h0 - 1 + py, x = w0 - 1 + px;
            if (q < pfill) {
                filled |= 1u << i;
                if (y >= 0 && y < hin && x >= 0 && x < hin) inside |= 1u << i;
            }
        }
        for (int g = 0; g < 4; ++g) {
            const int slot = g & 1;
            const float4 sc = *reinterpret_cast<const float4*>(s_sc + g * KC + chunk * 4);
            const float4 sh = *reinterpret_cast<const float4*>(s_sh + g * KC + chunk * 4);
            mbar_wait_sleep(&raw_full[slot], (g >> 1) & 1, 64);
            uint8_t* base = sA + slot * P::A_SLOT + rbase * 128 + j * 16;
            float4 x[NI];
#pragma unroll
            for (int i = 0; i < NI; ++i)
                if (filled & (1u << i)) x[i] = *reinterpret_cast<const float4*>(base + i * 32 * 128);
#pragma unroll
            for (int i = 0; i < NI; ++i) {
                if (!(filled & (1u << i))) continue;
                float4 y;
                y.x = fmaf(x[i].x, sc.x, sh.x); y.y = fmaf(x[i].y, sc.y, sh.y);
                y.z = fmaf(x[i].z, sc.z, sh.z); y.w = fmaf(x[i].w, sc.w, sh.w);
                if (a.relu) { y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f); }
                if (!(inside & (1u << i))) y = make_float4(0.f, 0.f, 0.f, 0.f);   // conv zero padding is post-activation
                *reinterpret_cast<float4*>(base + i * 32 * 128) = y;
            }
            fence_proxy_async();
            mbar_arrive(&a_ready[slot]);
        }
    } else if (warp == 8) {
        if (lane == 0) {
            constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                       ((uint32_t)(UM >> 4) << 24);
            const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB);
            uint32_t accum = 0;
            for (int g = 0; g < 4; ++g) {
                const int sa = g & 1;
                mbar_wait(&a_ready[sa], (g >> 1) & 1);
                for (int t = 0; t < 9; ++t) {
                    const int jj = g * 9 + t;
                    const int sb = jj % NB9;
                    mbar_wait(&b_full[sb], (jj / NB9) & 1);
                    tc_fence_after();
                    const int shift = (t / 3) * wp + (t % 3);
                    const uint32_t start = sA_u + sa * P::A_SLOT + shift * 128;
                    // descriptor base_offset stays 0: measured on B200, the swizzle phase follows the absolute address (setting
                    // base_offset = (start >> 7) & 7 for the shifted start gives wrong results)
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t ad = make_desc_sw128(start + k * 32);
                        const uint64_t bd = make_desc(sB_u + sb * P::B_STAGE + 2 * k * BN * 16, BN * 16, 128);
                        umma<4>(tmem_base, ad, bd, idesc, accum);
                        accum = 1;
                    }
                    umma_commit(&b_empty[sb]);
                }
                umma_commit(&a_empty[sa]);
            }
            umma_commit(tmem_full);
        }
    } else if (warp >= 4 && warp < 8) {
        const int e = warp - 4;
        const int row = e * 32 + lane;
        mbar_wait_sleep(tmem_full, 0, 128);
        tc_fence_after();
        const int i = row / wp, jx = row - i * wp;
        const bool valid = i < a.ht && jx < wp - 2 && h0 + i < hout && w0 + jx < hout;
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(e * 32) << 16), v);
#pragma unroll
        for (int c = 0; c < 32; ++c) s_out[row * (BN + 1) + c] = valid ? v[c] : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    {
        constexpr int NW = 14;
        for (int r = warp; r < UM; r += NW) {
            const int i = r / wp, jx = r - i * wp;
            if (!(i < a.ht && jx < wp - 2 && h0 + i < hout && w0 + jx < hout)) continue;
            float* o = a.out + ((size_t)s * hw_out + (h0 + i) * hout + w0 + jx) * a.out_cstride + a.out_coff;
            o[lane] = s_out[r * (BN + 1) + lane];
        }
        if (a.out_stats != nullptr) {
            constexpr int GROUPS = 448 / BN;                        // 14 row groups
            constexpr int RPG = (UM + GROUPS - 1) / GROUPS;
            float* red = s_sc;
            const int cidx = tid % BN, g = tid / BN;
            float su = 0.f, sq = 0.f;
            const int r1 = (g + 1) * RPG < UM ? (g + 1) * RPG : UM;
            for (int r = g * RPG; r < r1; ++r) {
                const float x = s_out[r * (BN + 1) + cidx];
                su += x;
                sq = fmaf(x, x, sq);
            }
            red[g * BN + cidx] = su;
            red[(GROUPS + g) * BN + cidx] = sq;
            __syncthreads();
            if (tid < BN) {
                double dsu = 0.0, dsq = 0.0;
#pragma unroll
                for (int g2 = 0; g2 < GROUPS; ++g2) {
                    dsu += (double)red[g2 * BN + tid];
                    dsq += (double)red[(GROUPS + g2) * BN + tid];
                }
                double* st = a.out_stats + 2 * ((size_t)s * a.out_stats_stride + a.out_coff + tid);
                atomicAdd(st, dsu);
                atomicAdd(st + 1, dsq);
            }
        }
    }
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, BN);
    }
}

}  // namespace

EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int make_tensor_map_f32(CUtensorMap* tm, const float* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                        const cuuint32_t* box, int swizzle_bytes) {
    EncodeTiledFn enc = encode_tiled_fn();
    SMG_CHECK(enc != nullptr, SMG_ERR_CUDA, "tensor map: cuTensorMapEncodeTiled not available from the driver");
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<float*>(base), dims, strides, box,
                           estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SMG_CHECK(r == CUDA_SUCCESS, SMG_ERR_CUDA, "tensor map: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return SMG_OK;
}

// Returns SMG_ERR_UNSUPPORTED for shapes this kernel does not serve (the caller then uses conv_umma.cu).
int launch_conv_umma_tma(smg_handle* h, const ConvArgs& a, cudaStream_t st) {
    if (a.taps != 1 || a.pool || a.cout % 128 != 0 || a.cin % KC != 0 || a.cin > 1024 || a.in_cstride % 4 != 0 ||
        (reinterpret_cast<uintptr_t>(a.in) & 15) != 0)
        return SMG_ERR_UNSUPPORTED;
    SMG_CHECK(a.w != nullptr && a.w->w_tf32 != nullptr, SMG_ERR_STATE, "conv_umma_tma: weights not packed");
    const int hw = a.hin * a.hin;
    CUtensorMap tm;
    const cuuint64_t dims[3] = {(cuuint64_t)a.in_cstride, (cuuint64_t)hw, (cuuint64_t)a.n};
    const cuuint64_t strides[2] = {(cuuint64_t)a.in_cstride * 4, (cuuint64_t)hw * a.in_cstride * 4};
    const cuuint32_t box[3] = {KC, UM, 1};
    SMG_TRY(make_tensor_map_f32(&tm, a.in, 3, dims, strides, box));

    UmmaDev d;
    d.in = a.in; d.in_cstride = a.in_cstride; d.cin = a.cin; d.hin = a.hin;
    d.prologue_mode = a.prologue_mode; d.in_stats = a.in_stats; d.in_stats_stride = a.in_stats_stride;
    d.gamma = a.gamma; d.beta = a.beta; d.scale = a.scale; d.shift = a.shift; d.relu = a.relu;
    d.w = a.w->w_tf32;
    d.out = a.out; d.out_cstride = a.out_cstride; d.out_coff = a.out_coff; d.cout = a.cout;
    d.out_stats = a.out_stats; d.out_stats_stride = a.out_stats_stride;
    d.hout = a.hin;
    d.wp = d.ht = d.tiles_x = 0;
    d.async_producer = 0;
    d.tiles_per_sample = d.tiles_per_cta = 0;
    using P = TmaPlan<128>;
    static bool attr = false;
    if (!attr) {
        SMG_CUDA(cudaFuncSetAttribute(conv_umma_tma_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, P::TOTAL));
        attr = true;
    }
    const dim3 grid((hw + UM - 1) / UM, a.cout / 128, a.n);
    conv_umma_tma_kernel<128><<<grid, 448, P::TOTAL, st>>>(tm, d);
    h->launches++;
    SMG_CUDA(cudaGetLastError());
    return SMG_OK;
}

int launch_conv3_umma_tma(smg_handle* h, const ConvArgs& a, cudaStream_t st) {
    if (a.taps != 9 || a.pool || a.cin != 128 || a.cout != 32 || a.in_cstride % 4 != 0 ||
        (reinterpret_cast<uintptr_t>(a.in) & 15) != 0)
        return SMG_ERR_UNSUPPORTED;
    SMG_CHECK(a.w != nullptr && a.w->w_tf32 != nullptr, SMG_ERR_STATE, "conv3_umma_tma: weights not packed");
    UmmaDev d;
    d.in = a.in; d.in_cstride = a.in_cstride; d.cin = a.cin; d.hin = a.hin;
    d.prologue_mode = a.prologue_mode; d.in_stats = a.in_stats; d.in_stats_stride = a.in_stats_stride;
    d.gamma = a.gamma; d.beta = a.beta; d.scale = a.scale; d.shift = a.shift; d.relu = a.relu;
    d.w = a.w->w_tf32;
    d.out = a.out; d.out_cstride = a.out_cstride; d.out_coff = a.out_coff; d.cout = a.cout;
    d.out_stats = a.out_stats; d.out_stats_stride = a.out_stats_stride;
    d.hout = a.hin;
    umma_patch_geometry(d.hout, &d.wp, &d.ht);
    SMG_CHECK((d.ht + 2) * d.wp * 128 <= Tma3Plan::A_SLOT && d.ht * d.wp <= UM, SMG_ERR_STATE,
              "conv3_umma_tma: patch %dx%d too large", d.ht, d.wp);
    const int wt = d.wp - 2;
    d.tiles_x = (d.hout + wt - 1) / wt;
    const int tiles_y = (d.hout + d.ht - 1) / d.ht;
    d.async_producer = 0;
    d.tiles_per_sample = d.tiles_per_cta = 0;

    CUtensorMap tm;
    const cuuint64_t dims[4] = {(cuuint64_t)a.in_cstride, (cuuint64_t)a.hin, (cuuint64_t)a.hin, (cuuint64_t)a.n};
    const cuuint64_t strides[3] = {(cuuint64_t)a.in_cstride * 4, (cuuint64_t)a.hin * a.in_cstride * 4,
                                   (cuuint64_t)a.hin * a.hin * a.in_cstride * 4};
    const cuuint32_t box[4] = {KC, (cuuint32_t)d.wp, (cuuint32_t)(d.ht + 2), 1};
    SMG_TRY(make_tensor_map_f32(&tm, a.in, 4, dims, strides, box));
    static bool attr = false;
    if (!attr) {
        SMG_CUDA(cudaFuncSetAttribute(conv3_umma_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Tma3Plan::TOTAL));
        attr = true;
    }
    const dim3 grid(d.tiles_x * tiles_y, 1, a.n);
    conv3_umma_tma_kernel<<<grid, 448, Tma3Plan::TOTAL, st>>>(tm, d);
    h->launches++;
    SMG_CUDA(cudaGetLastError());
    return SMG_OK;
}

}  // namespace smg
