// conv_umma_ts.cu - the 1x1 tf32 TMA convolution (conv_umma_tma.cu) with the activation operand in TENSOR MEMORY.
//
// Same contract and results class as conv_umma_tma_kernel<128>.  The 1x1 layers sit on the SM's shared-memory bandwidth
// (DESIGN.md section 4): per 16 KB activation stage the in-place transform writes 16 KB back to shared memory and the MMA
// reads them again.  Here the transform warps read the raw stage from shared memory, apply relu(x*scale+shift) in
// registers and store the operand with tcgen05.st into a 4 x 32-column ring of tensor memory; the MMA takes A from
// there (`tcgen05.mma [d], [a_tmem], b_desc, ...`), so those 32 KB per stage never touch shared memory.
//   tensor memory: columns [0,128) accumulator, [128,256) operand ring (row = TMEM lane, one 32-bit column per channel);
//   a transform warp may only write the 32 lanes of its quadrant (warp index mod 4): quadrant q is served by warp q
//   (channels 0-15 of the stage) and by the warp of 10-13 with the same index mod 4 (channels 16-31);
//   raw_empty (256 arrivals) frees the shared-memory stage for the next TMA as soon as it has been read,
//   a_empty (tcgen05.commit) frees the tensor-memory slot.
#include "tma_common.cuh"

namespace smg {

namespace {

constexpr int NA_T = 4;   // activation stages in flight (16 KB each) = operand slots in tensor memory
constexpr int NB_T = 2;   // weight stages (16 KB each at N = 128)

__device__ __forceinline__ void tmem_st16_wait(uint32_t taddr, const float* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
        "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
        "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

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
conv_umma_ts_kernel(const __grid_constant__ CUtensorMap tmA, UmmaDev a) {
    using P = TmaPlan<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + P::OFF_BAR);
    uint64_t* raw_full = bars;          // [4] TMA transaction barriers (activations landed, still raw)
    uint64_t* a_ready = bars + 4;       // [4] operand slot in tensor memory written (256 transform threads)
    uint64_t* a_empty = bars + 8;       // [4] MMA commit: operand slot free
    uint64_t* b_full = bars + 12;       // [4]
    uint64_t* b_empty = bars + 16;      // [4]
    uint64_t* tmem_full = bars + 20;
    uint64_t* raw_empty = bars + 21;    // [4] shared-memory stage read by the 256 transform threads
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 25);
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
            mbar_init(&raw_empty[i], 256);
        }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc(tmem_ptr, 2 * BN);   // accumulator + 4 x 32-column operand ring
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
                    mbar_wait_sleep(&raw_empty[ka % NA_T], ((ka / NA_T) & 1) ^ 1, 64);
                    load_a(ka);
                }
                if (kb < KG) {
                    mbar_wait_sleep(&b_empty[kb % NB_T], ((kb / NB_T) & 1) ^ 1, 64);
                    load_b(kb);
                }
            }
        }
    } else if (warp < 4 || warp >= 10) {
        // =============================== transform: shared memory -> registers -> tensor memory ===============================
        const int q = warp & 3;                               // TMEM lane quadrant this warp may write
        const int half = warp < 4 ? 0 : 1;                    // channels 16 half .. + 15 of the stage
        const int row = q * 32 + lane;
        const bool exists = m0 + row < hw_out;
        for (int kg = 0; kg < KG; ++kg) {
            const int slot = kg % NA_T;
            float4 sc[4], sh[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                sc[i] = *reinterpret_cast<const float4*>(s_sc + kg * KC + half * 16 + i * 4);
                sh[i] = *reinterpret_cast<const float4*>(s_sh + kg * KC + half * 16 + i * 4);
            }
            mbar_wait_sleep(&raw_full[slot], (kg / NA_T) & 1, 64);
            const uint8_t* base = sA + slot * P::A_STAGE + row * 128;
            float4 x[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) x[i] = *reinterpret_cast<const float4*>(base + (((half * 4 + i) ^ (row & 7)) * 16));
            float y[16];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                y[4 * i + 0] = fmaf(x[i].x, sc[i].x, sh[i].x); y[4 * i + 1] = fmaf(x[i].y, sc[i].y, sh[i].y);
                y[4 * i + 2] = fmaf(x[i].z, sc[i].z, sh[i].z); y[4 * i + 3] = fmaf(x[i].w, sc[i].w, sh[i].w);
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                if (a.relu) y[i] = fmaxf(y[i], 0.f);
                if (!exists) y[i] = 0.f;                      // rows beyond the sample contribute nothing
            }
            mbar_arrive(&raw_empty[slot]);                    // the raw stage is in registers: the next TMA may overwrite it
            mbar_wait_sleep(&a_empty[slot], ((kg / NA_T) & 1) ^ 1, 32);   // the MMAs that read this operand slot have retired
            tc_fence_after();
            tmem_st16_wait(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(BN + slot * 32 + half * 16), y);
            tc_fence_before();
            mbar_arrive(&a_ready[slot]);
        }
    } else if (warp == 8) {
        // =============================== MMA issuer ===============================
        if (lane == 0) {
            constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                       ((uint32_t)(UM >> 4) << 24);
            const uint32_t sB_u = smem_u32(sB);
            uint32_t accum = 0;
            for (int kg = 0; kg < KG; ++kg) {
                const int sa = kg % NA_T, sb = kg % NB_T;
                mbar_wait(&a_ready[sa], (kg / NA_T) & 1);
                mbar_wait(&b_full[sb], (kg / NB_T) & 1);
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t bd = make_desc(sB_u + sb * P::B_STAGE + 2 * k * BN * 16, BN * 16, 128);
                    umma_ts_tf32(tmem_base, tmem_base + (uint32_t)(BN + sa * 32 + k * 8), bd, idesc, accum);
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
        tmem_dealloc(tmem_base, 2 * BN);
    }
}

}  // namespace

// Returns SMG_ERR_UNSUPPORTED for shapes this kernel does not serve (the caller then uses conv_umma_tma.cu).
int launch_conv_umma_ts(smg_handle* h, const ConvArgs& a, cudaStream_t st) {
    if (a.taps != 1 || a.pool || a.cout % 128 != 0 || a.cin % KC != 0 || a.cin > 1024 || a.in_cstride % 4 != 0 ||
        (reinterpret_cast<uintptr_t>(a.in) & 15) != 0)
        return SMG_ERR_UNSUPPORTED;
    SMG_CHECK(a.w != nullptr && a.w->w_tf32 != nullptr, SMG_ERR_STATE, "conv_umma_ts: weights not packed");
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
        SMG_CUDA(cudaFuncSetAttribute(conv_umma_ts_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, P::TOTAL));
        attr = true;
    }
    const dim3 grid((hw + UM - 1) / UM, a.cout / 128, a.n);
    conv_umma_ts_kernel<128><<<grid, 448, P::TOTAL, st>>>(tm, d);
    h->launches++;
    SMG_CUDA(cudaGetLastError());
    return SMG_OK;
}

}  // namespace smg
