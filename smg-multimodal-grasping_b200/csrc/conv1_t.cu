// conv1_t.cu - persistent tf32 1x1 convolution (cin -> 128) with the operand roles swapped: weights are the A operand.
//
// Serves torchvision densenet `_DenseLayer.conv1` (/root/reference/code/models.py:319 builds the trunks); same contract as
// conv_umma_tma.cu (BN-ReLU prologue, raw NHWC output slice, (sum, sumsq) epilogue).
//
//   D[cout][pixel m] += W[cout][cin] * X[m][cin]
//
//   A operand = the weight matrix [128 cout][cin] (K-major = torch's own [cout][cin] layout).  cin <= 256: resident in
//               TENSOR memory (cin columns, written once per launch with tcgen05.st) - the layer's weights are then never
//               streamed again, which halves the bytes the copy engine has to deliver per activation stage (the 1x1 layers
//               are bound by that delivery rate, profiles/README.md); cin > 256: streamed through shared-memory stages
//               (pack.cu's stage images are valid K-major A tiles as they are);
//   B operand = the activation stage in shared memory (128 pixels x 32 channels, 128-byte swizzle), fetched by tensor-map
//               TMA and normalised in place by the transform warps;
//   D         = [cout lanes][128 pixel columns], double buffered (2 x 128 columns of tensor memory).
// With lanes = output channels the epilogue needs neither staging nor shuffles: a warp owns 32 channels, every
// instruction stores one pixel's 128 contiguous bytes, the statistics are per-thread sums.
// Warps (576 threads): 0-3 / 8-11 transform (two groups, tied to the stage parity), 4-7 epilogue of even tiles,
// 12-15 epilogue of odd tiles (quadrant = warp mod 4), 16 MMA issuer, 17 TMA loader.
#include "tma_common.cuh"

namespace smg {

namespace {

constexpr int T_STAGE = UM * 128;                  // 128 rows x 128 B
constexpr int T_RES_MAXK = 256;                    // resident mode: cin <= 256 (tensor memory: 256 weight + 2 x 128 accumulator columns)
constexpr int T_THREADS = 576;

template <bool RES>
struct T1 {
    static constexpr int NA = RES ? 12 : 8;               // activation stages (even: see the transform groups)
    static constexpr int NB = RES ? 0 : 4;                // weight stages
    static constexpr int OFF_A = 0;
    static constexpr int OFF_B = OFF_A + NA * T_STAGE;
    static constexpr int OFF_SC = OFF_B + NB * T_STAGE;   // scale[1024], shift[1024]
    static constexpr int OFF_BAR = OFF_SC + 8192;
    static constexpr int TOTAL = OFF_BAR + 512;
    static_assert(TOTAL <= 232448, "shared-memory plan exceeds the 227 KB of one SM");
};

struct Tile1 {
    int s, m0;
};

template <bool RES>
__global__ void __launch_bounds__(T_THREADS, 1)
conv1_t_kernel(const __grid_constant__ CUtensorMap tmA, UmmaDev a, const float* __restrict__ w_t, int total_tiles) {
    using Q = T1<RES>;
    constexpr int NA = Q::NA, NB = RES ? 1 : Q::NB;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Q::OFF_BAR);
    uint64_t* raw_full = bars;          // [12] activations landed (raw)
    uint64_t* a_ready = bars + 12;      // [12] normalised (the 128 transform threads that own the stage)
    uint64_t* a_empty = bars + 24;      // [12] MMAs retired
    uint64_t* b_full = bars + 36;       // [4] weight stage landed (streamed mode)
    uint64_t* b_empty = bars + 40;      // [4]
    uint64_t* t_full = bars + 44;       // [2] accumulator complete
    uint64_t* t_empty = bars + 46;      // [2] accumulator drained (128 epilogue threads)
    uint64_t* w_ready = bars + 48;      // weights in tensor memory (256 epilogue threads; resident mode)
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 49);
    float* s_sc = reinterpret_cast<float*>(smem + Q::OFF_SC);
    float* s_sh = s_sc + 1024;
    uint8_t* sA = smem + Q::OFF_A;
    uint8_t* sB = smem + Q::OFF_B;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int hw_out = a.hout * a.hout;
    const int KG = a.cin / KC;
    const int tps = a.tiles_per_sample;
    const int tile_begin = (int)(((long long)blockIdx.x * total_tiles) / gridDim.x);
    const int tile_end = (int)(((long long)(blockIdx.x + 1) * total_tiles) / gridDim.x);
    const int ntiles = tile_end - tile_begin;
    auto coord = [&](int tile) {
        Tile1 c;
        c.s = tile / tps;
        c.m0 = (tile - c.s * tps) * UM;
        return c;
    };
    const bool is_transform = warp < 16 && (warp & 7) < 4;    // warps 0-3, 8-11
    const bool is_epilogue = warp < 16 && (warp & 7) >= 4;    // warps 4-7, 12-15

    if (warp == 16 && lane == 0) {
        if (smem_u32(smem) & 1023u) __trap();   // the swizzled stages rely on a 1024-byte aligned window
        for (int i = 0; i < NA; ++i) { mbar_init(&raw_full[i], 1); mbar_init(&a_ready[i], 128); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < 4; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 128); }
        mbar_init(w_ready, 256);
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc(tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;          // columns [0,256) accumulators, [256, 256 + cin) weights
    const uint32_t tmem_w = tmem_base + 256;
    const int total_stages = ntiles * KG;
    // Programmatic dependent launch: barriers, tensor memory and (below) the weight fill touch nothing an earlier kernel of
    // the pass writes, so this CTA may be resident while the previous layer drains.  Everything that reads the previous
    // kernel's output (statistics, activations) sits behind pdl_wait(); the epilogue's stores follow from the loads.
    pdl_launch_dependents();

    if (warp == 17) {
        // =============================== loader ===============================
        if (lane == 0) {
            pdl_wait();
            int qa = 0, ta = 0, ka = 0;      // next activation stage: global index, tile, channel group
            int qb = RES ? total_stages : 0, kb = 0;
            Tile1 ca = coord(tile_begin);
            while (qa < total_stages || qb < total_stages) {
                if (qa < total_stages && mbar_test(&a_empty[qa % NA], ((qa / NA) & 1) ^ 1)) {
                    const int slot = qa % NA;
                    mbar_arrive_expect_tx(&raw_full[slot], T_STAGE);
                    tma_tile_3d(sA + slot * T_STAGE, &tmA, ka * KC, ca.m0, ca.s, &raw_full[slot]);
                    ++qa;
                    if (++ka == KG) {
                        ka = 0;
                        if (++ta < ntiles) ca = coord(tile_begin + ta);
                    }
                }
                if (!RES && qb < total_stages && mbar_test(&b_empty[qb % NB], ((qb / NB) & 1) ^ 1)) {
                    const int slot = qb % NB;
                    mbar_arrive_expect_tx(&b_full[slot], T_STAGE);
                    tma_bulk_load(sB + slot * T_STAGE, a.w + (size_t)kb * T_STAGE, T_STAGE, &b_full[slot]);
                    ++qb;
                    if (++kb == KG) kb = 0;
                }
            }
        }
    } else if (is_transform) {
        // =============================== in-place transform ===============================
        // two groups of four warps; group g owns the stages with global index = g (mod 2).  NA is even: a slot is always
        // served by the same group, which keeps the one-bit barrier parity unambiguous.
        const int ptid = warp < 4 ? tid : tid - 128;          // 0..255
        const int grp = ptid >> 7;
        const int gt = ptid & 127;
        const int j = gt & 7;                                 // physical 16-byte piece of the 128-byte row
        const int rbase = gt >> 3;                            // rows rbase + 16 i, i < 8
        const int chunk = j ^ (rbase & 7);                    // logical 4-channel chunk held by that piece
        int cur_s = -1;
        int q0 = 0;                                           // global index of the tile's first stage
        pdl_wait();                                           // the tables below read the producers' statistics
        for (int it = 0; it < ntiles; ++it, q0 += KG) {
            const Tile1 c = coord(tile_begin + it);
            if (c.s != cur_s) {
                asm volatile("bar.sync 2, 256;" ::: "memory");
                const double inv = 1.0 / ((double)a.hin * a.hin);
                for (int ch = ptid; ch < a.cin; ch += 256) {
                    float sc, sh;
                    if (a.prologue_mode == 0) {
                        const double2 st = *reinterpret_cast<const double2*>(a.in_stats + 2 * ((size_t)c.s * a.in_stats_stride + ch));
                        const double m = st.x * inv;
                        double var = st.y * inv - m * m;
                        if (var < 0) var = 0;
                        const float ve = (float)(var + (double)kBnEps);
                        float r = rsqrtf(ve);
                        r = r * (1.5f - 0.5f * ve * r * r);
                        sc = a.gamma[ch] * r;
                        sh = a.beta[ch] - (float)m * sc;
                    } else {
                        sc = a.scale[(size_t)c.s * a.cin + ch];
                        sh = a.shift[(size_t)c.s * a.cin + ch];
                    }
                    s_sc[ch] = sc;
                    s_sh[ch] = sh;
                }
                asm volatile("bar.sync 2, 256;" ::: "memory");
                cur_s = c.s;
            }
            const int nvalid = hw_out - c.m0 - rbase;         // row rbase + 16 i exists iff 16 i < nvalid
            for (int kg = (grp - q0) & 1; kg < KG; kg += 2) {
                const int q = q0 + kg;
                const int slot = q % NA;
                const float4 sc = *reinterpret_cast<const float4*>(s_sc + kg * KC + chunk * 4);
                const float4 sh = *reinterpret_cast<const float4*>(s_sh + kg * KC + chunk * 4);
                mbar_wait_sleep(&raw_full[slot], (q / NA) & 1, 64);
                uint8_t* base = sA + slot * T_STAGE + rbase * 128 + j * 16;
#pragma unroll
                for (int i0 = 0; i0 < 8; i0 += 4) {
                    float4 x[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) x[i] = *reinterpret_cast<const float4*>(base + (i0 + i) * 16 * 128);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float4 y;
                        y.x = fmaf(x[i].x, sc.x, sh.x); y.y = fmaf(x[i].y, sc.y, sh.y);
                        y.z = fmaf(x[i].z, sc.z, sh.z); y.w = fmaf(x[i].w, sc.w, sh.w);
                        if (a.relu) { y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f); }
                        if ((i0 + i) * 16 >= nvalid) y = make_float4(0.f, 0.f, 0.f, 0.f);   // pixels beyond the sample contribute nothing
                        *reinterpret_cast<float4*>(base + (i0 + i) * 16 * 128) = y;
                    }
                }
                fence_proxy_async();
                mbar_arrive(&a_ready[slot]);
            }
        }
    } else if (warp == 16) {
        // =============================== MMA issuer ===============================
        if (lane == 0) {
            // D[128 cout x 128 px] += A[128 x 8] (weights) * B[128 x 8]^T (activations, shared memory)
            constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB);
            int q = 0;
            if (RES) {
                mbar_wait(w_ready, 0);
                tc_fence_after();
            }
            for (int it = 0; it < ntiles; ++it) {
                const int buf = it & 1;
                mbar_wait(&t_empty[buf], ((it >> 1) & 1) ^ 1);   // the epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(buf * 128);
                uint32_t accum = 0;
                for (int kg = 0; kg < KG; ++kg, ++q) {
                    const int sa = q % NA, sb = RES ? 0 : q % NB;
                    mbar_wait(&a_ready[sa], (q / NA) & 1);
                    if (!RES) mbar_wait(&b_full[sb], (q / NB) & 1);
                    tc_fence_after();
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t xd = make_desc_sw128(sA_u + sa * T_STAGE + k * 32);   // activations: the B operand
                        if (RES) {
                            umma_ts_tf32(d_tmem, tmem_w + (uint32_t)(kg * 32 + k * 8), xd, idesc, accum);
                        } else {
                            const uint64_t wd = make_desc(sB_u + sb * T_STAGE + 2 * k * 128 * 16, 128 * 16, 128);   // weights: A
                            umma<4>(d_tmem, wd, xd, idesc, accum);
                        }
                        accum = 1;
                    }
                    umma_commit(&a_empty[sa]);
                    if (!RES) umma_commit(&b_empty[sb]);
                }
                umma_commit(&t_full[buf]);
            }
        }
    } else if (is_epilogue) {
        // =============================== epilogue (warps 4-7: even tiles, 12-15: odd tiles) ===============================
        const int eg = warp >> 3;            // epilogue group == accumulator buffer it drains
        const int q4 = warp & 3;             // TMEM lane quadrant: output channels 32 q4 + lane
        if (RES) {
            // this quadrant's 32 weight rows into tensor memory; the two warps of a quadrant split the column blocks.
            // Global layout [column block of 16][128 rows][16 floats]: a warp reads 2 KB contiguous per block.
            const int nblk = a.cin / 16;
            const float4* wsrc = reinterpret_cast<const float4*>(w_t) + (size_t)(q4 * 32 + lane) * 4;
            for (int c16 = eg; c16 < nblk; c16 += 2) {
                const float4* p4 = wsrc + (size_t)c16 * 128 * 4;
                const float4 w0 = __ldg(p4), w1 = __ldg(p4 + 1), w2 = __ldg(p4 + 2), w3 = __ldg(p4 + 3);
                tmem_st16(tmem_w + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(c16 * 16), w0, w1, w2, w3);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            mbar_arrive(w_ready);
        }
        double acc_su = 0.0, acc_ss = 0.0;   // statistics of channel 32 q4 + lane over the tiles this warp drained
        int cur_s = -1;
        auto flush = [&](int s_done) {
            if (a.out_stats == nullptr) return;
            double* st = a.out_stats + 2 * ((size_t)s_done * a.out_stats_stride + a.out_coff + q4 * 32 + lane);
            atomicAdd(st, acc_su);
            atomicAdd(st + 1, acc_ss);
            acc_su = 0.0;
            acc_ss = 0.0;
        };
        for (int it = eg; it < ntiles; it += 2) {
            const Tile1 c = coord(tile_begin + it);
            if (c.s != cur_s) {
                if (cur_s >= 0) flush(cur_s);
                cur_s = c.s;
            }
            const int rows = min(UM, hw_out - c.m0);          // pixels of this tile that exist
            float* obase = a.out + ((size_t)c.s * hw_out + c.m0) * a.out_cstride + a.out_coff + q4 * 32 + lane;
            mbar_wait_sleep(&t_full[eg], (it >> 1) & 1, 64);
            tc_fence_after();
            float su = 0.f, sq = 0.f;
#pragma unroll
            for (int cb = 0; cb < 4; ++cb) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(eg * 128 + cb * 32), v);
                if (cb == 3) {
                    tc_fence_before();
                    mbar_arrive(&t_empty[eg]);   // the accumulator is in registers: the MMA warp may overwrite it
                }
                if (rows == UM) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        obase[(size_t)(cb * 32 + i) * a.out_cstride] = v[i];   // 32 lanes = 128 contiguous bytes of one pixel
                        su += v[i];
                        sq = fmaf(v[i], v[i], sq);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        if (cb * 32 + i < rows) {
                            obase[(size_t)(cb * 32 + i) * a.out_cstride] = v[i];
                            su += v[i];
                            sq = fmaf(v[i], v[i], sq);
                        }
                    }
                }
            }
            acc_su += (double)su;
            acc_ss += (double)sq;
        }
        if (cur_s >= 0) flush(cur_s);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace

// Returns SMG_ERR_UNSUPPORTED for shapes this kernel does not serve (the caller then uses the other 1x1 kernels).
int launch_conv1_t(smg_handle* h, const ConvArgs& a, cudaStream_t st) {
    if (a.taps != 1 || a.pool || a.cout != 128 || a.cin % KC != 0 || a.cin > 1024 || a.in_cstride % 4 != 0 || a.w == nullptr ||
        a.w->w_tf32 == nullptr || a.w->w_tf32_t == nullptr || (reinterpret_cast<uintptr_t>(a.in) & 15) != 0)
        return SMG_ERR_UNSUPPORTED;
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
    d.tiles_per_sample = (hw + UM - 1) / UM;
    d.tiles_per_cta = 0;
    const int total = d.tiles_per_sample * a.n;
    SMG_TRY(ensure_dyn_smem(h, (const void*)conv1_t_kernel<true>, T1<true>::TOTAL));
    SMG_TRY(ensure_dyn_smem(h, (const void*)conv1_t_kernel<false>, T1<false>::TOTAL));
    const int grid = total < h->num_sms ? total : h->num_sms;
    const float* w_t = reinterpret_cast<const float*>(a.w->w_tf32_t);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(T_THREADS);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = h->use_pdl ? 1 : 0;
    if (a.cin <= T_RES_MAXK) {
        cfg.dynamicSmemBytes = T1<true>::TOTAL;
        SMG_CUDA(cudaLaunchKernelEx(&cfg, conv1_t_kernel<true>, tm, d, w_t, total));
    } else {
        cfg.dynamicSmemBytes = T1<false>::TOTAL;
        SMG_CUDA(cudaLaunchKernelEx(&cfg, conv1_t_kernel<false>, tm, d, w_t, total));
    }
    h->launches++;
    SMG_CUDA(cudaGetLastError());
    return SMG_OK;
}

}  // namespace smg
