// conv1_persist.cu - persistent tf32 1x1 convolution (cin -> 128-wide output tiles) for the dense layers' bottleneck conv.
//
// Serves torchvision densenet `_DenseLayer.conv1` (58 of the 120 trunk convolutions; /root/reference/code/models.py:319
// builds the trunks) with the contract of conv_umma.cu: BN-ReLU prologue, raw NHWC output slice, (sum, sumsq) epilogue.
//
// These layers are HBM-bound (every activation byte is read once, arithmetic intensity 40 flop/B at tf32), so the job
// is to keep loads in flight all the time.  With one tile per CTA (conv_umma_tma.cu) a third of a CTA's life is
// prologue (barriers, TMEM allocation, scale/shift table) and epilogue with nothing in flight.  Here ONE CTA per SM
// walks a contiguous range of the flattened (n-tile, sample, m-tile) list:
//   warp 9 lane 0   loader: 8 activation stages (tensor-map TMA, 32 channels x 128 pixels, 128-byte swizzle) and 4 weight
//                   stages (bulk copies of pack.cu's stage images) are refilled as soon as the MMAs that read them
//                   retire, across tile boundaries; the two rings are advanced independently (non-blocking probes);
//   warps 0-3,10-13 in-place relu(x*scale+shift) of each landed activation stage;
//   warp 8 lane 0   tcgen05.mma kind::tf32 into a double-buffered TMEM accumulator (2 x 128 columns);
//   warps 4-7,14-17 two epilogue groups, one per TMEM buffer (even / odd tiles), so two tiles are drained while a third
//                   is being multiplied: TMEM -> registers -> 32-column staging -> full-line global stores; statistics by
//                   warp transpose-reduction into per-warp double registers, flushed to HBM once per (warp, sample, n-tile).
#include "tma_common.cuh"

namespace smg {

namespace {

constexpr int Q_STAGE = UM * 128;                  // 128 rows x 128 B
constexpr int Q_RES_MAX = 7;                       // resident mode: up to 7 weight stages (K <= 224) stay in shared memory
constexpr int Q_THREADS = 576;

// RES = true: the layer's KG <= 7 weight stages are loaded once per launch (every block-1 layer and the first five of
// block 2) - no weight re-fetch per tile, 6 activation stages; RES = false: weights stream through 3 stages, 8 activation stages.
template <bool RES>
struct Q1 {
    static constexpr int NA = RES ? 6 : 8;                // even: see the transform groups
    static constexpr int NB = RES ? Q_RES_MAX : 3;
    static constexpr int OFF_A = 0;
    static constexpr int OFF_B = OFF_A + NA * Q_STAGE;
    static constexpr int OFF_SC = OFF_B + NB * Q_STAGE;   // scale[1024], shift[1024]
    static constexpr int OFF_BAR = OFF_SC + 8192;
    static constexpr int TOTAL = OFF_BAR + 320;
    static_assert(TOTAL <= 232448, "shared-memory plan exceeds the 227 KB of one SM");
};

struct Tile1 {
    int nt, s, m0;
};

template <bool RES>
__global__ void __launch_bounds__(Q_THREADS, 1)
conv1_persist_kernel(const __grid_constant__ CUtensorMap tmA, UmmaDev a, int total_tiles, int n_samples, long long* __restrict__ trace) {
    constexpr int BN = 128;
    using Q = Q1<RES>;
    constexpr int Q_NA = Q::NA, Q_NB = Q::NB;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Q::OFF_BAR);
    uint64_t* raw_full = bars;          // [8] activations landed (raw)
    uint64_t* a_ready = bars + 8;       // [8] normalised (the 128 transform threads that own the stage)
    uint64_t* a_empty = bars + 16;      // [8] MMAs retired
    uint64_t* b_full = bars + 24;       // [4]
    uint64_t* b_empty = bars + 28;      // [4]
    uint64_t* t_full = bars + 32;       // [2] accumulator complete
    uint64_t* t_empty = bars + 34;      // [2] accumulator drained (128 epilogue threads)
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 36);
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
    // optional timeline of CTA 0 (SMG_CONV3_TRACE): trace[event][index], clock64 stamps of the first 64 stages / tiles
    auto stamp = [&](int event, int idx) {
        if (trace != nullptr && blockIdx.x == 0 && idx < 64) trace[event * 64 + idx] = clock64();
    };
    auto coord = [&](int tile) {
        Tile1 c;
        const int per_nt = tps * n_samples;
        c.nt = tile / per_nt;
        const int rem = tile - c.nt * per_nt;
        c.s = rem / tps;
        c.m0 = (rem - c.s * tps) * UM;
        return c;
    };

    if (warp == 8 && lane == 0) {
        if (smem_u32(smem) & 1023u) __trap();   // the swizzled stages rely on a 1024-byte aligned window
        for (int i = 0; i < Q_NA; ++i) { mbar_init(&raw_full[i], 1); mbar_init(&a_ready[i], 128); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < 4; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }   // RES: b_full[0] = weights landed
        for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 128); }
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc(tmem_ptr, 2 * BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const int total_stages = ntiles * KG;

    if (warp == 9) {
        // =============================== loader ===============================
        if (lane == 0) {
            // two independent rings: activations run up to Q_NA stages ahead of the MMAs, weights up to Q_NB
            int qa = 0, ta = 0, ka = 0;      // next activation stage: global index, tile, channel group
            int qb = 0, tb = 0, kb = 0;
            Tile1 ca = coord(tile_begin), cb = ca;
            if (RES) {
                // one n-tile per launch in this mode: the whole weight matrix of the layer, once
                mbar_arrive_expect_tx(&b_full[0], (uint32_t)KG * Q_STAGE);
                for (int kg = 0; kg < KG; ++kg) tma_bulk_load(sB + kg * Q_STAGE, a.w + (size_t)kg * Q_STAGE, Q_STAGE, &b_full[0]);
                qb = total_stages;
            }
            while (qa < total_stages || qb < total_stages) {
                if (qa < total_stages && mbar_test(&a_empty[qa % Q_NA], ((qa / Q_NA) & 1) ^ 1)) {
                    const int slot = qa % Q_NA;
                    stamp(0, qa);
                    mbar_arrive_expect_tx(&raw_full[slot], Q_STAGE);
                    tma_tile_3d(sA + slot * Q_STAGE, &tmA, ka * KC, ca.m0, ca.s, &raw_full[slot]);
                    ++qa;
                    if (++ka == KG) {
                        ka = 0;
                        if (++ta < ntiles) ca = coord(tile_begin + ta);
                    }
                }
                if (!RES && qb < total_stages && mbar_test(&b_empty[qb % Q_NB], ((qb / Q_NB) & 1) ^ 1)) {
                    const int slot = qb % Q_NB;
                    mbar_arrive_expect_tx(&b_full[slot], Q_STAGE);
                    tma_bulk_load(sB + slot * Q_STAGE, a.w + ((size_t)cb.nt * KG + kb) * Q_STAGE, Q_STAGE, &b_full[slot]);
                    ++qb;
                    if (++kb == KG) {
                        kb = 0;
                        if (++tb < ntiles) cb = coord(tile_begin + tb);
                    }
                }
            }
        }
    } else if (warp < 4 || (warp >= 10 && warp < 14)) {
        // =============================== in-place transform ===============================
        // two groups of four warps; group g owns the stages with global index = g (mod 2), so two stages are being
        // normalised concurrently.  NA is even: a slot is always served by the same group, which is what keeps the
        // one-bit barrier parity unambiguous (a group that never saw phase k of a slot must not wait for phase k+1 of it).
        const int ptid = warp < 4 ? tid : tid - 192;          // 0..255
        const int grp = ptid >> 7;                            // 0..1
        const int gt = ptid & 127;
        const int j = gt & 7;                                 // physical 16-byte piece of the 128-byte row
        const int rbase = gt >> 3;                            // rows rbase + 16 i, i < 8
        const int chunk = j ^ (rbase & 7);                    // logical 4-channel chunk held by that piece
        int cur_s = -1;
        int q0 = 0;                                           // global index of the tile's first stage
        for (int it = 0; it < ntiles; ++it, q0 += KG) {
            const Tile1 c = coord(tile_begin + it);
            if (c.s != cur_s) {
                // BN scale/shift of the new sample (every transform thread has left the previous sample's tables)
                asm volatile("bar.sync 2, 256;" ::: "memory");
                for (int ch = ptid; ch < a.cin; ch += 256) {
                    float sc, sh;
                    if (a.prologue_mode == 0) {
                        const double cnt = (double)a.hin * a.hin;
                        const double* st = a.in_stats + 2 * ((size_t)c.s * a.in_stats_stride + ch);
                        const double m = st[0] / cnt;
                        double var = st[1] / cnt - m * m;
                        if (var < 0) var = 0;
                        sc = a.gamma[ch] * (float)(1.0 / sqrt(var + (double)kBnEps));
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
                const int slot = q % Q_NA;
                const float4 sc = *reinterpret_cast<const float4*>(s_sc + kg * KC + chunk * 4);
                const float4 sh = *reinterpret_cast<const float4*>(s_sh + kg * KC + chunk * 4);
                mbar_wait_sleep(&raw_full[slot], (q / Q_NA) & 1, 64);
                if (gt == 0) stamp(1, q);
                uint8_t* base = sA + slot * Q_STAGE + rbase * 128 + j * 16;
                float4 x[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = *reinterpret_cast<const float4*>(base + i * 16 * 128);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float4 y;
                    y.x = fmaf(x[i].x, sc.x, sh.x); y.y = fmaf(x[i].y, sc.y, sh.y);
                    y.z = fmaf(x[i].z, sc.z, sh.z); y.w = fmaf(x[i].w, sc.w, sh.w);
                    if (a.relu) { y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f); }
                    if (i * 16 >= nvalid) y = make_float4(0.f, 0.f, 0.f, 0.f);   // rows beyond the sample contribute nothing
                    *reinterpret_cast<float4*>(base + i * 16 * 128) = y;
                }
                fence_proxy_async();
                mbar_arrive(&a_ready[slot]);
                if (gt == 0) stamp(2, q);
            }
        }
    } else if (warp == 8) {
        // =============================== MMA issuer ===============================
        if (lane == 0) {
            constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                       ((uint32_t)(UM >> 4) << 24);
            const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB);
            int q = 0;
            if (RES) mbar_wait(&b_full[0], 0);
            for (int it = 0; it < ntiles; ++it) {
                const int buf = it & 1;
                mbar_wait(&t_empty[buf], ((it >> 1) & 1) ^ 1);   // the epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
                uint32_t accum = 0;
                for (int kg = 0; kg < KG; ++kg, ++q) {
                    const int sa = q % Q_NA, sb = RES ? kg : q % Q_NB;
                    mbar_wait(&a_ready[sa], (q / Q_NA) & 1);
                    stamp(3, q);
                    if (!RES) mbar_wait(&b_full[sb], (q / Q_NB) & 1);
                    tc_fence_after();
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t ad = make_desc_sw128(sA_u + sa * Q_STAGE + k * 32);
                        const uint64_t bd = make_desc(sB_u + sb * Q_STAGE + 2 * k * BN * 16, BN * 16, 128);
                        umma<4>(d_tmem, ad, bd, idesc, accum);
                        accum = 1;
                    }
                    umma_commit(&a_empty[sa]);
                    stamp(4, q);
                    if (!RES) umma_commit(&b_empty[sb]);
                }
                umma_commit(&t_full[buf]);
            }
        }
    } else {
        // =============================== epilogue (warps 4-7: even tiles, warps 14-17: odd tiles) ===============================
        const int eg = warp < 8 ? 0 : 1;     // epilogue group == TMEM accumulator buffer it drains
        const int e = warp & 3;              // TMEM lane partition of this warp
        const int row = e * 32 + lane;       // accumulator row == tile row
        // per-warp statistics of the rows this warp drained, lane = channel within a 32-column chunk; flushed to HBM when
        // the (sample, n-tile) changes
        double acc_su[4] = {0.0, 0.0, 0.0, 0.0}, acc_ss[4] = {0.0, 0.0, 0.0, 0.0};
        int cur_s = -1, cur_nt = -1;
        auto flush = [&](int s_done, int nt_done) {
            if (a.out_stats == nullptr) return;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                double* st = a.out_stats + 2 * ((size_t)s_done * a.out_stats_stride + a.out_coff + nt_done * BN + k * 32 + lane);
                atomicAdd(st, acc_su[k]);
                atomicAdd(st + 1, acc_ss[k]);
                acc_su[k] = 0.0;
                acc_ss[k] = 0.0;
            }
        };
        for (int it = eg; it < ntiles; it += 2) {
            const Tile1 c = coord(tile_begin + it);
            if (c.s != cur_s || c.nt != cur_nt) {
                if (cur_s >= 0) flush(cur_s, cur_nt);
                cur_s = c.s;
                cur_nt = c.nt;
            }
            const bool valid = c.m0 + row < hw_out;
            float* orow = a.out + ((size_t)c.s * hw_out + c.m0 + row) * a.out_cstride + a.out_coff + c.nt * BN;
            mbar_wait_sleep(&t_full[eg], (it >> 1) & 1, 128);
            tc_fence_after();
            if (e == 0 && lane == 0) stamp(5, it);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(e * 32) << 16) + (uint32_t)(eg * BN + k * 32), v);
                if (k == 3) {
                    tc_fence_before();
                    mbar_arrive(&t_empty[eg]);   // the accumulator is in registers: the MMA warp may overwrite it
                }
                // straight from registers: each thread writes its own row (8 x 16 B = one 128-byte line per chunk); no
                // shared-memory staging - these layers sit on the SM's shared-memory bandwidth
                if (valid) {
                    float4* o = reinterpret_cast<float4*>(orow + k * 32);
#pragma unroll
                    for (int q4 = 0; q4 < 8; ++q4) o[q4] = make_float4(v[4 * q4], v[4 * q4 + 1], v[4 * q4 + 2], v[4 * q4 + 3]);
                }
                if (a.out_stats != nullptr) {
                    float sq[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        if (!valid) v[i] = 0.f;
                        sq[i] = v[i] * v[i];
                    }
                    acc_su[k] += (double)warp_transpose_sum(v, lane);
                    acc_ss[k] += (double)warp_transpose_sum(sq, lane);
                }
            }
            if (e == 0 && lane == 0) stamp(6, it);
        }
        if (cur_s >= 0) flush(cur_s, cur_nt);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 2 * BN);
    }
}

}  // namespace

// Returns SMG_ERR_UNSUPPORTED for shapes this kernel does not serve (the caller then uses the one-tile kernels).
int launch_conv1_persist(smg_handle* h, const ConvArgs& a, cudaStream_t st) {
    if (a.taps != 1 || a.pool || a.cout % 128 != 0 || a.cin % KC != 0 || a.cin > 1024 || a.in_cstride % 4 != 0 ||
        a.out_cstride % 4 != 0 || a.out_coff % 4 != 0 || (reinterpret_cast<uintptr_t>(a.in) & 15) != 0 ||
        (reinterpret_cast<uintptr_t>(a.out) & 15) != 0)
        return SMG_ERR_UNSUPPORTED;
    SMG_CHECK(a.w != nullptr && a.w->w_tf32 != nullptr, SMG_ERR_STATE, "conv1_persist: weights not packed");
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
    const int total = d.tiles_per_sample * a.n * (a.cout / 128);
    static bool attr = false;
    if (!attr) {
        SMG_CUDA(cudaFuncSetAttribute(conv1_persist_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Q1<true>::TOTAL));
        SMG_CUDA(cudaFuncSetAttribute(conv1_persist_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Q1<false>::TOTAL));
        attr = true;
    }
    const int grid = total < h->num_sms ? total : h->num_sms;
    // debugging aid: SMG_CONV1_TRACE=<file> appends the clock64 timeline of CTA 0 of every launch (synchronises!)
    static const char* trace_path = getenv("SMG_CONV1_TRACE");
    static long long* trace_dev = nullptr;
    if (trace_path != nullptr && trace_dev == nullptr) SMG_CUDA(cudaMalloc(&trace_dev, 7 * 64 * sizeof(long long)));
    if (trace_dev != nullptr) SMG_CUDA(cudaMemsetAsync(trace_dev, 0, 7 * 64 * sizeof(long long), st));
    const bool res = a.cout == 128 && a.cin / KC <= Q_RES_MAX;
    if (res) conv1_persist_kernel<true><<<grid, Q_THREADS, Q1<true>::TOTAL, st>>>(tm, d, total, a.n, trace_dev);
    else conv1_persist_kernel<false><<<grid, Q_THREADS, Q1<false>::TOTAL, st>>>(tm, d, total, a.n, trace_dev);
    h->launches++;
    SMG_CUDA(cudaGetLastError());
    if (trace_dev != nullptr) {
        long long host[7 * 64];
        SMG_CUDA(cudaMemcpyAsync(host, trace_dev, sizeof(host), cudaMemcpyDeviceToHost, st));
        SMG_CUDA(cudaStreamSynchronize(st));
        if (FILE* f = fopen(trace_path, "a")) {
            fprintf(f, "launch hin=%d n=%d cin=%d resident=%d total_tiles=%d grid=%d\n", a.hin, a.n, a.cin, (int)res, total, grid);
            for (int e = 0; e < 7; ++e) {
                for (int i = 0; i < 64; ++i) fprintf(f, "%lld ", host[e * 64 + i]);
                fprintf(f, "\n");
            }
            fclose(f);
        }
    }
    return SMG_OK;
}

}  // namespace smg
