// conv3_wt.cu - persistent tf32 3x3 convolution (128 -> 32 channels) with the WEIGHTS resident in tensor memory.
//
// Serves torchvision densenet `_DenseLayer.conv2` (/root/reference/code/models.py:319 builds the trunks); same contract as
// conv3_persist.cu (BN-ReLU prologue, raw NHWC output slice, (sum, sumsq) epilogue), operand roles swapped:
//
//   D[(dx,co)][pixel m] += W[(dx,co)][(dy,cin)] * X[m + dy*wp][cin]          (= E_dx[m][co] of conv3_persist.cu)
//
//   A operand = weights, 96 rows (dx*32 + co; rows 96-127 unused), K = 3 dy x 128 cin = 384 tf32 -> 384 columns of TENSOR
//               memory, written once per launch with tcgen05.st and read by every MMA from there;
//   B operand = the normalised activation patch in shared memory (rows = pixels, 128-byte swizzle), N = 128 pixels; the
//               kernel-row shift dy*wp is a start-address offset exactly as before, now on the B descriptor;
//   D         = 128 columns of tensor memory (lanes = (dx, co), columns = pixels); 384 + 128 = all 512 columns.
//
// What this buys (profiles/r01_conv3_timeline.txt: conv3_persist.cu is bound by the bytes it can keep in flight):
//   * no weights in shared memory: SIX 27 KB patch slots instead of three next to 144 KB of weights;
//   * the 147 KB of weight operand reads per tile disappear from the shared-memory pipe;
//   * the dx shift of the output combination is a COLUMN offset, lanes are channels: the epilogue needs no shuffles.
//     The six epilogue warps write their accumulator columns, shifted by their dx, to a 48 KB exchange buffer; each
//     then sums the three dx planes for a sixth of the pixels, stores one 128-byte pixel row per instruction and keeps
//     the per-channel statistics in per-thread registers.
//   * the accumulator is single-buffered, but it is only held until the six epilogue warps have copied their 64
//     columns into registers; everything after that overlaps the next tile's MMAs.
// Warps (512 threads): 0-3 and 8-11 in-place transform, 4-6 and 12-14 epilogue (quadrant = warp mod 4, pixel half =
// warp / 8), 7 MMA issuer, 15 TMA loader.
#include "tma_common.cuh"

namespace smg {

namespace {

constexpr int W_NSLOT = 6;
constexpr int W_SLOT = 27 * 1024;                 // >= 212 patch rows x 128 B, multiple of the 1024-byte swizzle period
constexpr int W_ROWS = 212;                       // patch rows touched: TMA box <= 210, the last kernel row reads up to 2*wp + 127
constexpr int W_KCOLS = 384;                      // weight columns in tensor memory: ((g*3 + dy)*4 + k)*8 + e
constexpr int W_OFF_A = 0;
constexpr int W_OFF_X = W_OFF_A + W_NSLOT * W_SLOT;      // exchange [3 dx][128 px][32 co] floats
constexpr int W_OFF_SC = W_OFF_X + 3 * 128 * 32 * 4;     // scale[128], shift[128]
constexpr int W_OFF_BAR = W_OFF_SC + 1024;
constexpr int W_TOTAL = W_OFF_BAR + 256;
constexpr int W_THREADS = 512;
static_assert(W_TOTAL <= 232448, "shared-memory plan exceeds the 227 KB of one SM");

struct TileCoord {
    int s, h0, w0;
};

__global__ void __launch_bounds__(W_THREADS, 1)
conv3_wt_kernel(const __grid_constant__ CUtensorMap tmA, UmmaDev a, int total_tiles, long long* __restrict__ trace) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + W_OFF_BAR);
    uint64_t* raw_full = bars;          // [<= 7] patch landed (raw)
    uint64_t* a_ready = bars + 7;       // [<= 7] patch normalised (256 transform threads)
    uint64_t* a_empty = bars + 14;      // [<= 7] MMAs reading the slot retired
    uint64_t* t_full = bars + 21;       // accumulator complete
    uint64_t* t_empty = bars + 22;      // accumulator copied to registers (192 epilogue threads)
    uint64_t* w_ready = bars + 23;      // weights in tensor memory (192 threads)
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 24);
    float* s_sc = reinterpret_cast<float*>(smem + W_OFF_SC);
    float* s_sh = s_sc + 128;
    float* s_x = reinterpret_cast<float*>(smem + W_OFF_X);
    uint8_t* sA = smem + W_OFF_A;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int hout = a.hout, hin = a.hin;
    const int hw_out = hout * hout;
    const int wp = a.wp;
    const int pfill = (a.ht + 2) * wp;
    const int tps = a.tiles_per_sample;
    const int tile_begin = (int)(((long long)blockIdx.x * total_tiles) / gridDim.x);
    const int tile_end = (int)(((long long)(blockIdx.x + 1) * total_tiles) / gridDim.x);
    const int ntiles = tile_end - tile_begin;
    // optional timeline of CTA 0 (SMG_CONV3_TRACE): trace[event][index], clock64 stamps of the first 64 patch loads / tiles
    auto stamp = [&](int event, int idx) {
        if (trace != nullptr && blockIdx.x == 0 && idx < 64) trace[event * 64 + idx] = clock64();
    };
    auto coord = [&](int tile) {
        TileCoord c;
        c.s = tile / tps;
        const int rem = tile - c.s * tps;
        const int ty = rem / a.tiles_x, tx = rem - ty * a.tiles_x;
        c.h0 = ty * a.ht;
        c.w0 = tx * (wp - 2);
        return c;
    };
    const bool is_transform = (warp & 7) < 4;                    // warps 0-3, 8-11
    const bool is_epilogue = (warp & 7) >= 4 && (warp & 3) != 3; // warps 4-6, 12-14

    if (warp == 7 && lane == 0) {
        if (smem_u32(smem) & 1023u) __trap();   // the swizzled slots rely on a 1024-byte aligned window
        for (int i = 0; i < W_NSLOT; ++i) { mbar_init(&raw_full[i], 1); mbar_init(&a_ready[i], 256); mbar_init(&a_empty[i], 1); }
        mbar_init(t_full, 1);
        mbar_init(t_empty, 192);
        mbar_init(w_ready, 192);
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc(tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t tmem_d = tmem_base + W_KCOLS;
    pdl_launch_dependents();   // see conv1_t.cu: set-up and weight fill may overlap the previous kernel's tail

    if (warp == 15) {
        // =============================== loader ===============================
        if (lane == 0) {
            pdl_wait();
            for (int n = 0; n < 4 * ntiles; ++n) {
                const int slot = n % W_NSLOT;
                const TileCoord c = coord(tile_begin + (n >> 2));
                mbar_wait_sleep(&a_empty[slot], ((n / W_NSLOT) & 1) ^ 1, 64);
                stamp(0, n);
                mbar_arrive_expect_tx(&raw_full[slot], (uint32_t)pfill * 128u);
                tma_tile_4d(sA + slot * W_SLOT, &tmA, (n & 3) * KC, c.w0 - 1, c.h0 - 1, c.s, &raw_full[slot]);
            }
        }
    } else if (is_transform) {
        // =============================== in-place transform ===============================
        const int ptid = warp < 4 ? tid : tid - 128;          // 0..255
        const int j = ptid & 7;                               // physical 16-byte piece of the 128-byte row
        const int rbase = ptid >> 3;                          // patch rows rbase + 32 i
        const int chunk = j ^ (rbase & 7);                    // logical 4-channel chunk held by that piece
        constexpr int NI = 7;                                 // 7 x 32 = 224 >= patch rows
        int cur_s = -1;
        uint32_t inside = 0, filled = 0;
        pdl_wait();                                           // the scale/shift tables read the producer's statistics
        for (int n = 0; n < 4 * ntiles; ++n) {
            const int g = n & 3;
            const int slot = n % W_NSLOT;
            if (g == 0) {
                const TileCoord c = coord(tile_begin + (n >> 2));
                if (c.s != cur_s) {
                    // BN scale/shift of the new sample (every transform thread has left the previous sample's tables)
                    asm volatile("bar.sync 2, 256;" ::: "memory");
                    if (ptid < 128) {
                        float sc, sh;
                        if (a.prologue_mode == 0) {
                            const double cnt = (double)hin * hin;
                            const double* st = a.in_stats + 2 * ((size_t)c.s * a.in_stats_stride + ptid);
                            const double m = st[0] / cnt;
                            double var = st[1] / cnt - m * m;
                            if (var < 0) var = 0;
                            sc = a.gamma[ptid] * (float)(1.0 / sqrt(var + (double)kBnEps));
                            sh = a.beta[ptid] - (float)m * sc;
                        } else {
                            sc = a.scale[(size_t)c.s * a.cin + ptid];
                            sh = a.shift[(size_t)c.s * a.cin + ptid];
                        }
                        s_sc[ptid] = sc;
                        s_sh[ptid] = sh;
                    }
                    asm volatile("bar.sync 2, 256;" ::: "memory");
                    cur_s = c.s;
                }
                inside = 0;
                filled = 0;
#pragma unroll
                for (int i = 0; i < NI; ++i) {
                    const int q = rbase + 32 * i;
                    const int py = q / wp, px = q - py * wp;
                    const int y = c.h0 - 1 + py, x = c.w0 - 1 + px;
                    if (q < pfill) {
                        filled |= 1u << i;
                        if (y >= 0 && y < hin && x >= 0 && x < hin) inside |= 1u << i;
                    }
                }
            }
            const float4 sc = *reinterpret_cast<const float4*>(s_sc + g * KC + chunk * 4);
            const float4 sh = *reinterpret_cast<const float4*>(s_sh + g * KC + chunk * 4);
            mbar_wait_sleep(&raw_full[slot], (n / W_NSLOT) & 1, 64);
            if (ptid == 0) stamp(1, n);
            uint8_t* base = sA + slot * W_SLOT + rbase * 128 + j * 16;
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
            if (ptid == 0) stamp(2, n);
        }
    } else if (warp == 7) {
        // =============================== MMA issuer ===============================
        if (lane == 0) {
            // D[128 x 128] (fp32) += A[128 x 8] (tensor memory: weights) * B[128 x 8]^T (shared memory: activations)
            constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t sA_u = smem_u32(sA);
            mbar_wait(w_ready, 0);
            tc_fence_after();
            for (int it = 0; it < ntiles; ++it) {
                mbar_wait(t_empty, (it & 1) ^ 1);   // the epilogue holds the previous tile's accumulator in registers
                tc_fence_after();
                uint32_t accum = 0;
                for (int g = 0; g < 4; ++g) {
                    const int n = it * 4 + g;
                    const int slot = n % W_NSLOT;
                    mbar_wait(&a_ready[slot], (n / W_NSLOT) & 1);
                    tc_fence_after();
                    stamp(3, n);
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy) {
                        // kernel row dy = a shift of dy*wp patch rows = +128 B per row on the start address of the B operand
                        const uint32_t start = sA_u + slot * W_SLOT + dy * wp * 128;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t bd = make_desc_sw128(start + k * 32);
                            umma_ts_tf32(tmem_d, tmem_base + (uint32_t)(((g * 3 + dy) * 4 + k) * 8), bd, idesc, accum);
                            accum = 1;
                        }
                    }
                    umma_commit(&a_empty[slot]);
                    stamp(4, n);
                }
                umma_commit(t_full);
            }
        }
    } else if (is_epilogue) {
        // =============================== epilogue ===============================
        const int q = warp & 3;               // TMEM lane quadrant = dx
        const int half = warp >> 3;           // pixel columns 64 half .. + 63
        // ---- once: this quadrant's 32 weight rows into tensor memory; the two warps of a quadrant split the 24 column blocks.
        // Global layout [column block c16][row][16 floats]: a warp reads 2 KB contiguous per block.
        {
            const float4* wsrc = reinterpret_cast<const float4*>(a.w) + (size_t)(q * 32 + lane) * 4;
#pragma unroll 4
            for (int c16 = half * 12; c16 < half * 12 + 12; ++c16) {
                const float4* p4 = wsrc + (size_t)c16 * 96 * 4;
                const float4 w0 = __ldg(p4), w1 = __ldg(p4 + 1), w2 = __ldg(p4 + 2), w3 = __ldg(p4 + 3);
                tmem_st16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c16 * 16), w0, w1, w2, w3);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            mbar_arrive(w_ready);
        }
        const int ew = half * 3 + q;          // 0..5: this warp combines and stores pixels [22 ew, 22 ew + 22)
        double acc_su = 0.0, acc_ss = 0.0;    // statistics of channel `lane` over the pixels this warp stored
        int cur_s = -1;
        auto flush = [&](int s_done) {
            if (a.out_stats == nullptr) return;
            double* st = a.out_stats + 2 * ((size_t)s_done * a.out_stats_stride + a.out_coff + lane);
            atomicAdd(st, acc_su);
            atomicAdd(st + 1, acc_ss);
            acc_su = 0.0;
            acc_ss = 0.0;
        };
        for (int it = 0; it < ntiles; ++it) {
            const TileCoord c = coord(tile_begin + it);
            if (c.s != cur_s) {
                if (cur_s >= 0) flush(cur_s);
                cur_s = c.s;
            }
            mbar_wait_sleep(t_full, it & 1, 64);
            tc_fence_after();
            if (ew == 0 && lane == 0) stamp(5, it);
            float v[64];
            const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 64);
            tmem_ld32(taddr, v);
            tmem_ld32(taddr + 32, v + 32);
            tc_fence_before();
            mbar_arrive(t_empty);             // the accumulator is free for the next tile's MMAs
            // out[m][co] = E_0[m][co] + E_1[m+1][co] + E_2[m+2][co]: every quadrant publishes its columns shifted by its dx
            {
                float* x = s_x + q * 128 * 32;
#pragma unroll
                for (int i = 0; i < 64; ++i) {
                    const int m = half * 64 + i - q;          // output pixel this column contributes to
                    if (m >= 0) x[m * 32 + lane] = v[i];
                }
            }
            asm volatile("bar.sync 1, 192;" ::: "memory");
            {
                const float* x0 = s_x, *x1 = s_x + 128 * 32, *x2 = s_x + 2 * 128 * 32;
                float* obase = a.out + (size_t)c.s * hw_out * a.out_cstride + a.out_coff + lane;
                float su = 0.f, sq = 0.f;
                const int m0 = ew * 22;
                int ri = m0 / wp, rj = m0 - ri * wp;          // patch coordinates of pixel m, advanced incrementally
#pragma unroll
                for (int i = 0; i < 22; ++i) {
                    const int m = m0 + i;
                    const bool valid = m < 126 && ri < a.ht && rj < wp - 2 && c.h0 + ri < hout && c.w0 + rj < hout;   // warp-uniform
                    if (valid) {
                        const float o = (x0[m * 32 + lane] + x1[m * 32 + lane]) + x2[m * 32 + lane];
                        obase[(size_t)((c.h0 + ri) * hout + c.w0 + rj) * a.out_cstride] = o;   // 32 lanes = one 128-byte pixel row
                        su += o;
                        sq = fmaf(o, o, sq);
                    }
                    if (++rj == wp) { rj = 0; ++ri; }
                }
                acc_su += (double)su;
                acc_ss += (double)sq;
            }
            asm volatile("bar.sync 1, 192;" ::: "memory");   // the exchange buffer may be rewritten for the next tile
            if (ew == 0 && lane == 0) stamp(6, it);
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

// Returns SMG_ERR_UNSUPPORTED for shapes this kernel does not serve (the caller then uses conv3_persist.cu).
int launch_conv3_wt(smg_handle* h, const ConvArgs& a, cudaStream_t st) {
    if (a.taps != 9 || a.pool || a.cin != 128 || a.cout != 32 || a.in_cstride % 4 != 0 || a.w == nullptr ||
        a.w->w_tf32_t == nullptr || (reinterpret_cast<uintptr_t>(a.in) & 15) != 0)
        return SMG_ERR_UNSUPPORTED;
    UmmaDev d;
    d.in = a.in; d.in_cstride = a.in_cstride; d.cin = a.cin; d.hin = a.hin;
    d.prologue_mode = a.prologue_mode; d.in_stats = a.in_stats; d.in_stats_stride = a.in_stats_stride;
    d.gamma = a.gamma; d.beta = a.beta; d.scale = a.scale; d.shift = a.shift; d.relu = a.relu;
    d.w = a.w->w_tf32_t;
    d.out = a.out; d.out_cstride = a.out_cstride; d.out_coff = a.out_coff; d.cout = a.cout;
    d.out_stats = a.out_stats; d.out_stats_stride = a.out_stats_stride;
    d.hout = a.hin;
    umma_patch_geometry(d.hout, &d.wp, &d.ht);
    SMG_CHECK(2 * d.wp + UM <= W_ROWS && (d.ht + 2) * d.wp <= W_ROWS && d.ht * d.wp <= UM, SMG_ERR_STATE,
              "conv3_wt: patch %dx%d too large", d.ht, d.wp);
    const int wt = d.wp - 2;
    d.tiles_x = (d.hout + wt - 1) / wt;
    d.tiles_per_sample = d.tiles_x * ((d.hout + d.ht - 1) / d.ht);
    d.tiles_per_cta = 0;
    d.async_producer = 0;
    const int total = d.tiles_per_sample * a.n;

    CUtensorMap tm;
    const cuuint64_t dims[4] = {(cuuint64_t)a.in_cstride, (cuuint64_t)a.hin, (cuuint64_t)a.hin, (cuuint64_t)a.n};
    const cuuint64_t strides[3] = {(cuuint64_t)a.in_cstride * 4, (cuuint64_t)a.hin * a.in_cstride * 4,
                                   (cuuint64_t)a.hin * a.hin * a.in_cstride * 4};
    const cuuint32_t box[4] = {KC, (cuuint32_t)d.wp, (cuuint32_t)(d.ht + 2), 1};
    SMG_TRY(make_tensor_map_f32(&tm, a.in, 4, dims, strides, box, 128));
    SMG_TRY(ensure_dyn_smem(h, (const void*)conv3_wt_kernel, W_TOTAL));
    const int grid = total < h->num_sms ? total : h->num_sms;
    // debugging aid: SMG_CONV3_TRACE=<file> appends the clock64 timeline of CTA 0 of every launch (synchronises!)
    static const char* trace_path = getenv("SMG_CONV3_TRACE");
    static long long* trace_dev = nullptr;
    if (trace_path != nullptr && trace_dev == nullptr) SMG_CUDA(cudaMalloc(&trace_dev, 7 * 64 * sizeof(long long)));
    if (trace_dev != nullptr) SMG_CUDA(cudaMemsetAsync(trace_dev, 0, 7 * 64 * sizeof(long long), st));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(W_THREADS);
    cfg.dynamicSmemBytes = W_TOTAL;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = h->use_pdl ? 1 : 0;
    SMG_CUDA(cudaLaunchKernelEx(&cfg, conv3_wt_kernel, tm, d, total, trace_dev));
    h->launches++;
    SMG_CUDA(cudaGetLastError());
    if (trace_dev != nullptr) {
        long long host[7 * 64];
        SMG_CUDA(cudaMemcpyAsync(host, trace_dev, sizeof(host), cudaMemcpyDeviceToHost, st));
        SMG_CUDA(cudaStreamSynchronize(st));
        if (FILE* f = fopen(trace_path, "a")) {
            fprintf(f, "launch hin=%d n=%d total_tiles=%d grid=%d\n", a.hin, a.n, total, grid);
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
