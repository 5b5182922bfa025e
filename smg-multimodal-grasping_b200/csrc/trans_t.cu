// trans_t.cu - persistent tf32 transition: BN-ReLU -> 2x2 average pool -> 1x1 convolution (cin -> cin / 2), TMA-fed.
//
// Serves torchvision densenet `_Transition` (norm, relu, conv 1x1, AvgPool2d(2); /root/reference/code/models.py:319 builds
// the trunks).  Pool and convolution are both linear and the pool follows the ReLU, so the kernel pools FIRST (4x fewer
// MMAs and accumulator reads than convolving the full-resolution map):
//
//   D[cout][pooled pixel p] += W[cout][cin] * P[p][cin],      P = mean over the 2x2 window of relu(bn(x))
//
//   raw stages   : tensor-map TMA boxes {32 channels, 2 tw, 2 sr} = 4 tw sr <= 128 raw pixel rows of 128 B (128-byte
//                  swizzle), i.e. sr pooled rows of a tw-wide pooled tile; four consecutive boxes cover the tile
//                  (N = 4 tw sr pooled pixels <= 128);
//   pool warps   : read the four raw rows of every pooled pixel, apply the BatchNorm scale/shift + ReLU, average, and
//                  write the pooled row into an operand stage (the B operand: N rows x 32 channels, 128-byte swizzle);
//   A operand    : the weights, streamed through shared memory as pack.cu's K-major stage images (128 cout x 32 cin each;
//                  MB of them per channel group for MB x 128 output channels);
//   D            : [cout lanes][pooled pixel columns] in tensor memory, MB x 128 columns, double buffered.
// Layers with more than MB x 128 output channels are split into output-channel parts walked concurrently by the two halves
// of the grid.  Epilogue as in conv1_t.cu: a warp owns 32 channels, every store instruction
// writes one pixel's 128 contiguous bytes, statistics are per-thread sums.
// Warps (576 threads): 0-3 / 8-11 pool-transform (256 threads, every raw box in order), 4-7 epilogue of even
// tiles, 12-15 epilogue of odd tiles (quadrant = warp mod 4), 16 MMA issuer, 17 TMA loader.
#include "tma_common.cuh"

namespace smg {

namespace {

constexpr int X_STAGE = UM * 128;                  // 16 KB: 128 rows x 128 B
constexpr int X_THREADS = 576;

template <int MB>
struct XP {
    static constexpr int NR = MB == 1 ? 7 : 6;            // raw stages
    static constexpr int NO = MB == 1 ? 4 : 2;            // operand stages
    static constexpr int NW = 2;                          // weight stages of MB x 16 KB
    static constexpr int OFF_R = 0;
    static constexpr int OFF_O = OFF_R + NR * X_STAGE;
    static constexpr int OFF_W = OFF_O + NO * X_STAGE;
    static constexpr int OFF_SC = OFF_W + NW * MB * X_STAGE;   // scale[1024], shift[1024]
    static constexpr int OFF_BAR = OFF_SC + 8192;
    static constexpr int TOTAL = OFF_BAR + 512;
    static_assert(TOTAL <= 232448, "shared-memory plan exceeds the 227 KB of one SM");
};

struct XGeom {
    int tw, sr;            // pooled tile width, pooled rows per raw box; tile = tw x 4 sr pooled pixels
    int tiles_x, tiles_y;  // spatial tiles per sample
    int parts;             // output-channel parts of MB x 128 channels
    int kg_total;          // channel groups of the weight image (cin / 32)
};

struct XTile {
    int s, oy0, ox0, part;
};

#ifdef SMG_DEBUG_HANG
// debugging aid: a wait that does not complete within ~1 s reports where it is stuck and traps
__device__ __noinline__ void dbg_wait(uint64_t* bar, uint32_t parity, int what, int idx) {
    const long long t0 = clock64();
    while (!mbar_test(bar, parity)) {
        if (clock64() - t0 > 2000000000ll) {
            printf("trans_t stuck: block %d thread %d wait %d idx %d parity %u\n", blockIdx.x, threadIdx.x, what, idx, parity);
            __trap();
        }
    }
}
#define XWAIT(bar, parity, what, idx) dbg_wait(bar, parity, what, idx)
#define XWAIT_SLEEP(bar, parity, ns, what, idx) dbg_wait(bar, parity, what, idx)
#else
#define XWAIT(bar, parity, what, idx) mbar_wait(bar, parity)
#define XWAIT_SLEEP(bar, parity, ns, what, idx) mbar_wait_sleep(bar, parity, ns)
#endif

template <int MB>
__global__ void __launch_bounds__(X_THREADS, 1)
trans_t_kernel(const __grid_constant__ CUtensorMap tmA, UmmaDev a, XGeom g, int total_tiles) {
    using Q = XP<MB>;
    constexpr int NR = Q::NR, NO = Q::NO, NW = Q::NW;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Q::OFF_BAR);
    uint64_t* raw_full = bars;          // [8] raw box landed
    uint64_t* raw_empty = bars + 8;     // [8] read by the 256 pool threads
    uint64_t* op_ready = bars + 16;     // [4] pooled operand written (256 threads)
    uint64_t* op_empty = bars + 20;     // [4] MMAs retired
    uint64_t* b_full = bars + 24;       // [2] weight stage landed
    uint64_t* b_empty = bars + 26;      // [2]
    uint64_t* t_full = bars + 28;       // [2] accumulator complete
    uint64_t* t_empty = bars + 30;      // [2] accumulator drained (128 epilogue threads)
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 32);
    float* s_sc = reinterpret_cast<float*>(smem + Q::OFF_SC);
    float* s_sh = s_sc + 1024;
    uint8_t* sR = smem + Q::OFF_R;
    uint8_t* sO = smem + Q::OFF_O;
    uint8_t* sW = smem + Q::OFF_W;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int hout = a.hout;
    const int hw_out = hout * hout;
    const int KG = a.cin / KC;
    const int npp = g.tw * g.sr;                          // pooled pixels per raw box
    const int ncol = 4 * npp;                             // pooled pixels per tile = MMA N
    // tile order: part-major.  With two output-channel parts the CTAs of the first half of the grid walk part 0 and those of
    // the second half walk part 1 over the SAME raw boxes at the same time, so one of the two reads is an L2 hit (walking
    // both parts back to back in one CTA re-read everything from HBM: 148 CTAs x 2 MB between the two uses exceed L2).
    const int sp_tiles = g.tiles_x * g.tiles_y;
    const int per_part = total_tiles / g.parts;
    const int tile_begin = (int)(((long long)blockIdx.x * total_tiles) / gridDim.x);
    const int tile_end = (int)(((long long)(blockIdx.x + 1) * total_tiles) / gridDim.x);
    const int ntiles = tile_end - tile_begin;
    auto coord = [&](int tile) {
        XTile c;
        c.part = tile / per_part;
        const int rem = tile - c.part * per_part;
        c.s = rem / sp_tiles;
        const int sp = rem - c.s * sp_tiles;
        const int ty = sp / g.tiles_x;
        c.oy0 = ty * 4 * g.sr;
        c.ox0 = (sp - ty * g.tiles_x) * g.tw;
        return c;
    };
    const bool is_transform = warp < 16 && (warp & 7) < 4;    // warps 0-3, 8-11
    const bool is_epilogue = warp < 16 && (warp & 7) >= 4;    // warps 4-7, 12-15

    if (warp == 16 && lane == 0) {
        if (smem_u32(smem) & 1023u) __trap();   // the swizzled stages rely on a 1024-byte aligned window
        for (int i = 0; i < NR; ++i) { mbar_init(&raw_full[i], 1); mbar_init(&raw_empty[i], 256); }
        for (int i = 0; i < NO; ++i) { mbar_init(&op_ready[i], 256); mbar_init(&op_empty[i], 1); }
        for (int i = 0; i < NW; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 128); }
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc(tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;          // accumulator buffer b: columns [b * MB * 128, (b + 1) * MB * 128)
    const int total_ops = ntiles * KG;             // operand stages this CTA produces
    pdl_launch_dependents();                       // see conv1_t.cu: nothing above reads the previous kernel's output

    if (warp == 17) {
        // =============================== loader ===============================
        if (lane == 0) {
            pdl_wait();
            const uint32_t raw_bytes = (uint32_t)(4 * npp) * 128u;
            int qa = 0, ta = 0, ka = 0, sa = 0;      // next raw box: global index, tile, channel group, sub-box
            int qb = 0, tb = 0, kb = 0;              // next weight stage
            XTile ca = coord(tile_begin), cb = ca;
            const int total_raw = total_ops * 4;
#ifdef SMG_DEBUG_HANG
            long long t_last = clock64();
#endif
            while (qa < total_raw || qb < total_ops) {
#ifdef SMG_DEBUG_HANG
                if (clock64() - t_last > 2000000000ll) {
                    printf("trans_t stuck: block %d loader qa %d / %d qb %d / %d ntiles %d KG %d\n", blockIdx.x, qa, total_raw, qb, total_ops, ntiles, KG);
                    __trap();
                }
#endif
                if (qa < total_raw && mbar_test(&raw_empty[qa % NR], ((qa / NR) & 1) ^ 1)) {
                    const int slot = qa % NR;
                    mbar_arrive_expect_tx(&raw_full[slot], raw_bytes);
                    tma_tile_4d(sR + slot * X_STAGE, &tmA, ka * KC, 2 * ca.ox0, 2 * (ca.oy0 + sa * g.sr), ca.s, &raw_full[slot]);
                    ++qa;
#ifdef SMG_DEBUG_HANG
                    t_last = clock64();
#endif
                    if (++sa == 4) {
                        sa = 0;
                        if (++ka == KG) {
                            ka = 0;
                            if (++ta < ntiles) ca = coord(tile_begin + ta);
                        }
                    }
                }
                if (qb < total_ops && mbar_test(&b_empty[qb % NW], ((qb / NW) & 1) ^ 1)) {
                    const int slot = qb % NW;
                    mbar_arrive_expect_tx(&b_full[slot], MB * X_STAGE);
#pragma unroll
                    for (int mb = 0; mb < MB; ++mb)
                        tma_bulk_load(sW + (slot * MB + mb) * X_STAGE,
                                      a.w + ((size_t)(cb.part * MB + mb) * g.kg_total + kb) * X_STAGE, X_STAGE, &b_full[slot]);
                    ++qb;
                    if (++kb == KG) {
                        kb = 0;
                        if (++tb < ntiles) cb = coord(tile_begin + tb);
                    }
                }
            }
        }
    } else if (is_transform) {
        // =============================== pool + transform ===============================
        // All 256 threads consume EVERY raw box, in order: thread = (pooled pixel pp of the box, 16-byte channel chunk j).
        // (The first version split the operand stages between two 128-thread groups that shared the raw ring: a group
        // could then wait on a raw slot two uses ahead of the other group's pending use of the same slot, and the one-bit
        // barrier parity aliased - rare wrong pooled values and hangs inside the network, never in isolation.)
        const int ptid = warp < 4 ? tid : tid - 128;          // 0..255
        const int j = ptid & 7;                               // logical 4-channel chunk
        const int pp = ptid >> 3;                             // pooled pixel of the raw box (32 per box at most)
        const int ppy = pp / g.tw, ppx = pp - ppy * g.tw;
        const bool has_px = pp < npp;
        const int r00 = (2 * ppy) * (2 * g.tw) + 2 * ppx;     // top-left raw row of the pooled pixel
        const int rw = 2 * g.tw;                              // raw rows per raw image row of the box
        // chunk j of raw rows r00, r00 + 1, r00 + rw, r00 + rw + 1 under the 128-byte swizzle
        const int o0 = r00 * 128 + ((j ^ (r00 & 7)) << 4);
        const int o1 = (r00 + 1) * 128 + ((j ^ ((r00 + 1) & 7)) << 4);
        const int o2 = (r00 + rw) * 128 + ((j ^ ((r00 + rw) & 7)) << 4);
        const int o3 = (r00 + rw + 1) * 128 + ((j ^ ((r00 + rw + 1) & 7)) << 4);
        int cur_s = -1;
        int q = 0;                                            // operand stage (global index inside this CTA)
        pdl_wait();                                           // the tables below read the producers' statistics
        for (int it = 0; it < ntiles; ++it) {
            const XTile c = coord(tile_begin + it);
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
            for (int kg = 0; kg < KG; ++kg, ++q) {
                const int oslot = q % NO;
                const float4 sc = *reinterpret_cast<const float4*>(s_sc + kg * KC + j * 4);
                const float4 sh = *reinterpret_cast<const float4*>(s_sh + kg * KC + j * 4);
                XWAIT_SLEEP(&op_empty[oslot], ((q / NO) & 1) ^ 1, 32, 1, q);
                uint8_t* op = sO + oslot * X_STAGE;
#pragma unroll
                for (int sub = 0; sub < 4; ++sub) {
                    const int qr = q * 4 + sub;
                    const int rslot = qr % NR;
                    const uint8_t* stage = sR + rslot * X_STAGE;
                    XWAIT_SLEEP(&raw_full[rslot], (qr / NR) & 1, 64, 2, qr);
                    if (has_px) {
                        float4 x[4];
                        x[0] = *reinterpret_cast<const float4*>(stage + o0);
                        x[1] = *reinterpret_cast<const float4*>(stage + o1);
                        x[2] = *reinterpret_cast<const float4*>(stage + o2);
                        x[3] = *reinterpret_cast<const float4*>(stage + o3);
                        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            float4 y;
                            y.x = fmaf(x[t].x, sc.x, sh.x); y.y = fmaf(x[t].y, sc.y, sh.y);
                            y.z = fmaf(x[t].z, sc.z, sh.z); y.w = fmaf(x[t].w, sc.w, sh.w);
                            if (a.relu) { y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f); }
                            acc.x += y.x; acc.y += y.y; acc.z += y.z; acc.w += y.w;
                        }
                        acc.x *= 0.25f; acc.y *= 0.25f; acc.z *= 0.25f; acc.w *= 0.25f;
                        const int p = sub * npp + pp;
                        *reinterpret_cast<float4*>(op + p * 128 + ((j ^ (p & 7)) << 4)) = acc;
                    }
                    // the raw values are in registers: order the generic-proxy reads before the copy engine's refill
                    fence_proxy_async();
                    mbar_arrive(&raw_empty[rslot]);
                }
                fence_proxy_async();
                mbar_arrive(&op_ready[oslot]);
            }
        }
    } else if (warp == 16) {
        // =============================== MMA issuer ===============================
        if (lane == 0) {
            // D[128 cout x ncol px] += A[128 x 8] (weights, shared memory) * B[ncol x 8]^T (pooled activations, shared memory)
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(ncol >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t sO_u = smem_u32(sO), sW_u = smem_u32(sW);
            int q = 0;
            for (int it = 0; it < ntiles; ++it) {
                const int buf = it & 1;
                XWAIT(&t_empty[buf], ((it >> 1) & 1) ^ 1, 3, it);   // the epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(buf * MB * 128);
                for (int kg = 0; kg < KG; ++kg, ++q) {
                    const int so = q % NO, sb = q % NW;
                    XWAIT(&op_ready[so], (q / NO) & 1, 4, q);
                    XWAIT(&b_full[sb], (q / NW) & 1, 5, q);
                    tc_fence_after();
#pragma unroll
                    for (int mb = 0; mb < MB; ++mb) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t xd = make_desc_sw128(sO_u + so * X_STAGE + k * 32);
                            const uint64_t wd = make_desc(sW_u + (sb * MB + mb) * X_STAGE + 2 * k * 128 * 16, 128 * 16, 128);
                            umma<4>(d_tmem + (uint32_t)(mb * 128), wd, xd, idesc, (kg | k) != 0 ? 1u : 0u);
                        }
                    }
                    umma_commit(&op_empty[so]);
                    umma_commit(&b_empty[sb]);
                }
                umma_commit(&t_full[buf]);
            }
        }
    } else if (is_epilogue) {
        // =============================== epilogue (warps 4-7: even tiles, 12-15: odd tiles) ===============================
        const int eg = warp >> 3;            // epilogue group == accumulator buffer it drains
        const int q4 = warp & 3;             // TMEM lane quadrant: output channels 32 q4 + lane of every 128-channel block
        double acc_su[MB], acc_ss[MB];
#pragma unroll
        for (int mb = 0; mb < MB; ++mb) { acc_su[mb] = 0.0; acc_ss[mb] = 0.0; }
        int cur_s = -1, cur_part = -1;
        auto flush = [&](int s_done, int part_done) {
#pragma unroll
            for (int mb = 0; mb < MB; ++mb) {
                if (a.out_stats != nullptr) {
                    double* st = a.out_stats +
                                 2 * ((size_t)s_done * a.out_stats_stride + a.out_coff + (part_done * MB + mb) * 128 + q4 * 32 + lane);
                    atomicAdd(st, acc_su[mb]);
                    atomicAdd(st + 1, acc_ss[mb]);
                }
                acc_su[mb] = 0.0;
                acc_ss[mb] = 0.0;
            }
        };
        const int ncb = (ncol + 31) >> 5;
        for (int it = eg; it < ntiles; it += 2) {
            const XTile c = coord(tile_begin + it);
            if (c.s != cur_s || c.part != cur_part) {
                if (cur_s >= 0) flush(cur_s, cur_part);
                cur_s = c.s;
                cur_part = c.part;
            }
            XWAIT_SLEEP(&t_full[eg], (it >> 1) & 1, 64, 6, it);
            tc_fence_after();
#pragma unroll
            for (int mb = 0; mb < MB; ++mb) {
                float* obase = a.out + (size_t)c.s * hw_out * a.out_cstride + a.out_coff + (c.part * MB + mb) * 128 + q4 * 32 + lane;
                float su = 0.f, sq = 0.f;
                int py = 0, px = 0;          // tile coordinates of the column, advanced incrementally
                for (int cb = 0; cb < ncb; ++cb) {
                    float v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)((eg * MB + mb) * 128 + cb * 32), v);
                    if (mb == MB - 1 && cb == ncb - 1) {
                        tc_fence_before();
                        mbar_arrive(&t_empty[eg]);   // the accumulator is in registers: the MMA warp may overwrite it
                    }
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int oy = c.oy0 + py, ox = c.ox0 + px;
                        if (cb * 32 + i < ncol && oy < hout && ox < hout) {   // warp-uniform
                            obase[(size_t)(oy * hout + ox) * a.out_cstride] = v[i];   // 32 lanes = 128 contiguous bytes of one pixel
                            su += v[i];
                            sq = fmaf(v[i], v[i], sq);
                        }
                        if (++px == g.tw) { px = 0; ++py; }
                    }
                }
                acc_su[mb] += (double)su;
                acc_ss[mb] += (double)sq;
            }
        }
        if (cur_s >= 0) flush(cur_s, cur_part);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// pooled tile shape: tw x 4 sr pooled pixels (N = 4 tw sr, a multiple of 16, <= 128) with the fewest MMA columns over the map
void choose_tile(int hout, int* tw, int* sr) {
    if (const char* e = getenv("SMG_TRANS_TILE")) {   // A/B aid: "tw,sr"
        if (sscanf(e, "%d,%d", tw, sr) == 2 && *tw >= 1 && *sr >= 1 && 4 * *tw * *sr <= 128 && (4 * *tw * *sr) % 16 == 0) return;
    }
    static const int cand[][2] = {{16, 2}, {8, 4}, {32, 1}, {4, 8}, {20, 1}, {24, 1}, {28, 1}, {12, 2}, {2, 16}};
    long best = -1;
    for (const auto& cd : cand) {
        const int n = 4 * cd[0] * cd[1];
        const int th = 4 * cd[1];
        const long tiles = (long)((hout + cd[0] - 1) / cd[0]) * ((hout + th - 1) / th);
        const long cost = tiles * (n + 32 + (n < 128 ? 16 : 0));   // per-tile overhead; smaller TMA boxes keep fewer bytes in flight
        if (best < 0 || cost < best) { best = cost; *tw = cd[0]; *sr = cd[1]; }
    }
}

}  // namespace

// Returns SMG_ERR_UNSUPPORTED for shapes this kernel does not serve (the caller then uses conv_umma.cu's pooled instance).
int launch_trans_t(smg_handle* h, const ConvArgs& a, cudaStream_t st) {
    if (a.taps != 1 || !a.pool || a.cout % 128 != 0 || a.cin % KC != 0 || a.cin > 1024 || a.in_cstride % 4 != 0 || (a.hin & 1) ||
        a.w == nullptr || a.w->w_tf32 == nullptr || (reinterpret_cast<uintptr_t>(a.in) & 15) != 0)
        return SMG_ERR_UNSUPPORTED;
    const int hout = a.hin / 2;
    XGeom g;
    choose_tile(hout, &g.tw, &g.sr);
    g.tiles_x = (hout + g.tw - 1) / g.tw;
    g.tiles_y = (hout + 4 * g.sr - 1) / (4 * g.sr);
    const int mb = a.cout % 256 == 0 ? 2 : 1;
    g.parts = a.cout / (128 * mb);
    g.kg_total = a.cin / KC;
    CUtensorMap tm;
    const cuuint64_t dims[4] = {(cuuint64_t)a.in_cstride, (cuuint64_t)a.hin, (cuuint64_t)a.hin, (cuuint64_t)a.n};
    const cuuint64_t strides[3] = {(cuuint64_t)a.in_cstride * 4, (cuuint64_t)a.hin * a.in_cstride * 4,
                                   (cuuint64_t)a.hin * a.hin * a.in_cstride * 4};
    const cuuint32_t box[4] = {KC, (cuuint32_t)(2 * g.tw), (cuuint32_t)(2 * g.sr), 1};
    SMG_TRY(make_tensor_map_f32(&tm, a.in, 4, dims, strides, box, 128));
    UmmaDev d;
    d.in = a.in; d.in_cstride = a.in_cstride; d.cin = a.cin; d.hin = a.hin;
    d.prologue_mode = a.prologue_mode; d.in_stats = a.in_stats; d.in_stats_stride = a.in_stats_stride;
    d.gamma = a.gamma; d.beta = a.beta; d.scale = a.scale; d.shift = a.shift; d.relu = a.relu;
    d.w = a.w->w_tf32;
    d.out = a.out; d.out_cstride = a.out_cstride; d.out_coff = a.out_coff; d.cout = a.cout;
    d.out_stats = a.out_stats; d.out_stats_stride = a.out_stats_stride;
    d.hout = hout;
    d.wp = d.ht = d.tiles_x = 0;
    d.async_producer = 0;
    d.tiles_per_sample = g.tiles_x * g.tiles_y * g.parts;
    d.tiles_per_cta = 0;
    const int total = d.tiles_per_sample * a.n;
    const int grid = total < h->num_sms ? total : h->num_sms;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(X_THREADS);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = h->use_pdl ? 1 : 0;
    if (mb == 1) {
        SMG_TRY(ensure_dyn_smem(h, (const void*)trans_t_kernel<1>, XP<1>::TOTAL));
        cfg.dynamicSmemBytes = XP<1>::TOTAL;
        SMG_CUDA(cudaLaunchKernelEx(&cfg, trans_t_kernel<1>, tm, d, g, total));
    } else {
        SMG_TRY(ensure_dyn_smem(h, (const void*)trans_t_kernel<2>, XP<2>::TOTAL));
        cfg.dynamicSmemBytes = XP<2>::TOTAL;
        SMG_CUDA(cudaLaunchKernelEx(&cfg, trans_t_kernel<2>, tm, d, g, total));
    }
    h->launches++;
    SMG_CUDA(cudaGetLastError());
    return SMG_OK;
}

}  // namespace smg
