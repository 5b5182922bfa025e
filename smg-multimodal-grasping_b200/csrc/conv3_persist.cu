// conv3_persist.cu - persistent tf32 3x3 convolution (128 -> 32 channels) with the weights resident in shared memory.
//
// Serves torchvision densenet `_DenseLayer.conv2` (58 of the 120 trunk convolutions; /root/reference/code/models.py:319
// builds the trunks) with the contract of conv_umma.cu: BN-ReLU prologue, raw NHWC output slice, (sum, sumsq) epilogue.
//
// Why another kernel: with one tile per CTA the 3x3 layers are a latency chain (ncu, profiles/README.md): a tile waits
// for its patch, then for 144 MMAs, then for its epilogue, and the 144 KB of packed weights are re-fetched from L2 for
// every 128-pixel tile - as many bytes as the activations.  Here
//   * ONE CTA per SM owns the whole shared memory: the 36 weight stage images (144 KB) are loaded ONCE per launch;
//   * the CTA walks a contiguous range of the flattened (sample, tile) list; six 13.5 KB patch slots (16 channels each:
//     rows of 64 B, 64-byte swizzle) are refilled by 4-D tensor-map TMA boxes (channels, x, y, sample; halo zero-filled
//     by the copy engine) as soon as the MMAs that read them retire, independently of tile boundaries.  Six small slots
//     instead of three 32-channel ones: same bytes, but a slot's load -> transform -> MMA -> refill turn-around is what
//     bounds the kernel, and it shrinks with the slot;
//   * the three dx taps of a kernel row are ONE MMA: B = [W(dy,0) | W(dy,1) | W(dy,2)] is an N = 96 operand
//     (pack.cu: w_tf32_dx), so that E_dx[m] = sum_dy A[m + dy*wp] W(dy,dx) costs 3 shifted A reads per K step instead
//     of 9 - measured (profiles/r01_umma_rate.csv) an M=128 MMA costs 45 cycles at N = 32 and 48 at N = 64, i.e. it is
//     bound by the 4 KB A-operand read, not by its MACs.  out[m] = E_0[m] + E_1[m+1] + E_2[m+2] is formed by the
//     epilogue with two warp shuffles per channel (rows 30-31 of a warp take their neighbours from the next warp
//     through a 1 KB exchange buffer); a valid output row never needs a row beyond its own patch row;
//   * the accumulator (96 columns) is double buffered in TMEM; the 4 epilogue warps drain tile i (TMEM -> registers ->
//     global, one 128-byte row per thread) while the MMA warp is already working on tile i+1;
//   * output statistics are reduced with warp shuffles into per-warp double registers and flushed to HBM once per
//     (warp, sample) instead of once per tile.
// Warp roles: 0-3 and 10-13 in-place BN-ReLU transform of the landed patch, 4-7 epilogue, 8 MMA issuer, 9 TMA loader.
#include "tma_common.cuh"

namespace smg {

namespace {

// Two instantiations: PKC = 32 (three 27 KB patch slots, rows of 128 B, 128-byte swizzle) and PKC = 16 (six 13.5 KB slots,
// rows of 64 B, 64-byte swizzle).  Same bytes in flight; the small-slot variant hands patches over at a finer grain.
template <int PKC>
struct P3 {
    static constexpr int KCH = PKC;                       // channels per patch load
    static constexpr int ROWB = PKC * 4;                  // bytes per patch row
    static constexpr int NG = 128 / PKC;                  // patch loads per tile
    static constexpr int NSLOT = PKC == 32 ? 3 : 6;
    static constexpr int ROWS = 212;                      // patch rows touched: the TMA box fills (ht+2)*wp <= 210, the MMAs read up to 2*wp + 127
    static constexpr int SLOT = 27 * 8 * ROWB;            // slot stride: 216 rows, a multiple of the swizzle period (8 rows)
    static constexpr int LAST = ROWS * ROWB;              // bytes of a slot actually touched
    static constexpr int WSTAGE = (PKC / 4) * 96 * 16;    // one (channel group, kernel row) weight image: chunks x 96 rows x 16 B
    static constexpr int WBYTES = NG * 3 * WSTAGE;        // 144 KB
    static constexpr int OFF_A = 0;
    static constexpr int OFF_BAR = LAST;                  // the barriers + TMEM pointer live in the unused tail of slot 0
    static constexpr int OFF_W = OFF_A + (NSLOT - 1) * SLOT + LAST;
    static constexpr int OFF_SC = OFF_W + WBYTES;         // scale[128], shift[128]
    static constexpr int OFF_XCH = OFF_SC + 1024;         // rows 0-1 of epilogue warps 1-3: [3][E_1 row 0 | E_2 row 0 | E_2 row 1][32]
    static constexpr int TOTAL = OFF_XCH + 3 * 96 * 4;
    static_assert(TOTAL <= 232448, "shared-memory plan exceeds the 227 KB of one SM");
    static_assert(SLOT - LAST >= 24 * 8, "barriers do not fit the slot tail");
    static_assert(OFF_W % 128 == 0 && OFF_BAR % 8 == 0 && OFF_SC % 16 == 0, "alignment");
};
constexpr int P_NCOL = 96;                        // accumulator columns: dx * 32 + cout

struct TileCoord {
    int s, h0, w0;
};

template <int PKC>
__global__ void __launch_bounds__(448, 1)
conv3_persist_kernel(const __grid_constant__ CUtensorMap tmA, UmmaDev a, int total_tiles, long long* __restrict__ trace) {
    using P = P3<PKC>;
    constexpr int NS = P::NSLOT;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + P::OFF_BAR);
    uint64_t* raw_full = bars;          // [NS <= 6] patch landed (raw)
    uint64_t* a_ready = bars + 6;       // [NS] patch normalised (256 transform threads)
    uint64_t* a_empty = bars + 12;      // [NS] MMAs reading the slot retired
    uint64_t* tmem_full = bars + 18;    // [2]
    uint64_t* tmem_empty = bars + 20;   // [2] 128 epilogue threads
    uint64_t* w_full = bars + 22;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 23);
    float* s_sc = reinterpret_cast<float*>(smem + P::OFF_SC);
    float* s_sh = s_sc + 128;
    float* s_xch = reinterpret_cast<float*>(smem + P::OFF_XCH);
    uint8_t* sA = smem + P::OFF_A;
    uint8_t* sW = smem + P::OFF_W;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int hout = a.hout, hin = a.hin;
    const int hw_out = hout * hout;
    const int wp = a.wp;
    const int pfill = (a.ht + 2) * wp;
    const int tps = a.tiles_per_sample;
    // contiguous, balanced share of the flattened (sample, tile) list
    const int tile_begin = (int)(((long long)blockIdx.x * total_tiles) / gridDim.x);
    const int tile_end = (int)(((long long)(blockIdx.x + 1) * total_tiles) / gridDim.x);
    const int ntiles = tile_end - tile_begin;
    // optional timeline of CTA 0 (SMG_CONV3_TRACE): trace[event][index], clock64 stamps of the first 64 slot uses / tiles
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

    if (warp == 8 && lane == 0) {
        if (smem_u32(smem) & 1023u) __trap();   // the swizzled slots rely on a 1024-byte aligned window
        for (int i = 0; i < NS; ++i) { mbar_init(&raw_full[i], 1); mbar_init(&a_ready[i], 256); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 128); }
        mbar_init(w_full, 1);
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc(tmem_ptr, 256);   // 2 x 96 accumulator columns (power-of-two allocation)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 9) {
        // =============================== loader ===============================
        if (lane == 0) {
            mbar_arrive_expect_tx(w_full, P::WBYTES);
            for (int i = 0; i < 12; ++i) tma_bulk_load(sW + i * 12288, a.w + (size_t)i * 12288, 12288, w_full);   // 12 x 12 KB = 144 KB
            const int pd = a.l2_prefetch;     // L2 prefetch distance in patch loads (0 = off)
            for (int n = 0; n < pd && n < P::NG * ntiles; ++n) {
                const TileCoord c = coord(tile_begin + n / P::NG);
                tma_prefetch_4d(&tmA, (n % P::NG) * P::KCH, c.w0 - 1, c.h0 - 1, c.s);
            }
            for (int n = 0; n < P::NG * ntiles; ++n) {
                const int slot = n % NS;
                if (pd > 0 && n + pd < P::NG * ntiles) {
                    // run ahead of the shared-memory slots: HBM -> L2 for the patch that will be loaded pd loads later
                    const TileCoord cp = coord(tile_begin + (n + pd) / P::NG);
                    tma_prefetch_4d(&tmA, ((n + pd) % P::NG) * P::KCH, cp.w0 - 1, cp.h0 - 1, cp.s);
                }
                const TileCoord c = coord(tile_begin + n / P::NG);
                mbar_wait_sleep(&a_empty[slot], ((n / NS) & 1) ^ 1, 64);
                stamp(0, n);   // slot free, load issued
                mbar_arrive_expect_tx(&raw_full[slot], (uint32_t)(pfill * P::ROWB));
                tma_tile_4d(sA + slot * P::SLOT, &tmA, (n % P::NG) * P::KCH, c.w0 - 1, c.h0 - 1, c.s, &raw_full[slot]);
            }
        }
    } else if (warp < 4 || warp >= 10) {
        // =============================== in-place transform ===============================
        const int ptid = warp < 4 ? tid : tid - 192;          // 0..255
        constexpr int PPR = PKC / 4;                          // 16-byte pieces per patch row (8 or 4)
        constexpr int RSTEP = 256 / PPR;                      // rows covered by the 256 threads per pass (32 or 64)
        constexpr int NI = (P::ROWS + RSTEP - 1) / RSTEP;     // passes (7 or 4)
        const int j = ptid % PPR;                             // physical 16-byte piece of the row
        const int rbase = ptid / PPR;                         // patch rows rbase + RSTEP i
        // logical 4-channel chunk held by that piece: 128-byte swizzle xors with (row & 7), 64-byte swizzle with (row >> 1) & 3
        const int chunk = PKC == 32 ? (j ^ (rbase & 7)) : (j ^ ((rbase >> 1) & 3));
        int cur_s = -1;
        uint32_t inside = 0, filled = 0;
        for (int n = 0; n < P::NG * ntiles; ++n) {
            const int g = n % P::NG;
            const int slot = n % NS;
            if (g == 0) {
                const TileCoord c = coord(tile_begin + n / P::NG);
                if (c.s != cur_s) {
                    // BN scale/shift of the new sample (every transform thread has left the previous tile's tables)
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
                    const int q = rbase + RSTEP * i;
                    const int py = q / wp, px = q - py * wp;
                    const int y = c.h0 - 1 + py, x = c.w0 - 1 + px;
                    if (q < pfill) {
                        filled |= 1u << i;
                        if (y >= 0 && y < hin && x >= 0 && x < hin) inside |= 1u << i;
                    }
                }
            }
            const float4 sc = *reinterpret_cast<const float4*>(s_sc + g * P::KCH + chunk * 4);
            const float4 sh = *reinterpret_cast<const float4*>(s_sh + g * P::KCH + chunk * 4);
            mbar_wait_sleep(&raw_full[slot], (n / NS) & 1, 64);
            if (ptid == 0) stamp(1, n);   // patch landed
            uint8_t* base = sA + slot * P::SLOT + rbase * P::ROWB + j * 16;
            float4 x[NI];
#pragma unroll
            for (int i = 0; i < NI; ++i)
                if (filled & (1u << i)) x[i] = *reinterpret_cast<const float4*>(base + i * RSTEP * P::ROWB);
#pragma unroll
            for (int i = 0; i < NI; ++i) {
                if (!(filled & (1u << i))) continue;
                float4 y;
                y.x = fmaf(x[i].x, sc.x, sh.x); y.y = fmaf(x[i].y, sc.y, sh.y);
                y.z = fmaf(x[i].z, sc.z, sh.z); y.w = fmaf(x[i].w, sc.w, sh.w);
                if (a.relu) { y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f); }
                if (!(inside & (1u << i))) y = make_float4(0.f, 0.f, 0.f, 0.f);   // conv zero padding is post-activation
                *reinterpret_cast<float4*>(base + i * RSTEP * P::ROWB) = y;
            }
            fence_proxy_async();
            mbar_arrive(&a_ready[slot]);
            if (ptid == 0) stamp(2, n);   // this thread's part of the patch normalised
        }
    } else if (warp == 8) {
        // =============================== MMA issuer ===============================
        if (lane == 0) {
            constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(P_NCOL >> 3) << 17) |
                                       ((uint32_t)(UM >> 4) << 24);
            const uint32_t sA_u = smem_u32(sA), sW_u = smem_u32(sW);
            mbar_wait(w_full, 0);
            for (int it = 0; it < ntiles; ++it) {
                const int buf = it & 1;
                mbar_wait(&tmem_empty[buf], ((it >> 1) & 1) ^ 1);   // the epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(buf * 128);
                uint32_t accum = 0;
                for (int g = 0; g < P::NG; ++g) {
                    const int n = it * P::NG + g;
                    const int slot = n % NS;
                    mbar_wait(&a_ready[slot], (n / NS) & 1);
                    tc_fence_after();
                    stamp(3, n);   // MMA issue starts
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy) {
                        // kernel row dy = a shift of dy*wp patch rows = +64 B per row on the start address; the swizzle
                        // phase follows the absolute address, so descriptor base_offset stays 0
                        const uint32_t start = sA_u + slot * P::SLOT + dy * wp * P::ROWB;
                        const uint32_t wst = sW_u + (g * 3 + dy) * P::WSTAGE;
#pragma unroll
                        for (int k = 0; k < P::KCH / 8; ++k) {
                            const uint64_t ad = PKC == 32 ? make_desc_sw128(start + k * 32) : make_desc_sw64(start + k * 32);
                            const uint64_t bd = make_desc(wst + 2 * k * P_NCOL * 16, P_NCOL * 16, 128);
                            umma<4>(d_tmem, ad, bd, idesc, accum);
                            accum = 1;
                        }
                    }
                    umma_commit(&a_empty[slot]);
                    stamp(4, n);   // MMAs of the slot issued
                }
                umma_commit(&tmem_full[buf]);
            }
        }
    } else {
        // =============================== epilogue (warps 4-7) ===============================
        const int e = warp - 4;              // TMEM lane partition of this warp
        const int row = e * 32 + lane;       // accumulator row == tile row
        const int ri = row / wp, rj = row - ri * wp;
        // per-warp statistics of the rows this warp drained (lane = channel), flushed when the sample changes
        double acc_su = 0.0, acc_ss = 0.0;
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
            const int buf = it & 1;
            if (c.s != cur_s) {
                if (cur_s >= 0) flush(cur_s);
                cur_s = c.s;
            }
            const bool valid = ri < a.ht && rj < wp - 2 && c.h0 + ri < hout && c.w0 + rj < hout;
            mbar_wait_sleep(&tmem_full[buf], (it >> 1) & 1, 128);
            tc_fence_after();
            if (e == 0 && lane == 0) stamp(5, it);   // accumulator complete
            float v[32], e1[32], e2[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(e * 32) << 16) + (uint32_t)(buf * 128);
            tmem_ld32(taddr, v);
            tmem_ld32(taddr + 32, e1);
            tmem_ld32(taddr + 64, e2);
            tc_fence_before();
            mbar_arrive(&tmem_empty[buf]);
            // out[m] = E_0[m] + E_1[m+1] + E_2[m+2]: neighbours inside the warp by shuffle, the first two rows of the next
            // warp through shared memory (a VALID output row m never needs a row beyond the tile: m + 2 stays in its patch row)
            if (e > 0 && lane < 2) {
                float* x = s_xch + (e - 1) * 96;
                if (lane == 0) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) x[i] = e1[i];
                }
#pragma unroll
                for (int i = 0; i < 32; ++i) x[32 + lane * 32 + i] = e2[i];
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            {
                // branch-free: every lane reads the exchange rows (two broadcast addresses), lanes 30/31 keep them
                const float* x = s_xch + (e < 3 ? e : 0) * 96;   // warp 3's rows 126-127 are never valid outputs
                const float* x2 = x + 32 + (lane & 1) * 32;      // lane 30 -> next warp's row 0, lane 31 -> its row 1
                const bool last1 = lane == 31, last2 = lane >= 30;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float s1 = __shfl_down_sync(0xffffffffu, e1[i], 1);
                    const float s2 = __shfl_down_sync(0xffffffffu, e2[i], 2);
                    const float xa = x[i], xb = x2[i];
                    v[i] = (v[i] + (last1 ? xa : s1)) + (last2 ? xb : s2);
                }
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");   // the exchange buffer may be rewritten for the next tile
            if (valid) {
                float4* o = reinterpret_cast<float4*>(a.out + ((size_t)c.s * hw_out + (c.h0 + ri) * hout + c.w0 + rj) * a.out_cstride +
                                                      a.out_coff);
#pragma unroll
                for (int q = 0; q < 8; ++q) o[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            }
            if (a.out_stats != nullptr) {
                float sq[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    if (!valid) v[i] = 0.f;
                    sq[i] = v[i] * v[i];
                }
                acc_su += (double)warp_transpose_sum(v, lane);     // lane = channel
                acc_ss += (double)warp_transpose_sum(sq, lane);
            }
            if (e == 0 && lane == 0) stamp(6, it);   // tile drained
        }
        if (cur_s >= 0) flush(cur_s);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

}  // namespace

// Returns SMG_ERR_UNSUPPORTED for shapes this kernel does not serve (the caller then uses the one-tile kernels).
int launch_conv3_persist(smg_handle* h, const ConvArgs& a, cudaStream_t st) {
    if (a.taps != 9 || a.pool || a.cin != 128 || a.cout != 32 || a.in_cstride % 4 != 0 || a.out_cstride % 4 != 0 ||
        a.out_coff % 4 != 0 || (reinterpret_cast<uintptr_t>(a.in) & 15) != 0 || (reinterpret_cast<uintptr_t>(a.out) & 15) != 0)
        return SMG_ERR_UNSUPPORTED;
    SMG_CHECK(a.w != nullptr && a.w->w_tf32_dx != nullptr && a.w->w_tf32_dx32 != nullptr, SMG_ERR_STATE, "conv3_persist: weights not packed");
    UmmaDev d;
    d.in = a.in; d.in_cstride = a.in_cstride; d.cin = a.cin; d.hin = a.hin;
    d.prologue_mode = a.prologue_mode; d.in_stats = a.in_stats; d.in_stats_stride = a.in_stats_stride;
    d.gamma = a.gamma; d.beta = a.beta; d.scale = a.scale; d.shift = a.shift; d.relu = a.relu;
    d.w = h->conv3_slot_channels == 16 ? a.w->w_tf32_dx : a.w->w_tf32_dx32;
    d.out = a.out; d.out_cstride = a.out_cstride; d.out_coff = a.out_coff; d.cout = a.cout;
    d.out_stats = a.out_stats; d.out_stats_stride = a.out_stats_stride;
    d.hout = a.hin;
    umma_patch_geometry(d.hout, &d.wp, &d.ht);
    // the last kernel row reads patch rows up to 2*wp + 127; the TMA box fills (ht+2)*wp rows
    SMG_CHECK(2 * d.wp + UM <= P3<32>::ROWS && (d.ht + 2) * d.wp <= P3<32>::ROWS && d.ht * d.wp <= UM, SMG_ERR_STATE,
              "conv3_persist: patch %dx%d too large", d.ht, d.wp);
    const int wt = d.wp - 2;
    d.tiles_x = (d.hout + wt - 1) / wt;
    d.tiles_per_sample = d.tiles_x * ((d.hout + d.ht - 1) / d.ht);
    d.tiles_per_cta = 0;
    d.l2_prefetch = h->l2_prefetch;
    d.async_producer = 0;
    const int total = d.tiles_per_sample * a.n;

    CUtensorMap tm;
    const cuuint64_t dims[4] = {(cuuint64_t)a.in_cstride, (cuuint64_t)a.hin, (cuuint64_t)a.hin, (cuuint64_t)a.n};
    const cuuint64_t strides[3] = {(cuuint64_t)a.in_cstride * 4, (cuuint64_t)a.hin * a.in_cstride * 4,
                                   (cuuint64_t)a.hin * a.hin * a.in_cstride * 4};
    const bool small = h->conv3_slot_channels == 16;
    const cuuint32_t box[4] = {(cuuint32_t)(small ? 16 : 32), (cuuint32_t)d.wp, (cuuint32_t)(d.ht + 2), 1};
    SMG_TRY(make_tensor_map_f32(&tm, a.in, 4, dims, strides, box, small ? 64 : 128));
    static bool attr = false;
    if (!attr) {
        SMG_CUDA(cudaFuncSetAttribute(conv3_persist_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, P3<32>::TOTAL));
        SMG_CUDA(cudaFuncSetAttribute(conv3_persist_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, P3<16>::TOTAL));
        attr = true;
    }
    const int grid = total < h->num_sms ? total : h->num_sms;
    // debugging aid: SMG_CONV3_TRACE=<file> appends the clock64 timeline of CTA 0 of every launch (synchronises!)
    static const char* trace_path = getenv("SMG_CONV3_TRACE");
    static long long* trace_dev = nullptr;
    if (trace_path != nullptr && trace_dev == nullptr) SMG_CUDA(cudaMalloc(&trace_dev, 7 * 64 * sizeof(long long)));
    if (trace_dev != nullptr) SMG_CUDA(cudaMemsetAsync(trace_dev, 0, 7 * 64 * sizeof(long long), st));
    if (small) conv3_persist_kernel<16><<<grid, 448, P3<16>::TOTAL, st>>>(tm, d, total, trace_dev);
    else conv3_persist_kernel<32><<<grid, 448, P3<32>::TOTAL, st>>>(tm, d, total, trace_dev);
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
