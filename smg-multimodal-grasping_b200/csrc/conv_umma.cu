// conv_umma.cu - tensor-core implicit-GEMM convolution for sm_100a (tcgen05 + TMEM + bulk TMA).
//
// Same contract as conv_ffma.cu (BN-ReLU prologue, optional 2x2 avg-pool prologue, 1x1 or
// 3x3 taps, raw output + (sum,sumsq) epilogue) with operands rounded to tf32 or bf16 and
// fp32 accumulation in tensor memory.  Serves torchvision densenet `_DenseLayer.conv1/conv2`,
// `_Transition.conv` and the head's 1x1 conv (/root/reference/code/models.py:319,384-387).
//
// CTA = one 128-pixel output tile x BN output channels, 14 warps:
//   warps 0-3 and 10-13  two groups of A producers (each fills every other stage): global (NHWC fp32, pre-BN) -> registers -> relu(x*scale+shift) ->
//              tf32/bf16 -> shared memory in the UMMA no-swizzle K-major layout
//              [16-byte K chunk][row][16 B]  (core matrix = 8 rows x 16 B contiguous, SBO = 128 B,
//              LBO = padded row count x 16 B).  A row shift of s pixels is a start-address
//              offset of 16*s bytes, which is what makes the 3x3 taps free: the activation patch
//              (tile + halo) is transformed ONCE and the 9 taps are 9 shifted descriptors.
//   warps 4-7  epilogue: tcgen05.ld accumulator -> shared staging -> coalesced NHWC stores into the
//              layer's channel slice + per-channel (sum, sumsq) -> double atomics.
//   warp 8     MMA issuer (one lane): tcgen05.mma kind::tf32 / kind::f16, M=128, N=BN, D in TMEM;
//              tcgen05.commit releases shared-memory stages and finally signals the epilogue.
//   warp 9     weight loader (one lane): cp.async.bulk (TMA bulk copy, mbarrier complete_tx) of
//              pre-packed weight stage images (pack.cu writes them in the exact smem layout).
#include "umma_common.cuh"

namespace smg {

// SPLIT = 1 (tf32 elements only): fp32-accurate "3xTF32" mode.  Every operand is kept as hi + lo (hi = the value with its 13
// low mantissa bits cleared, exactly a tf32 number; lo = value - hi, exact in fp32) and D += a_lo b_hi + a_hi b_lo + a_hi b_hi
// (the dropped a_lo b_lo term is 2^-22 relative): each A slot / B stage holds the hi image followed by the lo image.
// TAPS = 1: 1x1; 9: 3x3 as nine shifted taps of N = BN output channels; 3: 3x3 with the three dx taps MERGED into N (BN = 3 x cout
// columns = (dx, co)): three row-shifted taps (dy) of wider MMAs, the dx shift is applied when the staged accumulator is read back
// (out[m][co] = E[m][co] + E[m+1][cout+co] + E[m+2][2 cout+co]).  Used by the fp32 (split) mode, whose N = 32 MMAs are issue-bound.
template <int ELT, int BN, int TAPS, int SPLIT = 0>
struct SmemPlan {
    using E = EltCfg<ELT>;
    static constexpr int A_LBO = (TAPS != 1 ? E::P_ROWS : E::A_ROWS) * 16;
    static constexpr int A_HALF = E::CH * A_LBO;                      // one 32-channel group, one image
    static constexpr int A_SLOT = (1 + SPLIT) * A_HALF;
    static constexpr int A_SLOTS = TAPS != 1 ? 2 : NA;
    static constexpr int B_HALF = E::CH * BN * 16;
    static constexpr int B_STAGE = (1 + SPLIT) * B_HALF;
    static constexpr int B_SLOTS = TAPS == 9 ? NB9 : (TAPS == 3 ? 3 : NB1);   // TAPS == 3: 24 KB stages (96 rows, hi + lo)
    static constexpr int OFF_BAR = 0;
    static constexpr int OFF_SC = 256;
    static constexpr int OFF_A = OFF_SC + 2 * 1024 * 4;
    static constexpr int OFF_B = OFF_A + A_SLOTS * A_SLOT;
    static constexpr int STAGING = UM * (BN + 1) * 4;
    static constexpr int END_AB = OFF_B + B_SLOTS * B_STAGE;
    static constexpr int TOTAL = (OFF_A + STAGING > END_AB ? OFF_A + STAGING : END_AB);
};


// fp32 bits rounded to the nearest tf32 value (10 explicit mantissa bits): the hi part of a split operand; |x - hi| <= 2^-12 |x|
__device__ __forceinline__ uint32_t tf32_rn(uint32_t bits) { return (bits + 0x1000u) & 0xFFFFE000u; }

template <int ELT, int BN, int TAPS, int POOL, int SPLIT = 0>
__global__ void __launch_bounds__(448, 2)
conv_umma_kernel(UmmaDev a) {
    static_assert(SPLIT == 0 || ELT == 4, "the hi/lo split is a tf32 mode");
    using E = EltCfg<ELT>;
    using P = SmemPlan<ELT, BN, TAPS, SPLIT>;
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + P::OFF_BAR);
    uint64_t* a_full = bars;            // [4]
    uint64_t* a_empty = bars + 4;       // [4]
    uint64_t* b_full = bars + 8;        // [8]
    uint64_t* b_empty = bars + 16;      // [8]
    uint64_t* tmem_full = bars + 24;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 25);
    float* s_sc = reinterpret_cast<float*>(smem + P::OFF_SC);
    float* s_sh = s_sc + 1024;
    uint8_t* sA = smem + P::OFF_A;
    uint8_t* sB = smem + P::OFF_B;
    float* s_out = reinterpret_cast<float*>(smem + P::OFF_A);  // staging aliases the operand stages

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int s = blockIdx.z;
    const int ntile = blockIdx.y;
    const int hout = a.hout, hin = a.hin;
    const int hw_out = hout * hout;
    const int KG = a.cin / KC;

    constexpr int NOUT = TAPS == 3 ? BN / 3 : BN;      // output channels of the tile
    constexpr int TMC = BN == 96 ? 128 : BN;           // tensor-memory columns (a power of two)
    // tile origin
    int m0 = 0, h0 = 0, w0 = 0;
    if (TAPS != 1) {
        const int ty = blockIdx.x / a.tiles_x, tx = blockIdx.x - ty * a.tiles_x;
        h0 = ty * a.ht;
        w0 = tx * (a.wp - 2);
    } else {
        m0 = blockIdx.x * UM;
    }

    // ---- one-time setup
    if (warp == 8 && lane == 0) {
        for (int i = 0; i < 4; ++i) { mbar_init(&a_full[i], 128); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < 8; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc(tmem_ptr, TMC);
    // Programmatic dependent launch (the training step chains ~240 of these kernels with the BatchNorm-backward kernels,
    // many of them on 7-50 CTAs): the successor may become resident now; everything below reads what the predecessor wrote
    pdl_launch_dependents();
    pdl_wait();
    // BN prologue parameters of this sample
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
    } else if (a.prologue_mode == 1) {
        for (int c = tid; c < a.cin; c += 448) {
            s_sc[c] = a.scale[(size_t)s * a.cin + c];
            s_sh[c] = a.shift[(size_t)s * a.cin + c];
        }
    } else {   // identity prologue (data-gradient convolutions): fmaf(x, 1, 0) == x
        for (int c = tid; c < a.cin; c += 448) {
            s_sc[c] = 1.f;
            s_sh[c] = 0.f;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    const float* inp = a.in + (size_t)s * hin * hin * a.in_cstride;

    if (warp < 4 || warp >= 10) {
        // =============================== A producers ===============================
        // two groups of 4 warps; group g fills every other K stage (1x1) / patch slot g (3x3), so two stages
        // worth of global loads are in flight per CTA without growing the per-thread register footprint
        const int pgroup = warp < 4 ? 0 : 1;
        const int ptid = tid & 127;               // thread index inside the group (warps 10-13 start at 320)
        const int c = ptid % E::CH;               // chunk within the 32-channel group
        const int r0 = ptid / E::CH;              // first row handled by this thread
        constexpr int RSTEP = 128 / E::CH;    // row stride between iterations
        if (TAPS == 1) {
            constexpr int RI = UM / RSTEP;    // rows per thread per stage
            // per-row input offsets (in floats, relative to inp), -1 = row beyond the sample
            int roff[RI];
#pragma unroll
            for (int i = 0; i < RI; ++i) {
                const int m = m0 + r0 + i * RSTEP;
                if (m < hw_out) {
                    if (POOL) {
                        const int oy = m / hout, ox = m - oy * hout;
                        roff[i] = ((2 * oy) * hin + 2 * ox) * a.in_cstride;
                    } else {
                        roff[i] = m * a.in_cstride;
                    }
                } else {
                    roff[i] = -1;
                }
            }
            if (ELT == 4 && !POOL && !SPLIT && a.async_producer) {
                // tf32 small-grid path: raw fp32 rows are cp.async'ed straight into their final UMMA slot (NA stages in
                // flight per thread, no registers held), then normalised IN PLACE by the thread that loaded them.
                auto issue = [&](int kg) {
                    uint8_t* dst = sA + (kg % NA) * P::A_SLOT + c * P::A_LBO;
                    const float* src = inp + kg * KC + c * E::EPC;
#pragma unroll
                    for (int i = 0; i < RI; ++i)
                        if (roff[i] >= 0) cp_async16(dst + (r0 + i * RSTEP) * 16, src + roff[i]);
                    cp_async_commit();
                };
                // each producer group owns every other K stage and keeps one of its own stages prefetched
                if (pgroup < KG) issue(pgroup);
                else cp_async_commit();
                for (int kg = pgroup; kg < KG; kg += 2) {
                    const int slot = kg % NA;
                    const int next = kg + 2;
                    if (next < KG) {
                        mbar_wait(&a_empty[next % NA], ((next / NA) & 1) ^ 1);
                        issue(next);
                    } else {
                        cp_async_commit();
                    }
                    cp_async_wait<1>();
                    const int ch0 = kg * KC + c * E::EPC;
                    const float4 sc = *reinterpret_cast<const float4*>(s_sc + ch0);
                    const float4 sh = *reinterpret_cast<const float4*>(s_sh + ch0);
                    uint8_t* dst = sA + slot * P::A_SLOT + c * P::A_LBO;
#pragma unroll
                    for (int i = 0; i < RI; ++i) {
                        float4* p4 = reinterpret_cast<float4*>(dst + (r0 + i * RSTEP) * 16);
                        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (roff[i] >= 0) {
                            x = *p4;
                            x.x = fmaf(x.x, sc.x, sh.x); x.y = fmaf(x.y, sc.y, sh.y);
                            x.z = fmaf(x.z, sc.z, sh.z); x.w = fmaf(x.w, sc.w, sh.w);
                            if (a.relu) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f); }
                        }
                        *p4 = x;
                    }
                    fence_proxy_async();
                    mbar_arrive(&a_full[slot]);
                }
            } else
            for (int kg = pgroup; kg < KG; kg += 2) {
                const int slot = kg % NA;
                const uint32_t ph = (kg / NA) & 1;
                mbar_wait(&a_empty[slot], ph ^ 1);
                const int ch0 = kg * KC + c * E::EPC;
                float sc[E::EPC], sh[E::EPC];
#pragma unroll
                for (int e = 0; e < E::EPC; ++e) { sc[e] = s_sc[ch0 + e]; sh[e] = s_sh[ch0 + e]; }
                uint8_t* dst = sA + slot * P::A_SLOT + c * P::A_LBO;
                float v[RI][E::EPC];
#pragma unroll
                for (int i = 0; i < RI; ++i) {
                    if (roff[i] >= 0) {
                        const float* src = inp + roff[i] + ch0;
                        if (POOL) {
#pragma unroll
                            for (int e = 0; e < E::EPC; ++e) v[i][e] = 0.f;
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const float* sq = src + ((q >> 1) * hin + (q & 1)) * a.in_cstride;
#pragma unroll
                                for (int e4 = 0; e4 < E::EPC / 4; ++e4) {
                                    const float4 x = __ldg(reinterpret_cast<const float4*>(sq) + e4);
                                    float t0 = fmaf(x.x, sc[4 * e4 + 0], sh[4 * e4 + 0]);
                                    float t1 = fmaf(x.y, sc[4 * e4 + 1], sh[4 * e4 + 1]);
                                    float t2 = fmaf(x.z, sc[4 * e4 + 2], sh[4 * e4 + 2]);
                                    float t3 = fmaf(x.w, sc[4 * e4 + 3], sh[4 * e4 + 3]);
                                    if (a.relu) { t0 = fmaxf(t0, 0.f); t1 = fmaxf(t1, 0.f); t2 = fmaxf(t2, 0.f); t3 = fmaxf(t3, 0.f); }
                                    v[i][4 * e4 + 0] += t0; v[i][4 * e4 + 1] += t1;
                                    v[i][4 * e4 + 2] += t2; v[i][4 * e4 + 3] += t3;
                                }
                            }
#pragma unroll
                            for (int e = 0; e < E::EPC; ++e) v[i][e] *= 0.25f;
                        } else {
#pragma unroll
                            for (int e4 = 0; e4 < E::EPC / 4; ++e4) {
                                const float4 x = __ldg(reinterpret_cast<const float4*>(src) + e4);
                                v[i][4 * e4 + 0] = x.x; v[i][4 * e4 + 1] = x.y;
                                v[i][4 * e4 + 2] = x.z; v[i][4 * e4 + 3] = x.w;
                            }
                        }
                    } else {
#pragma unroll
                        for (int e = 0; e < E::EPC; ++e) v[i][e] = 0.f;
                    }
                }
#pragma unroll
                for (int i = 0; i < RI; ++i) {
                    if (!POOL && roff[i] >= 0) {
#pragma unroll
                        for (int e = 0; e < E::EPC; ++e) {
                            float t = fmaf(v[i][e], sc[e], sh[e]);
                            v[i][e] = a.relu ? fmaxf(t, 0.f) : t;
                        }
                    }
                    uint4 pk;
                    if (ELT == 4) {
                        pk = make_uint4(__float_as_uint(v[i][0]), __float_as_uint(v[i][1]), __float_as_uint(v[i][2]),
                                        __float_as_uint(v[i][3]));
                        if (SPLIT) {
                            const uint4 hi = make_uint4(tf32_rn(pk.x), tf32_rn(pk.y), tf32_rn(pk.z), tf32_rn(pk.w));
                            const uint4 lo = make_uint4(__float_as_uint(v[i][0] - __uint_as_float(hi.x)), __float_as_uint(v[i][1] - __uint_as_float(hi.y)),
                                                        __float_as_uint(v[i][2] - __uint_as_float(hi.z)), __float_as_uint(v[i][3] - __uint_as_float(hi.w)));
                            *reinterpret_cast<uint4*>(dst + P::A_HALF + (r0 + i * RSTEP) * 16) = lo;
                            pk = hi;
                        }
                    } else {
                        pk = make_uint4(pack_bf16x2(v[i][0], v[i][1]), pack_bf16x2(v[i][2], v[i][3]),
                                        pack_bf16x2(v[i][4 % E::EPC], v[i][5 % E::EPC]),
                                        pack_bf16x2(v[i][6 % E::EPC], v[i][7 % E::EPC]));
                    }
                    *reinterpret_cast<uint4*>(dst + (r0 + i * RSTEP) * 16) = pk;
                }
                fence_proxy_async();
                mbar_arrive(&a_full[slot]);
            }
        } else {
            // 3x3: transform the (ht+2) x wp activation patch once per 32-channel group; the 9 taps are
            // 9 shifted descriptors over it.  Two patch slots: group g+2 re-uses the slot of group g
            // once the MMAs of group g have completed (a_empty).  All global loads of a group are
            // issued before the first use so that 14 x 16 B per thread are in flight.
            const int wp = a.wp;
            const int pfill = (a.ht + 2) * wp;
            constexpr int NI = (5 * MAX_WP + RSTEP - 1) / RSTEP;  // max rows per thread: (ht+2)*wp <= 5*42
            int poff[NI];
#pragma unroll
            for (int i = 0; i < NI; ++i) {
                const int q = r0 + i * RSTEP;
                const int py = q / wp, px = q - py * wp;
                const int y = h0 - 1 + py, x = w0 - 1 + px;
                poff[i] = (q < pfill && y >= 0 && y < hin && x >= 0 && x < hin) ? (y * hin + x) * a.in_cstride : -1;
            }
            if (ELT == 4 && !SPLIT && a.async_producer) {
                // tf32 small-grid path: both patch slots are filled with cp.async (raw fp32 lands in its final place,
                // 2 x 27 KB in flight per CTA) and normalised in place by the loading thread.
                auto issue = [&](int g) {
                    uint8_t* dst = sA + (g & 1) * P::A_SLOT + c * P::A_LBO;
                    const float* src = inp + g * KC + c * E::EPC;
#pragma unroll
                    for (int i = 0; i < NI; ++i)
                        if (poff[i] >= 0) cp_async16(dst + (r0 + i * RSTEP) * 16, src + poff[i]);
                    cp_async_commit();
                };
                if (pgroup < KG) issue(pgroup);  // producer group g owns patch slot g: channel groups g and g+2
                for (int g = pgroup; g < KG; g += 2) {
                    const int slot = g & 1;
                    cp_async_wait<0>();
                    const int ch0 = g * KC + c * E::EPC;
                    const float4 sc = *reinterpret_cast<const float4*>(s_sc + ch0);
                    const float4 sh = *reinterpret_cast<const float4*>(s_sh + ch0);
                    uint8_t* dst = sA + slot * P::A_SLOT + c * P::A_LBO;
#pragma unroll
                    for (int i = 0; i < NI; ++i) {
                        const int q = r0 + i * RSTEP;
                        if (q < pfill) {
                            float4* p4 = reinterpret_cast<float4*>(dst + q * 16);
                            float4 x = make_float4(0.f, 0.f, 0.f, 0.f);  // conv zero padding (post-activation)
                            if (poff[i] >= 0) {
                                x = *p4;
                                x.x = fmaf(x.x, sc.x, sh.x); x.y = fmaf(x.y, sc.y, sh.y);
                                x.z = fmaf(x.z, sc.z, sh.z); x.w = fmaf(x.w, sc.w, sh.w);
                                if (a.relu) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f); }
                            }
                            *p4 = x;
                        }
                    }
                    fence_proxy_async();
                    mbar_arrive(&a_full[slot]);
                    if (g + 2 < KG) {
                        mbar_wait(&a_empty[slot], 0);
                        issue(g + 2);
                    } else {
                        cp_async_commit();
                    }
                }
            } else
            for (int g = pgroup; g < KG; g += 2) {
                const int slot = g & 1;
                if (g >= 2) mbar_wait(&a_empty[slot], 0);
                const int ch0 = g * KC + c * E::EPC;
                float sc[E::EPC], sh[E::EPC];
#pragma unroll
                for (int e = 0; e < E::EPC; ++e) { sc[e] = s_sc[ch0 + e]; sh[e] = s_sh[ch0 + e]; }
                uint8_t* dst = sA + slot * P::A_SLOT + c * P::A_LBO;
                constexpr int NB = 4;                    // batches of loads (register budget: 72 at 2 x 448 threads / SM)
                constexpr int NH = (NI + NB - 1) / NB;
#pragma unroll
                for (int hb = 0; hb < NB; ++hb) {
                    float4 v[NH][E::EPC / 4];
#pragma unroll
                    for (int ii = 0; ii < NH; ++ii) {
                        const int i = hb * NH + ii;
#pragma unroll
                        for (int e4 = 0; e4 < E::EPC / 4; ++e4) {
                            v[ii][e4] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (i < NI && poff[i < NI ? i : 0] >= 0)
                                v[ii][e4] = __ldg(reinterpret_cast<const float4*>(inp + poff[i < NI ? i : 0] + ch0) + e4);
                        }
                    }
#pragma unroll
                    for (int ii = 0; ii < NH; ++ii) {
                        const int i = hb * NH + ii;
                        const int q = r0 + i * RSTEP;
                        if (i < NI && q < pfill) {
                            float t[E::EPC];
#pragma unroll
                            for (int e4 = 0; e4 < E::EPC / 4; ++e4) {
                                t[4 * e4 + 0] = v[ii][e4].x; t[4 * e4 + 1] = v[ii][e4].y;
                                t[4 * e4 + 2] = v[ii][e4].z; t[4 * e4 + 3] = v[ii][e4].w;
                            }
                            if (poff[i < NI ? i : 0] >= 0) {
#pragma unroll
                                for (int e = 0; e < E::EPC; ++e) {
                                    const float u = fmaf(t[e], sc[e], sh[e]);
                                    t[e] = a.relu ? fmaxf(u, 0.f) : u;
                                }
                            }  // else: conv zero padding (post-activation zeros)
                            uint4 pk;
                            if (ELT == 4) {
                                pk = make_uint4(__float_as_uint(t[0]), __float_as_uint(t[1]), __float_as_uint(t[2]),
                                                __float_as_uint(t[3]));
                                if (SPLIT) {
                                    const uint4 hi = make_uint4(tf32_rn(pk.x), tf32_rn(pk.y), tf32_rn(pk.z), tf32_rn(pk.w));
                                    const uint4 lo = make_uint4(__float_as_uint(t[0] - __uint_as_float(hi.x)), __float_as_uint(t[1] - __uint_as_float(hi.y)),
                                                                __float_as_uint(t[2] - __uint_as_float(hi.z)), __float_as_uint(t[3] - __uint_as_float(hi.w)));
                                    *reinterpret_cast<uint4*>(dst + P::A_HALF + q * 16) = lo;
                                    pk = hi;
                                }
                            } else {
                                pk = make_uint4(pack_bf16x2(t[0], t[1]), pack_bf16x2(t[2], t[3]),
                                                pack_bf16x2(t[4 % E::EPC], t[5 % E::EPC]),
                                                pack_bf16x2(t[6 % E::EPC], t[7 % E::EPC]));
                            }
                            *reinterpret_cast<uint4*>(dst + q * 16) = pk;
                        }
                    }
                }
                fence_proxy_async();
                mbar_arrive(&a_full[slot]);
            }
        }
    } else if (warp == 9) {
        // =============================== weight loader ===============================
        if (lane == 0) {
            const int nstages = TAPS * KG;
            const uint8_t* wsrc = a.w + (size_t)ntile * nstages * P::B_STAGE;
            for (int j = 0; j < nstages; ++j) {
                const int slot = j % P::B_SLOTS;
                const uint32_t ph = (j / P::B_SLOTS) & 1;
                mbar_wait(&b_empty[slot], ph ^ 1);
                mbar_arrive_expect_tx(&b_full[slot], P::B_STAGE);
                tma_bulk_load(sB + slot * P::B_STAGE, wsrc + (size_t)j * P::B_STAGE, P::B_STAGE, &b_full[slot]);
            }
        }
    } else if (warp == 8) {
        // =============================== MMA issuer ===============================
        if (lane == 0) {
            constexpr uint32_t idesc = (1u << 4) | (E::FMT << 7) | (E::FMT << 10) | ((uint32_t)(BN >> 3) << 17) |
                                       ((uint32_t)(UM >> 4) << 24);
            constexpr int MMAS = E::CH / 2;  // instructions per 32-channel group (2 chunks = 32 B of K each)
            const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB);
            uint32_t accum = 0;
            if (TAPS == 1) {
                for (int kg = 0; kg < KG; ++kg) {
                    const int sa = kg % NA, sb = kg % NB1;
                    mbar_wait(&a_full[sa], (kg / NA) & 1);
                    mbar_wait(&b_full[sb], (kg / NB1) & 1);
                    tc_fence_after();
#pragma unroll
                    for (int k = 0; k < MMAS; ++k) {
                        const uint64_t ad = make_desc(sA_u + sa * P::A_SLOT + 2 * k * P::A_LBO, P::A_LBO, 128);
                        const uint64_t bd = make_desc(sB_u + sb * P::B_STAGE + 2 * k * BN * 16, BN * 16, 128);
                        if (SPLIT) {   // small terms first: a_lo b_hi, a_hi b_lo, then a_hi b_hi
                            const uint64_t al = make_desc(sA_u + sa * P::A_SLOT + P::A_HALF + 2 * k * P::A_LBO, P::A_LBO, 128);
                            const uint64_t bl = make_desc(sB_u + sb * P::B_STAGE + P::B_HALF + 2 * k * BN * 16, BN * 16, 128);
                            umma<ELT>(tmem_base, al, bd, idesc, accum);
                            umma<ELT>(tmem_base, ad, bl, idesc, 1);
                            accum = 1;
                        }
                        umma<ELT>(tmem_base, ad, bd, idesc, accum);
                        accum = 1;
                    }
                    umma_commit(&a_empty[sa]);
                    umma_commit(&b_empty[sb]);
                }
            } else {
                for (int g = 0; g < KG; ++g) {
                    const int sa = g & 1;
                    mbar_wait(&a_full[sa], (g >> 1) & 1);
                    for (int t = 0; t < TAPS; ++t) {
                        const int j = g * TAPS + t;
                        const int sb = j % P::B_SLOTS;
                        mbar_wait(&b_full[sb], (j / P::B_SLOTS) & 1);
                        tc_fence_after();
                        const int shift = TAPS == 3 ? t * a.wp : (t / 3) * a.wp + (t % 3);
#pragma unroll
                        for (int k = 0; k < MMAS; ++k) {
                            const uint64_t ad =
                                make_desc(sA_u + sa * P::A_SLOT + 2 * k * P::A_LBO + shift * 16, P::A_LBO, 128);
                            const uint64_t bd = make_desc(sB_u + sb * P::B_STAGE + 2 * k * BN * 16, BN * 16, 128);
                            if (SPLIT) {
                                const uint64_t al = make_desc(sA_u + sa * P::A_SLOT + P::A_HALF + 2 * k * P::A_LBO + shift * 16, P::A_LBO, 128);
                                const uint64_t bl = make_desc(sB_u + sb * P::B_STAGE + P::B_HALF + 2 * k * BN * 16, BN * 16, 128);
                                umma<ELT>(tmem_base, al, bd, idesc, accum);
                                umma<ELT>(tmem_base, ad, bl, idesc, 1);
                                accum = 1;
                            }
                            umma<ELT>(tmem_base, ad, bd, idesc, accum);
                            accum = 1;
                        }
                        umma_commit(&b_empty[sb]);
                    }
                    umma_commit(&a_empty[sa]);
                }
            }
            umma_commit(tmem_full);
        }
    } else {
        // =============================== epilogue (warps 4-7) ===============================
        const int e = warp - 4;              // TMEM lane partition of this warp
        const int row = e * 32 + lane;       // accumulator row == tile row
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        bool valid;
        if (TAPS == 3) {
            valid = true;                    // partial sums of every patch row are needed by the rows before it
        } else if (TAPS == 9) {
            const int i = row / a.wp, j = row - i * a.wp;
            valid = i < a.ht && j < a.wp - 2 && h0 + i < hout && w0 + j < hout;
        } else {
            valid = m0 + row < hw_out;
        }
#pragma unroll
        for (int cb = 0; cb < BN; cb += 32) {
            float v[32];
            tmem_ld32(tmem_base + ((uint32_t)(e * 32) << 16) + cb, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) s_out[row * (BN + 1) + cb + i] = valid ? v[i] : 0.f;
        }
    }
    // the accumulator tile is staged: every warp of the CTA (producers, MMA and loader warps are idle by now) shares
    // the coalesced stores and the per-channel statistics
    tc_fence_before();
    __syncthreads();
    {
        constexpr int NW = 14;
        for (int r = warp; r < UM; r += NW) {
            int pix;
            bool ok;
            if (TAPS != 1) {
                const int i = r / a.wp, j = r - i * a.wp;
                ok = i < a.ht && j < a.wp - 2 && h0 + i < hout && w0 + j < hout;
                pix = (h0 + i) * hout + w0 + j;
            } else {
                ok = m0 + r < hw_out;
                pix = m0 + r;
            }
            if (TAPS == 3) {
                // dx-merged accumulator: combine the three column blocks of rows r, r + 1, r + 2 and keep the result in the
                // first block of row r (nobody else reads E_0[r]), where the statistics below expect the tile's output
                float v = 0.f;
                if (ok) v = (s_out[r * (BN + 1) + lane] + s_out[(r + 1) * (BN + 1) + NOUT + lane]) + s_out[(r + 2) * (BN + 1) + 2 * NOUT + lane];
                s_out[r * (BN + 1) + lane] = v;
            }
            if (!ok) continue;
            float* o = a.out + ((size_t)s * hw_out + pix) * a.out_cstride + a.out_coff + ntile * NOUT;
            const int nvalid = a.cout - ntile * NOUT;   // < NOUT only on the zero-padded last tile of a data-gradient convolution
#pragma unroll
            for (int cb = 0; cb < NOUT; cb += 32)
                if (cb + lane < nvalid) o[cb + lane] = s_out[r * (BN + 1) + cb + lane];
        }
        if (TAPS == 3) __syncthreads();   // the combined rows were written by different warps than the ones that sum them
        // per-channel statistics of this tile (invalid rows were staged as zeros): partial column sums by all
        // threads, combined in shared memory (the scale/shift tables are dead by now), ONE double atomic pair per
        // channel and tile
        if (a.bnr_sums != nullptr) {
            // BatchNorm(+ReLU) backward reduction over this tile: channel = column, rows split over the row groups
            constexpr int GROUPS = 448 / NOUT;
            constexpr int RPG = (UM + GROUPS - 1) / GROUPS;
            float* red = s_sc;
            const int cidx = tid % NOUT, g = tid / NOUT;
            const int ch = ntile * NOUT + cidx;
            if (g < GROUPS) {
                float s1 = 0.f, s2 = 0.f;
                if (ch < a.cout) {
                    const double cnt = (double)hout * hout;
                    const double* stp = a.bnr_stats + 2 * ((size_t)s * a.bnr_stats_stride + ch);
                    const double md = stp[0] / cnt;
                    double var = stp[1] / cnt - md * md;
                    if (var < 0) var = 0;
                    const float mean = (float)md, rstd = (float)(1.0 / sqrt(var + (double)kBnEps));
                    const float sc = a.bnr_gamma[ch] * rstd, sh = a.bnr_beta[ch] - mean * sc;
                    const float* xb = a.bnr_x + (size_t)s * hw_out * a.bnr_x_cstride + ch;
                    // branch-free batches of eight rows: the eight raw-activation loads are in flight together (a serial
                    // loop of ~43 dependent global loads per thread cost more than the separate reduction kernel it replaces)
                    const int rbeg = g * RPG, rend = (g + 1) * RPG < UM ? (g + 1) * RPG : UM;
                    for (int rb = rbeg; rb < rend; rb += 8) {
                        float xv[8], w[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const int r = rb + u;
                            int pix;
                            bool ok;
                            if (TAPS != 1) {
                                const int i = r / a.wp, j = r - i * a.wp;
                                ok = r < rend && i < a.ht && j < a.wp - 2 && h0 + i < hout && w0 + j < hout;
                                pix = (h0 + i) * hout + w0 + j;
                            } else {
                                ok = r < rend && m0 + r < hw_out;
                                pix = m0 + r;
                            }
                            w[u] = ok ? 1.f : 0.f;
                            xv[u] = __ldg(xb + (size_t)(ok ? pix : 0) * a.bnr_x_cstride);
                        }
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const int r = rb + u < UM ? rb + u : UM - 1;
                            const float dz = fmaf(xv[u], sc, sh) > 0.f ? w[u] * s_out[r * (BN + 1) + cidx] : 0.f;
                            s1 += dz;
                            s2 = fmaf(dz, (xv[u] - mean) * rstd, s2);
                        }
                    }
                }
                red[g * NOUT + cidx] = s1;
                red[(GROUPS + g) * NOUT + cidx] = s2;
            }
            __syncthreads();
            if (tid < NOUT && ntile * NOUT + tid < a.cout) {
                double t1 = 0.0, t2 = 0.0;
#pragma unroll
                for (int g2 = 0; g2 < GROUPS; ++g2) {
                    t1 += (double)red[g2 * NOUT + tid];
                    t2 += (double)red[(GROUPS + g2) * NOUT + tid];
                }
                double* sm = a.bnr_sums + 2 * ((size_t)s * a.cout + ntile * NOUT + tid);
                atomicAdd(sm, t1);
                atomicAdd(sm + 1, t2);
            }
            __syncthreads();
        }
        if (a.out_stats != nullptr) {
            constexpr int GROUPS = 448 / NOUT;                        // row groups: 3 (N=128), 7 (N=64), 14 (N=32)
            constexpr int RPG = (UM + GROUPS - 1) / GROUPS;          // rows per group
            float* red = s_sc;                                       // [2][GROUPS][BN] floats <= 3.5 KB
            const int cidx = tid % NOUT, g = tid / NOUT;
            if (g < GROUPS) {
                float su = 0.f, sq = 0.f;
                const int r1 = (g + 1) * RPG < UM ? (g + 1) * RPG : UM;
                for (int r = g * RPG; r < r1; ++r) {
                    const float x = s_out[r * (BN + 1) + cidx];
                    su += x;
                    sq = fmaf(x, x, sq);
                }
                red[g * NOUT + cidx] = su;
                red[(GROUPS + g) * NOUT + cidx] = sq;
            }
            __syncthreads();
            if (tid < NOUT && ntile * NOUT + tid < a.cout) {
                double su = 0.0, sq = 0.0;
#pragma unroll
                for (int g2 = 0; g2 < GROUPS; ++g2) {
                    su += (double)red[g2 * NOUT + tid];
                    sq += (double)red[(GROUPS + g2) * NOUT + tid];
                }
                double* st = a.out_stats + 2 * ((size_t)s * a.out_stats_stride + a.out_coff + ntile * NOUT + tid);
                atomicAdd(st, su);
                atomicAdd(st + 1, sq);
            }
        }
    }
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMC);
    }
}

// ------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------
void umma_patch_geometry(int hw, int* wp, int* ht) {
    // patch width wp = tile width + 2 halo columns; ht*wp <= 128 output rows per MMA tile; the patch a tile loads is
    // (ht+2) x wp pixels.  The 3x3 kernels are bound by the bytes they pull in, so the tile shape is chosen to minimise
    // loaded pixels per output pixel (squarer tiles: 18 x 7 loads 1.46 pixels per output at 160^2, the widest tile 42 x 3
    // loads 1.77), with a mild penalty on the number of tiles (= MMA work).  SMG_PATCH_GEOM=0 restores the widest-tile rule.
    static const bool widest = getenv("SMG_PATCH_GEOM") != nullptr && atoi(getenv("SMG_PATCH_GEOM")) == 0;
    int best_wp = 0, best_ht = 0;
    double best_cost = 1e30;
    int min_tiles = 1 << 30;
    for (int pass = 0; pass < 2; ++pass) {
        for (int nt = 1; nt <= hw; ++nt) {
            const int wt = (hw + nt - 1) / nt, w = wt + 2;
            if (w > MAX_WP) continue;
            int t = UM / w;
            if (t > hw) t = hw;
            if (t < 1 || (t + 2) * w > 5 * MAX_WP || 2 * w + UM > 212) continue;
            const int tiles = ((hw + wt - 1) / wt) * ((hw + t - 1) / t);
            if (pass == 0) {
                if (tiles < min_tiles) min_tiles = tiles;
                if (widest) { best_wp = w; best_ht = t; nt = hw; }   // first admissible = fewest column tiles
                continue;
            }
            const double loaded = (double)tiles * (t + 2) * w / ((double)hw * hw);
            const double cost = loaded + 0.8 * ((double)tiles / min_tiles - 1.0);
            if (cost < best_cost) { best_cost = cost; best_wp = w; best_ht = t; }
        }
        if (widest) break;
    }
    *wp = best_wp;
    *ht = best_ht;
}

template <int ELT, int BN, int TAPS, int POOL, int SPLIT = 0>
static int launch_umma(smg_handle* h, const UmmaDev& d, int n, cudaStream_t st) {
    using P = SmemPlan<ELT, BN, TAPS, SPLIT>;
    SMG_TRY(ensure_dyn_smem(h, (const void*)conv_umma_kernel<ELT, BN, TAPS, POOL, SPLIT>, P::TOTAL));
    dim3 grid;
    if (TAPS != 1) {
        const int wt = d.wp - 2;
        const int tx = (d.hout + wt - 1) / wt, ty = (d.hout + d.ht - 1) / d.ht;
        grid = dim3(tx * ty, 1, n);
    } else {
        grid = dim3((d.hout * d.hout + UM - 1) / UM, (d.cout + BN - 1) / BN, n);
    }
    UmmaDev dd = d;
    dd.async_producer = h->force_async >= 0 ? h->force_async : ((int)(grid.x * grid.y * grid.z) < 2 * h->num_sms ? 1 : 0);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(448);
    cfg.dynamicSmemBytes = P::TOTAL;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = h->use_pdl ? 1 : 0;
    SMG_CUDA(cudaLaunchKernelEx(&cfg, conv_umma_kernel<ELT, BN, TAPS, POOL, SPLIT>, dd));
    h->launches++;
    SMG_CUDA(cudaGetLastError());
    return SMG_OK;
}

template <int ELT>
static int dispatch(smg_handle* h, const ConvArgs& a, UmmaDev& d, cudaStream_t st) {
    d.w = a.w_umma != nullptr ? a.w_umma : (ELT == 4 ? a.w->w_tf32 : a.w->w_bf16);
    SMG_CHECK(d.w != nullptr, SMG_ERR_STATE, "conv_umma: weights not packed");
    if (a.taps == 9) {
        const bool fwd = a.cin == 128 && a.cout == 32, dgrad = ELT == 4 && a.cin == 32 && a.cout == 128 && a.w_umma != nullptr;
        SMG_CHECK((fwd || dgrad) && !a.pool, SMG_ERR_UNSUPPORTED, "conv_umma: 3x3 expects 128->32 (or its 32->128 data gradient)");
        umma_patch_geometry(d.hout, &d.wp, &d.ht);
        SMG_CHECK((d.ht + 2) * d.wp <= 5 * MAX_WP && d.ht * d.wp <= UM, SMG_ERR_STATE, "conv_umma: patch %dx%d too large",
                  d.ht, d.wp);
        const int wt = d.wp - 2;
        d.tiles_x = (d.hout + wt - 1) / wt;
        if (dgrad) return launch_umma<4, 128, 9, 0>(h, d, a.n, st);
        return launch_umma<ELT, 32, 9, 0>(h, d, a.n, st);
    }
    SMG_CHECK(a.cin % KC == 0 && a.cin <= 1024, SMG_ERR_UNSUPPORTED, "conv_umma: cin %d unsupported", a.cin);
    if (a.w_umma != nullptr) {
        // data-gradient 1x1: the packed image is zero-padded to whole 128-channel output tiles
        SMG_CHECK(ELT == 4 && !a.pool, SMG_ERR_UNSUPPORTED, "conv_umma: packed-image override is tf32, unpooled only");
        return launch_umma<4, 128, 1, 0>(h, d, a.n, st);
    }
    if (a.cout == 64) {
        SMG_CHECK(!a.pool, SMG_ERR_UNSUPPORTED, "conv_umma: pooled N=64 not built");
        return launch_umma<ELT, 64, 1, 0>(h, d, a.n, st);
    }
    SMG_CHECK(a.cout % 128 == 0, SMG_ERR_UNSUPPORTED, "conv_umma: cout %d unsupported", a.cout);
    if (a.pool) return launch_umma<ELT, 128, 1, 1>(h, d, a.n, st);
    return launch_umma<ELT, 128, 1, 0>(h, d, a.n, st);
}

// fp32-accurate mode on the tensor cores: tf32 elements, hi/lo split operands (w_split = per stage [hi image][lo image])
static int dispatch_split(smg_handle* h, const ConvArgs& a, UmmaDev& d, cudaStream_t st) {
    SMG_CHECK(a.w != nullptr && a.w->w_split != nullptr, SMG_ERR_STATE, "conv_umma: split weights not packed");
    d.w = a.w->w_split;
    if (a.taps == 9) {
        SMG_CHECK(a.cin == 128 && a.cout == 32 && !a.pool, SMG_ERR_UNSUPPORTED, "conv_umma: 3x3 expects 128->32");
        umma_patch_geometry(d.hout, &d.wp, &d.ht);
        SMG_CHECK((d.ht + 2) * d.wp <= 5 * MAX_WP && d.ht * d.wp <= UM, SMG_ERR_STATE, "conv_umma: patch %dx%d too large", d.ht, d.wp);
        const int wt = d.wp - 2;
        d.tiles_x = (d.hout + wt - 1) / wt;
        return launch_umma<4, 96, 3, 0, 1>(h, d, a.n, st);   // dx merged into N: a third of the MMAs of the nine-tap form
    }
    SMG_CHECK(a.cin % KC == 0 && a.cin <= 1024, SMG_ERR_UNSUPPORTED, "conv_umma: cin %d unsupported", a.cin);
    if (a.cout == 64) {
        SMG_CHECK(!a.pool, SMG_ERR_UNSUPPORTED, "conv_umma: pooled N=64 not built");
        return launch_umma<4, 64, 1, 0, 1>(h, d, a.n, st);
    }
    SMG_CHECK(a.cout % 128 == 0, SMG_ERR_UNSUPPORTED, "conv_umma: cout %d unsupported", a.cout);
    if (a.pool) return launch_umma<4, 128, 1, 1, 1>(h, d, a.n, st);
    return launch_umma<4, 128, 1, 0, 1>(h, d, a.n, st);
}

int launch_conv_umma(smg_handle* h, const ConvArgs& a, int precision, cudaStream_t st) {
    SMG_CHECK(a.w != nullptr || a.w_umma != nullptr, SMG_ERR_STATE, "conv_umma: no weights");
    UmmaDev d;
    d.in = a.in; d.in_cstride = a.in_cstride; d.cin = a.cin; d.hin = a.hin;
    d.prologue_mode = a.prologue_mode; d.in_stats = a.in_stats; d.in_stats_stride = a.in_stats_stride;
    d.gamma = a.gamma; d.beta = a.beta; d.scale = a.scale; d.shift = a.shift; d.relu = a.relu;
    d.out = a.out; d.out_cstride = a.out_cstride; d.out_coff = a.out_coff; d.cout = a.cout;
    d.out_stats = a.out_stats; d.out_stats_stride = a.out_stats_stride;
    d.bnr_x = a.bnr_x; d.bnr_x_cstride = a.bnr_x_cstride; d.bnr_stats = a.bnr_stats; d.bnr_stats_stride = a.bnr_stats_stride;
    d.bnr_gamma = a.bnr_gamma; d.bnr_beta = a.bnr_beta; d.bnr_sums = a.bnr_sums;
    d.hout = a.pool ? a.hin / 2 : a.hin;
    d.wp = d.ht = d.tiles_x = 0;
    d.async_producer = 0;
    if (precision == SMG_PREC_TF32) return dispatch<4>(h, a, d, st);
    if (precision == SMG_PREC_BF16) return dispatch<2>(h, a, d, st);
    if (precision == SMG_PREC_FP32) return dispatch_split(h, a, d, st);
    set_error("conv_umma: precision %d is not a tensor-core mode", precision);
    return SMG_ERR_INVALID;
}

}  // namespace smg
