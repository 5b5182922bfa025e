// conv_umma_mt.cu - multi-tile variant of the tcgen05 convolution (conv_umma.cu) for the large launches.
//
// ncu on the one-tile-per-CTA kernel (profiles/README.md) showed neither HBM nor the tensor pipe saturated: the CTA
// spends a third of its life in its prologue (barrier init, TMEM allocation, BN scale/shift table) and epilogue
// (TMEM -> staging -> global, statistics) during which it has no loads in flight.  Here a CTA walks T consecutive
// tiles of ONE sample: the table, the barriers and the TMEM allocation are set up once, the accumulator is double
// buffered in TMEM (2 x BN columns) and the 4 epilogue warps drain tile i through a small 32-column staging buffer
// while the producer / MMA warps are already streaming tile i+1.
//
// Same contract, operand layout (no-swizzle K-major, shifted descriptors for the 3x3 taps), warp roles and arithmetic
// as conv_umma.cu; pooled (transition) convolutions keep using the one-tile kernel.
#include "umma_common.cuh"

namespace smg {

constexpr int NBM1 = 2;  // weight stages of the 1x1 variant (the staging buffer takes the room of the third one)

template <int ELT, int BN, int TAPS>
struct MtPlan {
    using E = EltCfg<ELT>;
    static constexpr int A_LBO = (TAPS == 9 ? E::P_ROWS : E::A_ROWS) * 16;
    static constexpr int A_SLOT = E::CH * A_LBO;
    static constexpr int A_SLOTS = TAPS == 9 ? 2 : NA;
    static constexpr int B_STAGE = E::CH * BN * 16;
    static constexpr int B_SLOTS = TAPS == 9 ? NB9 : NBM1;
    static constexpr int OFF_BAR = 0;
    static constexpr int OFF_SC = 256;
    static constexpr int OFF_A = OFF_SC + 2 * 1024 * 4;
    static constexpr int OFF_B = OFF_A + A_SLOTS * A_SLOT;
    static constexpr int OFF_STAGE = OFF_B + B_SLOTS * B_STAGE;   // [128][33] floats
    static constexpr int OFF_RED = OFF_STAGE + UM * 33 * 4;       // [2][4][32] floats
    static constexpr int TOTAL = OFF_RED + 2 * 4 * 32 * 4;
};

template <int ELT, int BN, int TAPS>
__global__ void __launch_bounds__(448, 2)
conv_umma_mt_kernel(UmmaDev a) {
    using E = EltCfg<ELT>;
    using P = MtPlan<ELT, BN, TAPS>;
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + P::OFF_BAR);
    uint64_t* a_full = bars;            // [4]
    uint64_t* a_empty = bars + 4;       // [4]
    uint64_t* b_full = bars + 8;        // [8]
    uint64_t* b_empty = bars + 16;      // [8]
    uint64_t* tmem_full = bars + 24;    // [2]
    uint64_t* tmem_empty = bars + 26;   // [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 28);
    float* s_sc = reinterpret_cast<float*>(smem + P::OFF_SC);
    float* s_sh = s_sc + 1024;
    uint8_t* sA = smem + P::OFF_A;
    uint8_t* sB = smem + P::OFF_B;
    float* s_stage = reinterpret_cast<float*>(smem + P::OFF_STAGE);
    float* s_red = reinterpret_cast<float*>(smem + P::OFF_RED);

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int s = blockIdx.z;
    const int ntile = blockIdx.y;
    const int hout = a.hout, hin = a.hin;
    const int hw_out = hout * hout;
    const int KG = a.cin / KC;
    const int tile0 = blockIdx.x * a.tiles_per_cta;
    const int ntiles = min(a.tiles_per_cta, a.tiles_per_sample - tile0);

    // ---- one-time setup
    if (warp == 8 && lane == 0) {
        for (int i = 0; i < 4; ++i) { mbar_init(&a_full[i], 128); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < 8; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 128); }
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc(tmem_ptr, 2 * BN);
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
    const float* inp = a.in + (size_t)s * hin * hin * a.in_cstride;

    if (warp < 4 || warp >= 10) {
        // =============================== A producers (two groups) ===============================
        const int pgroup = warp < 4 ? 0 : 1;
        const int ptid = tid & 127;
        const int c = ptid % E::CH;
        const int r0 = ptid / E::CH;
        constexpr int RSTEP = 128 / E::CH;
        for (int it = 0; it < ntiles; ++it) {
            const int tile = tile0 + it;
            if (TAPS == 1) {
                constexpr int RI = UM / RSTEP;
                const int m0 = tile * UM;
                int roff[RI];
#pragma unroll
                for (int i = 0; i < RI; ++i) {
                    const int m = m0 + r0 + i * RSTEP;
                    roff[i] = m < hw_out ? m * a.in_cstride : -1;
                }
                for (int kg = pgroup; kg < KG; kg += 2) {
                    const int sg = it * KG + kg;
                    const int slot = sg % NA;
                    mbar_wait(&a_empty[slot], ((sg / NA) & 1) ^ 1);
                    const int ch0 = kg * KC + c * E::EPC;
                    float sc[E::EPC], sh[E::EPC];
#pragma unroll
                    for (int e = 0; e < E::EPC; ++e) { sc[e] = s_sc[ch0 + e]; sh[e] = s_sh[ch0 + e]; }
                    uint8_t* dst = sA + slot * P::A_SLOT + c * P::A_LBO;
                    float4 v[RI][E::EPC / 4];
#pragma unroll
                    for (int i = 0; i < RI; ++i) {
#pragma unroll
                        for (int e4 = 0; e4 < E::EPC / 4; ++e4) {
                            v[i][e4] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (roff[i] >= 0) v[i][e4] = __ldg(reinterpret_cast<const float4*>(inp + roff[i] + ch0) + e4);
                        }
                    }
#pragma unroll
                    for (int i = 0; i < RI; ++i) {
                        float t[E::EPC];
#pragma unroll
                        for (int e4 = 0; e4 < E::EPC / 4; ++e4) {
                            t[4 * e4 + 0] = v[i][e4].x; t[4 * e4 + 1] = v[i][e4].y;
                            t[4 * e4 + 2] = v[i][e4].z; t[4 * e4 + 3] = v[i][e4].w;
                        }
                        if (roff[i] >= 0) {
#pragma unroll
                            for (int e = 0; e < E::EPC; ++e) {
                                const float u = fmaf(t[e], sc[e], sh[e]);
                                t[e] = a.relu ? fmaxf(u, 0.f) : u;
                            }
                        }
                        uint4 pk;
                        if (ELT == 4) {
                            pk = make_uint4(__float_as_uint(t[0]), __float_as_uint(t[1]), __float_as_uint(t[2]),
                                            __float_as_uint(t[3]));
                        } else {
                            pk = make_uint4(pack_bf16x2(t[0], t[1]), pack_bf16x2(t[2], t[3]),
                                            pack_bf16x2(t[4 % E::EPC], t[5 % E::EPC]), pack_bf16x2(t[6 % E::EPC], t[7 % E::EPC]));
                        }
                        *reinterpret_cast<uint4*>(dst + (r0 + i * RSTEP) * 16) = pk;
                    }
                    fence_proxy_async();
                    mbar_arrive(&a_full[slot]);
                }
            } else {
                const int wp = a.wp;
                const int ty = tile / a.tiles_x, tx = tile - ty * a.tiles_x;
                const int h0 = ty * a.ht, w0 = tx * (wp - 2);
                const int pfill = (a.ht + 2) * wp;
                constexpr int NI = (5 * MAX_WP + RSTEP - 1) / RSTEP;
                int poff[NI];
#pragma unroll
                for (int i = 0; i < NI; ++i) {
                    const int q = r0 + i * RSTEP;
                    const int py = q / wp, px = q - py * wp;
                    const int y = h0 - 1 + py, x = w0 - 1 + px;
                    poff[i] = (q < pfill && y >= 0 && y < hin && x >= 0 && x < hin) ? (y * hin + x) * a.in_cstride : -1;
                }
                for (int g = pgroup; g < 4; g += 2) {
                    const int slot = g & 1;
                    const int use = it * 2 + (g >> 1);          // how many times this slot has been filled before
                    if (use > 0) mbar_wait(&a_empty[slot], (use - 1) & 1);
                    const int ch0 = g * KC + c * E::EPC;
                    float sc[E::EPC], sh[E::EPC];
#pragma unroll
                    for (int e = 0; e < E::EPC; ++e) { sc[e] = s_sc[ch0 + e]; sh[e] = s_sh[ch0 + e]; }
                    uint8_t* dst = sA + slot * P::A_SLOT + c * P::A_LBO;
                    constexpr int NBAT = 4;
                    constexpr int NH = (NI + NBAT - 1) / NBAT;
#pragma unroll
                    for (int hb = 0; hb < NBAT; ++hb) {
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
        }
    } else if (warp == 9) {
        // =============================== weight loader ===============================
        if (lane == 0) {
            const int nstages = TAPS == 9 ? 36 : KG;
            const uint8_t* wsrc = a.w + (size_t)ntile * nstages * P::B_STAGE;
            for (int it = 0; it < ntiles; ++it) {
                for (int j = 0; j < nstages; ++j) {
                    const int jg = it * nstages + j;
                    const int slot = jg % P::B_SLOTS;
                    mbar_wait(&b_empty[slot], ((jg / P::B_SLOTS) & 1) ^ 1);
                    mbar_arrive_expect_tx(&b_full[slot], P::B_STAGE);
                    tma_bulk_load(sB + slot * P::B_STAGE, wsrc + (size_t)j * P::B_STAGE, P::B_STAGE, &b_full[slot]);
                }
            }
        }
    } else if (warp == 8) {
        // =============================== MMA issuer ===============================
        if (lane == 0) {
            constexpr uint32_t idesc = (1u << 4) | (E::FMT << 7) | (E::FMT << 10) | ((uint32_t)(BN >> 3) << 17) |
                                       ((uint32_t)(UM >> 4) << 24);
            constexpr int MMAS = E::CH / 2;
            const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB);
            for (int it = 0; it < ntiles; ++it) {
                const int buf = it & 1;
                mbar_wait(&tmem_empty[buf], ((it >> 1) & 1) ^ 1);   // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
                uint32_t accum = 0;
                if (TAPS == 1) {
                    for (int kg = 0; kg < KG; ++kg) {
                        const int sg = it * KG + kg;
                        const int sa = sg % NA, sb = sg % NBM1;
                        mbar_wait(&a_full[sa], (sg / NA) & 1);
                        mbar_wait(&b_full[sb], (sg / NBM1) & 1);
                        tc_fence_after();
#pragma unroll
                        for (int k = 0; k < MMAS; ++k) {
                            const uint64_t ad = make_desc(sA_u + sa * P::A_SLOT + 2 * k * P::A_LBO, P::A_LBO, 128);
                            const uint64_t bd = make_desc(sB_u + sb * P::B_STAGE + 2 * k * BN * 16, BN * 16, 128);
                            umma<ELT>(d_tmem, ad, bd, idesc, accum);
                            accum = 1;
                        }
                        umma_commit(&a_empty[sa]);
                        umma_commit(&b_empty[sb]);
                    }
                } else {
                    for (int g = 0; g < 4; ++g) {
                        const int sa = g & 1;
                        const int use = it * 2 + (g >> 1);
                        mbar_wait(&a_full[sa], use & 1);
                        for (int t = 0; t < 9; ++t) {
                            const int j = (it * 4 + g) * 9 + t;
                            const int sb = j % NB9;
                            mbar_wait(&b_full[sb], (j / NB9) & 1);
                            tc_fence_after();
                            const int shift = (t / 3) * a.wp + (t % 3);
#pragma unroll
                            for (int k = 0; k < MMAS; ++k) {
                                const uint64_t ad =
                                    make_desc(sA_u + sa * P::A_SLOT + 2 * k * P::A_LBO + shift * 16, P::A_LBO, 128);
                                const uint64_t bd = make_desc(sB_u + sb * P::B_STAGE + 2 * k * BN * 16, BN * 16, 128);
                                umma<ELT>(d_tmem, ad, bd, idesc, accum);
                                accum = 1;
                            }
                            umma_commit(&b_empty[sb]);
                        }
                        umma_commit(&a_empty[sa]);
                    }
                }
                umma_commit(&tmem_full[buf]);
            }
        }
    } else {
        // =============================== epilogue (warps 4-7) ===============================
        const int e = warp - 4;              // TMEM lane partition of this warp
        const int row = e * 32 + lane;       // accumulator row == tile row
        const int t = tid - 128;             // 0..127 inside the epilogue group
        for (int it = 0; it < ntiles; ++it) {
            const int tile = tile0 + it;
            const int buf = it & 1;
            int m0 = 0, h0 = 0, w0 = 0;
            if (TAPS == 9) {
                const int ty = tile / a.tiles_x, tx = tile - ty * a.tiles_x;
                h0 = ty * a.ht;
                w0 = tx * (a.wp - 2);
            } else {
                m0 = tile * UM;
            }
            bool valid;
            if (TAPS == 9) {
                const int i = row / a.wp, j = row - i * a.wp;
                valid = i < a.ht && j < a.wp - 2 && h0 + i < hout && w0 + j < hout;
            } else {
                valid = m0 + row < hw_out;
            }
            mbar_wait(&tmem_full[buf], (it >> 1) & 1);
            tc_fence_after();
            for (int cb = 0; cb < BN; cb += 32) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(e * 32) << 16) + (uint32_t)(buf * BN + cb), v);
#pragma unroll
                for (int i = 0; i < 32; ++i) s_stage[row * 33 + i] = valid ? v[i] : 0.f;
                asm volatile("bar.sync 1, 128;" ::: "memory");
                // coalesced 128-byte row segments: warp e stores rows e, e+4, ...
                for (int r = e; r < UM; r += 4) {
                    int pix;
                    bool ok;
                    if (TAPS == 9) {
                        const int i = r / a.wp, j = r - i * a.wp;
                        ok = i < a.ht && j < a.wp - 2 && h0 + i < hout && w0 + j < hout;
                        pix = (h0 + i) * hout + w0 + j;
                    } else {
                        ok = m0 + r < hw_out;
                        pix = m0 + r;
                    }
                    if (ok)
                        a.out[((size_t)s * hw_out + pix) * a.out_cstride + a.out_coff + ntile * BN + cb + lane] =
                            s_stage[r * 33 + lane];
                }
                if (a.out_stats != nullptr) {
                    // column statistics of this 32-column chunk: 4 row groups x 32 columns, combined in smem
                    const int col = t & 31, grp = t >> 5;
                    float su = 0.f, sq = 0.f;
                    for (int r = grp * 32; r < grp * 32 + 32; ++r) {
                        const float x = s_stage[r * 33 + col];
                        su += x;
                        sq = fmaf(x, x, sq);
                    }
                    s_red[grp * 32 + col] = su;
                    s_red[128 + grp * 32 + col] = sq;
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (a.out_stats != nullptr && t < 32) {
                    const double su = (double)s_red[t] + (double)s_red[32 + t] + (double)s_red[64 + t] + (double)s_red[96 + t];
                    const double sq = (double)s_red[128 + t] + (double)s_red[160 + t] + (double)s_red[192 + t] +
                                      (double)s_red[224 + t];
                    double* st = a.out_stats + 2 * ((size_t)s * a.out_stats_stride + a.out_coff + ntile * BN + cb + t);
                    atomicAdd(st, su);
                    atomicAdd(st + 1, sq);
                }
            }
            tc_fence_before();
            mbar_arrive(&tmem_empty[buf]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 2 * BN);
    }
}

template <int ELT, int BN, int TAPS>
static int launch_mt(smg_handle* h, UmmaDev d, int n, int tiles_per_cta, cudaStream_t st) {
    using P = MtPlan<ELT, BN, TAPS>;
    static bool attr = false;
    if (!attr) {
        SMG_CUDA(cudaFuncSetAttribute(conv_umma_mt_kernel<ELT, BN, TAPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, P::TOTAL));
        attr = true;
    }
    if (TAPS == 9) {
        const int wt = d.wp - 2;
        d.tiles_x = (d.hout + wt - 1) / wt;
        d.tiles_per_sample = d.tiles_x * ((d.hout + d.ht - 1) / d.ht);
    } else {
        d.tiles_per_sample = (d.hout * d.hout + UM - 1) / UM;
    }
    d.tiles_per_cta = tiles_per_cta;
    dim3 grid((d.tiles_per_sample + tiles_per_cta - 1) / tiles_per_cta, TAPS == 9 ? 1 : d.cout / BN, n);
    conv_umma_mt_kernel<ELT, BN, TAPS><<<grid, 448, P::TOTAL, st>>>(d);
    h->launches++;
    SMG_CUDA(cudaGetLastError());
    return SMG_OK;
}

// returns SMG_ERR_UNSUPPORTED (without setting an error) if the multi-tile kernel does not serve this shape
int launch_conv_umma_mt(smg_handle* h, const ConvArgs& a, int precision, int tiles_per_cta, cudaStream_t st) {
    if (a.pool || a.w == nullptr) return SMG_ERR_UNSUPPORTED;
    UmmaDev d;
    d.in = a.in; d.in_cstride = a.in_cstride; d.cin = a.cin; d.hin = a.hin;
    d.prologue_mode = a.prologue_mode; d.in_stats = a.in_stats; d.in_stats_stride = a.in_stats_stride;
    d.gamma = a.gamma; d.beta = a.beta; d.scale = a.scale; d.shift = a.shift; d.relu = a.relu;
    d.out = a.out; d.out_cstride = a.out_cstride; d.out_coff = a.out_coff; d.cout = a.cout;
    d.out_stats = a.out_stats; d.out_stats_stride = a.out_stats_stride;
    d.hout = a.hin;
    d.wp = d.ht = d.tiles_x = 0;
    d.async_producer = 0;
    d.w = precision == SMG_PREC_TF32 ? a.w->w_tf32 : a.w->w_bf16;
    if (d.w == nullptr || a.cin % KC != 0 || a.cin > 1024) return SMG_ERR_UNSUPPORTED;
    if (a.taps == 9) {
        if (a.cin != 128 || a.cout != 32) return SMG_ERR_UNSUPPORTED;
        umma_patch_geometry(d.hout, &d.wp, &d.ht);
        if ((d.ht + 2) * d.wp > 5 * MAX_WP || d.ht * d.wp > UM) return SMG_ERR_UNSUPPORTED;
        return precision == SMG_PREC_TF32 ? launch_mt<4, 32, 9>(h, d, a.n, tiles_per_cta, st)
                                          : launch_mt<2, 32, 9>(h, d, a.n, tiles_per_cta, st);
    }
    if (a.cout % 128 != 0) return SMG_ERR_UNSUPPORTED;
    return precision == SMG_PREC_TF32 ? launch_mt<4, 128, 1>(h, d, a.n, tiles_per_cta, st)
                                      : launch_mt<2, 128, 1>(h, d, a.n, tiles_per_cta, st);
}

}  // namespace smg
