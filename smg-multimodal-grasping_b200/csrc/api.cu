// api.cu - C ABI entry points, workspace management and the trunk / Q-pass schedule.
//
// The schedule of one trunk pass over n samples (torchvision densenet121().features as called at
// /root/reference/code/models.py:384-385) is a fixed kernel sequence on one stream:
//   conv0+stats -> norm0/relu/maxpool+stats -> for each dense layer {1x1 conv (BN-ReLU prologue,
//   stats epilogue) -> 3x3 conv (same), written into the block buffer slice} -> per transition
//   {BN-ReLU + 2x2 avg-pool prologue, 1x1 conv, stats} ...  norm5 is folded into the head prologue.
// All n samples run through every kernel together (M = n * H*W rows), with BatchNorm statistics
// kept per sample, exactly like the reference's batch-1 calls.
#include "smg_internal.cuh"

#include <stdarg.h>
#include <string.h>

namespace smg {

static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* get_error() { return g_err; }


// brackets the launches issued inside its lifetime with CUDA events when profiling is enabled
struct ProfScope {
    smg_handle* h;
    cudaStream_t st;
    size_t idx = (size_t)-1;
    ProfScope(smg_handle* h_, cudaStream_t st_, int cls, double flops, double bytes) : h(h_), st(st_) {
        if (!h->profile) return;
        smg_handle::ProfRec r;
        r.cls = cls; r.flops = flops; r.bytes = bytes;
        cudaEventCreate(&r.e0);
        cudaEventCreate(&r.e1);
        cudaEventRecord(r.e0, st);
        idx = h->prof.size();
        h->prof.push_back(r);
    }
    ~ProfScope() {
        if (idx != (size_t)-1) cudaEventRecord(h->prof[idx].e1, st);
    }
};

int conv_dispatch(smg_handle* h, const ConvArgs& a, cudaStream_t st) {
    const int hout = a.pool ? a.hin / 2 : a.hin;
    const double px = (double)a.n * hout * hout;
    // algorithmic traffic: every input element of the layer read once, every output written once (fp32)
    const double bytes = 4.0 * ((double)a.n * a.hin * a.hin * a.cin + px * a.cout);
    ProfScope ps(h, st, a.taps == 9 ? 2 : 1, 2.0 * px * a.cout * a.cin * a.taps, bytes);
    if (h->precision == SMG_PREC_FP32) {
        // fp32 mode: tensor cores with hi/lo split tf32 operands (error ~1e-6, like fp32 FMA chains); CUDA cores on request
        if (h->fp32_tc && !h->fp32_exact_pass) {
            const int status = launch_conv_umma(h, a, SMG_PREC_FP32, st);
            if (status != SMG_ERR_UNSUPPORTED) return status;
        }
        return launch_conv_ffma(h, a, st);
    }
    // tf32: the persistent TMA-fed kernels (SMG_TMA bits: 32 = trans_t.cu, 64 = conv3_wt.cu, 128 = conv1_t.cu); everything they
    // do not serve (the head's 1x1, bf16 mode, odd shapes) runs on the register-producer kernel.
    if (h->precision == SMG_PREC_TF32 && a.pool && a.taps == 1 && (h->use_tma & 32)) {
        const int status = launch_trans_t(h, a, st);
        if (status != SMG_ERR_UNSUPPORTED) return status;
    }
    if (h->precision == SMG_PREC_TF32 && !a.pool) {
        if (a.taps == 1 && (h->use_tma & 128)) {
            const int status = launch_conv1_t(h, a, st);
            if (status != SMG_ERR_UNSUPPORTED) return status;
        }
        if (a.taps == 9 && (h->use_tma & 64)) {
            const int status = launch_conv3_wt(h, a, st);
            if (status != SMG_ERR_UNSUPPORTED) return status;
        }
    }
    return launch_conv_umma(h, a, h->precision, st);
}

// ---------------------------------------------------------------------------------------
// trunk forward over `n` samples already resident in h->input
// ---------------------------------------------------------------------------------------
// Samples are independent (per-sample BatchNorm), so each dense block is run over CHUNKS of samples sized to keep the
// chunk's block buffer + bottleneck scratch inside the 126 MB L2: the 2 x layers re-reads of the block buffer then hit
// L2 instead of HBM (block 1: 39 MB per sample -> 2-3 samples per chunk; blocks 3-4: all samples at once).
static int chunk_samples(const smg_handle* h, int b, int n) {
    const BlockGeom& g = h->geom[b];
    double per_sample = (double)g.hw * g.hw * (g.c_tot + kBottleneck) * 4.0;
    if (b == 0) per_sample += (double)(h->H / 2) * (h->H / 2) * 64 * 4.0;  // conv0 output is consumed by pool0 in-chunk
    if (h->l2_chunk_bytes <= 0.0) return n;  // chunking disabled
    const double q = h->l2_chunk_bytes / per_sample;
    int cs = q >= (double)n ? n : (int)q;
    if (cs < 1) cs = 1;
    return cs;
}

int trunk_forward(smg_handle* h, int trunk_id, int n, int in_channels, cudaStream_t st, bool save_bott) {
    TrunkW& T = h->trunks[trunk_id];
    SMG_CHECK(T.set, SMG_ERR_STATE, "trunk %d: weights not set (call smg_set_trunk_weights)", trunk_id);
    SMG_CHECK(n >= 1 && n <= h->max_samples, SMG_ERR_INVALID, "trunk_forward: n=%d outside [1,%d]", n, h->max_samples);
    // every pass overwrites the workspace a pending smg_qforward_train result lives in: its backward must fail, not
    // silently differentiate another pass's activations (smg_qforward_train re-validates after its own pass)
    h->train.valid = false;
    {
        const int need = h->precision == SMG_PREC_FP32 ? SMG_PACK_FFMA : (h->precision == SMG_PREC_TF32 ? SMG_PACK_TF32 : SMG_PACK_BF16);
        SMG_CHECK(T.packed & need, SMG_ERR_STATE,
                  "trunk %d: weights were packed with layout mask %d, precision %d needs %d (set the weights again)", trunk_id,
                  T.packed, h->precision, need);
    }
    SMG_CUDA(cudaMemsetAsync(h->stats, 0, h->stats_bytes, st));
    // fp32 mode exists for parity: the grad-enabled pass keeps plain fp32 FMA chains (3e-6 of the reference), because the
    // gradients of this network amplify forward noise through its ReLU kinks (tests/test_gpu_backward.py); the volatile
    // passes use the tensor cores with split operands (3e-5, inside the 1e-4 bar)
    h->fp32_exact_pass = save_bott;
    const size_t in_img = (size_t)in_channels * h->H * h->H;
    const size_t c0_img = (size_t)(h->H / 2) * (h->H / 2) * 64;
    int layer_base = 0;
    for (int b = 0; b < kNumBlocks; ++b) {
        const BlockGeom& g = h->geom[b];
        const size_t px = (size_t)g.hw * g.hw;
        const int cs = save_bott ? n : chunk_samples(h, b, n);
        for (int s0 = 0; s0 < n; s0 += cs) {
            const int ns = s0 + cs <= n ? cs : n - s0;
            double* st_blk = stats_ptr(h, h->st_block[b]) + 2 * (size_t)s0 * g.c_tot;
            float* blk = h->block[b] + (size_t)s0 * px * g.c_tot;
            if (b == 0) {
                const double hc = (double)c0_img / 64;
                // stem traffic: input read + conv0 written, conv0 read + pooled output written
                ProfScope ps(h, st, 0, 2.0 * ns * hc * 64 * 49 * in_channels,
                             4.0 * ns * ((double)in_channels * h->H * h->H + 2 * hc * 64 + hc / 4 * 64));
                double* st_c0 = stats_ptr(h, h->st_conv0) + 2 * (size_t)s0 * 64;
                int c0_status = SMG_ERR_UNSUPPORTED;
                if ((h->precision == SMG_PREC_TF32 || (h->precision == SMG_PREC_FP32 && h->fp32_tc)) && in_channels == 1 &&
                    (h->use_tma & 16))
                    c0_status = launch_conv0_umma(h, h->input + (size_t)s0 * in_img, ns, T.conv0_umma, h->conv0 + (size_t)s0 * c0_img,
                                                  st_c0, st);
                if (c0_status == SMG_ERR_UNSUPPORTED)
                    c0_status = launch_conv0(h, h->input + (size_t)s0 * in_img, in_channels, ns,
                                             in_channels == 1 ? T.conv0_folded : T.conv0, h->conv0 + (size_t)s0 * c0_img, st_c0, st);
                SMG_TRY(c0_status);
                SMG_TRY(launch_pool0(h, ns, h->conv0 + (size_t)s0 * c0_img, st_c0, T.norm0.gamma, T.norm0.beta, blk, g.c_tot,
                                     st_blk, st));
            }
            for (int l = 0; l < kBlockLayers[b]; ++l) {
                const int layer_index = layer_base + l;
                const DenseLayerW& L = T.layers[b][l];
                const int cin = g.c_in + l * kGrowth;
                double* st_bott = stats_ptr(h, h->st_bott + (size_t)layer_index * kBottleneck) + 2 * (size_t)s0 * kBottleneck;
                float* bott = (save_bott ? h->train.bott_saved[layer_index] : h->bott) + (size_t)s0 * px * kBottleneck;
                ConvArgs a1;
                a1.in = blk; a1.in_cstride = g.c_tot; a1.cin = cin; a1.hin = g.hw;
                a1.in_stats = st_blk; a1.in_stats_stride = g.c_tot;
                a1.gamma = L.norm1.gamma; a1.beta = L.norm1.beta;
                a1.taps = 1; a1.w = &L.conv1;
                a1.out = bott; a1.out_cstride = kBottleneck; a1.out_coff = 0; a1.cout = kBottleneck;
                a1.out_stats = st_bott; a1.out_stats_stride = kBottleneck;
                a1.n = ns;
                SMG_TRY(conv_dispatch(h, a1, st));
                ConvArgs a2;
                a2.in = bott; a2.in_cstride = kBottleneck; a2.cin = kBottleneck; a2.hin = g.hw;
                a2.in_stats = st_bott; a2.in_stats_stride = kBottleneck;
                a2.gamma = L.norm2.gamma; a2.beta = L.norm2.beta;
                a2.taps = 9; a2.w = &L.conv2;
                a2.out = blk; a2.out_cstride = g.c_tot; a2.out_coff = cin; a2.cout = kGrowth;
                a2.out_stats = st_blk; a2.out_stats_stride = g.c_tot;
                a2.n = ns;
                SMG_TRY(conv_dispatch(h, a2, st));
            }
            if (b < kNumBlocks - 1) {
                const TransitionW& R = T.trans[b];
                const BlockGeom& gn = h->geom[b + 1];
                ConvArgs at;
                at.in = blk; at.in_cstride = g.c_tot; at.cin = g.c_tot; at.hin = g.hw;
                at.in_stats = st_blk; at.in_stats_stride = g.c_tot;
                at.gamma = R.norm.gamma; at.beta = R.norm.beta;
                at.pool = 1; at.taps = 1; at.w = &R.conv;
                at.out = h->block[b + 1] + (size_t)s0 * gn.hw * gn.hw * gn.c_tot; at.out_cstride = gn.c_tot; at.out_coff = 0;
                at.cout = g.c_tot / 2;
                at.out_stats = stats_ptr(h, h->st_block[b + 1]) + 2 * (size_t)s0 * gn.c_tot; at.out_stats_stride = gn.c_tot;
                at.n = ns;
                SMG_TRY(conv_dispatch(h, at, st));
            }
        }
        layer_base += kBlockLayers[b];
    }
    h->last_n = n;
    h->fp32_exact_pass = false;
    return SMG_OK;
}

// per-sample BN statistics of all 121 BatchNorm layers in module order: ONE launch over a region table built at creation
int export_bn_stats(smg_handle* h, int n, float* mean, float* var, cudaStream_t st) {
    SMG_CHECK(h->bn_regions_dev != nullptr, SMG_ERR_STATE, "bn export: no region table");
    return launch_bn_export_all(h, n, h->bn_regions_dev, h->bn_regions, SMG_TRUNK_BN_CHANNELS, mean, var, st);
}

static int build_bn_regions(smg_handle* h) {
    std::vector<BnRegion> r;
    int off = 0;
    auto add = [&](size_t st_off, int stride, int count, double cnt) {
        r.push_back(BnRegion{stats_ptr(h, st_off), stride, count, cnt, off});
        off += count;
    };
    add(h->st_conv0, 64, 64, (double)(h->H / 2) * (h->H / 2));
    int layer_index = 0;
    for (int b = 0; b < kNumBlocks; ++b) {
        const BlockGeom& g = h->geom[b];
        const double cnt = (double)g.hw * g.hw;
        for (int l = 0; l < kBlockLayers[b]; ++l, ++layer_index) {
            add(h->st_block[b], g.c_tot, g.c_in + l * kGrowth, cnt);                                      // norm1
            add(h->st_bott + (size_t)layer_index * kBottleneck, kBottleneck, kBottleneck, cnt);           // norm2
        }
        add(h->st_block[b], g.c_tot, g.c_tot, cnt);   // transition norm (b < 3) or norm5 (b == 3): the whole block buffer
    }
    SMG_CHECK(off == SMG_TRUNK_BN_CHANNELS && r.size() == 121, SMG_ERR_STATE, "bn export: %d channels in %zu regions", off, r.size());
    SMG_CUDA(cudaMalloc(&h->bn_regions_dev, r.size() * sizeof(BnRegion)));
    SMG_CUDA(cudaMemcpy(h->bn_regions_dev, r.data(), r.size() * sizeof(BnRegion), cudaMemcpyHostToDevice));
    h->bn_regions = (int)r.size();
    return SMG_OK;
}

// heads for all (mask, rotation) pairs; samples [0,n_rot) are scenes, [n_rot, n_rot+n_masks) masks
int heads_forward(smg_handle* h, int trunk_id, int head_id, int n_rot, int n_masks, float* dev_q, cudaStream_t st,
                  int groups) {
    TrunkW& T = h->trunks[trunk_id];
    HeadW& Hd = h->heads[head_id];
    SMG_CHECK(Hd.set, SMG_ERR_STATE, "head %d: weights not set (call smg_set_head_weights)", head_id);
    h->train.valid = false;
    const BlockGeom& g = h->geom[3];
    SMG_CHECK(g.hw == kHeadK, SMG_ERR_INVALID, "heads need H=640 (block-4 spatial %d != %d)", g.hw, kHeadK);
    SMG_TRY(head_partials(h, trunk_id, head_id, n_rot, n_masks, st, groups));
    const double pairs = (double)groups * n_rot * n_masks;
    // head tail per (mask, rotation) pair: two [400][64] partial products read three times (mean, variance, dot), 25 600 MACs
    ProfScope ps(h, st, 3, pairs * 2.0 * 400 * 64 * Hd.n_out, pairs * 2.0 * 400 * 64 * 4);
    return launch_head_tail(h, h->head_p, h->head_p + (size_t)groups * n_rot * g.hw * g.hw * kHeadMid, n_rot, n_masks, Hd, dev_q, st,
                            groups);
}

// per-sample halves of the head's BN(2048)+ReLU+1x1 conv: channels [0,1024) of the concatenation depend only on the scene
// sample, [1024,2048) only on the mask sample, so P_s = W_half . relu(bn(f_s)) is computed once per sample into h->head_p
// ([groups x n_rot] scenes, then [groups x n_masks] masks; [400][64] each) and paired later by head_tail
int head_partials(smg_handle* h, int trunk_id, int head_id, int n_rot, int n_masks, cudaStream_t st, int groups) {
    TrunkW& T = h->trunks[trunk_id];
    HeadW& Hd = h->heads[head_id];
    const BlockGeom& g = h->geom[3];
    const double* st4 = stats_ptr(h, h->st_block[3]);
    for (int half = 0; half < 2; ++half) {
        const int s0 = half == 0 ? 0 : groups * n_rot;
        const int cnt = groups * (half == 0 ? n_rot : n_masks);
        if (cnt == 0) continue;
        SMG_TRY(launch_head_prepare(h, cnt, st4 + 2 * (size_t)s0 * g.c_tot, g.c_tot, T.norm5, Hd.norm0, half,
                                    h->head_scale + (size_t)s0 * kFeatC, h->head_shift + (size_t)s0 * kFeatC, st));
        ConvArgs a;
        a.in = h->block[3] + (size_t)s0 * g.hw * g.hw * g.c_tot; a.in_cstride = g.c_tot; a.cin = kFeatC; a.hin = g.hw;
        a.prologue_mode = 1;
        a.scale = h->head_scale + (size_t)s0 * kFeatC; a.shift = h->head_shift + (size_t)s0 * kFeatC;
        a.taps = 1; a.w = &Hd.conv0[half];
        a.out = h->head_p + (size_t)s0 * g.hw * g.hw * kHeadMid; a.out_cstride = kHeadMid; a.out_coff = 0; a.cout = kHeadMid;
        a.out_stats = nullptr;
        a.n = cnt;
        SMG_TRY(conv_dispatch(h, a, st));
    }
    return SMG_OK;
}

__global__ void pack_head_conv1_kernel(const float* __restrict__ w, float* __restrict__ out, int n_out, int npix) {
    // torch [n_out][64][npix] -> [n_out][npix][64]
    const int total = n_out * 64 * npix;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int c = i % 64, p = (i / 64) % npix, o = i / (64 * npix);
        out[i] = w[((size_t)o * 64 + c) * npix + p];
    }
}

__global__ void pack_conv0_kernel(const float* __restrict__ w, float* __restrict__ out, float* __restrict__ folded) {
    // torch [64][3][7][7] -> [147][64], and the channel-folded [49][64] (sum over the 3 input channels)
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 147 * 64) {
        const int co = i % 64, k = i / 64;
        out[i] = w[co * 147 + k];
        if (k < 49) folded[i] = (w[co * 147 + k] + w[co * 147 + 49 + k]) + w[co * 147 + 98 + k];
    }
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct ArenaPlanner {
    size_t off = 0;
    size_t take(size_t bytes) {
        const size_t o = off;
        off = align_up(off + bytes, 256);
        return o;
    }
};

static void plan_conv(ArenaPlanner& p, ConvW& cw, int cin, int cout, int taps, uint8_t* base) {
    cw.cin = cin; cw.cout = cout; cw.taps = taps;
    const size_t o1 = p.take(conv_packed_bytes_ffma(cin, cout, taps));
    const size_t o2 = p.take(conv_packed_bytes_umma(cin, cout, taps, 4));
    const size_t o3 = p.take(conv_packed_bytes_umma(cin, cout, taps, 2));
    const size_t o4 = p.take(conv_packed_bytes_ffma(cin, cout, taps));
    const size_t o5 = p.take(conv_packed_bytes_umma(dgrad_cin_padded(cin, taps), cout, taps, 4));
    const size_t o6 = p.take(2 * conv_packed_bytes_umma(cin, cout, taps, 4));
    const bool has_t = taps == 9 || (taps == 1 && cout == 128);
    const size_t o7 = has_t ? p.take(conv_packed_bytes_umma(cin, cout, taps, 4)) : 0;
    if (base) {
        if (has_t) cw.w_tf32_t = base + o7;
        cw.w_dgrad_tf32 = base + o5;
        cw.w_split = base + o6;
        cw.w_ffma = reinterpret_cast<float*>(base + o1);
        cw.w_tf32 = base + o2;
        cw.w_bf16 = base + o3;
        cw.w_dgrad = reinterpret_cast<float*>(base + o4);
    }
}
static void plan_bn(ArenaPlanner& p, BnP& b, int c, uint8_t* base) {
    b.c = c;
    const size_t o1 = p.take((size_t)c * 4), o2 = p.take((size_t)c * 4);
    if (base) {
        b.gamma = reinterpret_cast<float*>(base + o1);
        b.beta = reinterpret_cast<float*>(base + o2);
    }
}

static size_t plan_trunk(smg_handle* h, TrunkW& T, uint8_t* base) {
    ArenaPlanner p;
    const size_t o = p.take(147 * 64 * 4);
    const size_t of = p.take(49 * 64 * 4);
    const size_t ou = p.take(64 * 128 * 4);
    if (base) {
        T.conv0 = reinterpret_cast<float*>(base + o);
        T.conv0_folded = reinterpret_cast<float*>(base + of);
        T.conv0_umma = reinterpret_cast<float*>(base + ou);
    }
    plan_bn(p, T.norm0, 64, base);
    for (int b = 0; b < kNumBlocks; ++b) {
        T.layers[b].resize(kBlockLayers[b]);
        for (int l = 0; l < kBlockLayers[b]; ++l) {
            DenseLayerW& L = T.layers[b][l];
            const int cin = h->geom[b].c_in + l * kGrowth;
            plan_bn(p, L.norm1, cin, base);
            plan_conv(p, L.conv1, cin, kBottleneck, 1, base);
            plan_bn(p, L.norm2, kBottleneck, base);
            plan_conv(p, L.conv2, kBottleneck, kGrowth, 9, base);
        }
        if (b < kNumBlocks - 1) {
            plan_bn(p, T.trans[b].norm, h->geom[b].c_tot, base);
            plan_conv(p, T.trans[b].conv, h->geom[b].c_tot, h->geom[b].c_tot / 2, 1, base);
        }
    }
    plan_bn(p, T.norm5, kFeatC, base);
    return p.off;
}

int repack_trunk(smg_handle* h, int trunk_id, cudaStream_t st) {
    TrunkW& T = h->trunks[trunk_id];
    SMG_CHECK(T.jobs_dev != nullptr && T.src.size() == SMG_TRUNK_NUM_PARAMS, SMG_ERR_STATE, "trunk %d: no packing tables", trunk_id);
    pack_conv0_kernel<<<(147 * 64 + 255) / 256, 256, 0, st>>>(T.src[0], T.conv0, T.conv0_folded);
    h->launches++;
    SMG_TRY(pack_conv0_umma(h, T.conv0_folded, T.conv0_umma, st));
    const PackJob* pj = reinterpret_cast<const PackJob*>(T.jobs_dev);
    const CopyJob* cj = reinterpret_cast<const CopyJob*>(reinterpret_cast<const uint8_t*>(T.jobs_dev) + T.pack_jobs.size() * sizeof(PackJob));
    SMG_TRY(launch_pack_tables(h, pj, (int)T.pack_jobs.size(), cj, (int)T.copy_jobs.size(), st));
    return SMG_OK;
}

int repack_head(smg_handle* h, int head_id, cudaStream_t st) {
    HeadW& Hd = h->heads[head_id];
    SMG_CHECK(Hd.src.size() == SMG_HEAD_NUM_PARAMS && Hd.arena != nullptr, SMG_ERR_STATE, "head %d: weights never set", head_id);
    const float* const* dev_params = Hd.src.data();
    SMG_CUDA(cudaMemcpyAsync(Hd.norm0.gamma, dev_params[0], 2 * kFeatC * 4, cudaMemcpyDeviceToDevice, st));
    SMG_CUDA(cudaMemcpyAsync(Hd.norm0.beta, dev_params[1], 2 * kFeatC * 4, cudaMemcpyDeviceToDevice, st));
    SMG_TRY(pack_conv_weights(h, dev_params[2], Hd.conv0[0], 0, 2 * kFeatC, st));
    SMG_TRY(pack_conv_weights(h, dev_params[2], Hd.conv0[1], kFeatC, 2 * kFeatC, st));
    SMG_CUDA(cudaMemcpyAsync(Hd.norm1.gamma, dev_params[3], kHeadMid * 4, cudaMemcpyDeviceToDevice, st));
    SMG_CUDA(cudaMemcpyAsync(Hd.norm1.beta, dev_params[4], kHeadMid * 4, cudaMemcpyDeviceToDevice, st));
    pack_head_conv1_kernel<<<64, 256, 0, st>>>(dev_params[5], Hd.conv1, Hd.n_out, kHeadK * kHeadK);
    h->launches++;
    SMG_CUDA(cudaGetLastError());
    return SMG_OK;
}

}  // namespace smg

using namespace smg;

extern "C" {

int smg_version(void) { return 100; }
const char* smg_last_error(void) { return smg::get_error(); }

int smg_create(int device, int max_samples, int H, smg_handle** out) {
    SMG_CHECK(out != nullptr, SMG_ERR_INVALID, "smg_create: out is NULL");
    SMG_CHECK(max_samples >= 0 && max_samples <= 4096, SMG_ERR_INVALID, "smg_create: max_samples %d", max_samples);
    SMG_CHECK(H >= 64 && H % 32 == 0 && H <= 1024, SMG_ERR_INVALID, "smg_create: H=%d must be a multiple of 32 in [64,1024]", H);
    int ndev = 0;
    SMG_CUDA(cudaGetDeviceCount(&ndev));
    SMG_CHECK(device >= 0 && device < ndev, SMG_ERR_INVALID, "smg_create: device %d of %d", device, ndev);
    DeviceGuard guard(device);
    cudaDeviceProp prop;
    SMG_CUDA(cudaGetDeviceProperties(&prop, device));
    SMG_CHECK(prop.major == 10, SMG_ERR_UNSUPPORTED,
              "smg_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
    smg_handle* h = new smg_handle();
    h->device = device;
    h->max_samples = max_samples;
    h->H = H;
    h->num_sms = prop.multiProcessorCount;
    if (const char* e = getenv("SMG_L2_CHUNK_MB")) {  // tuning knob: 0 disables the chunked schedule
        const double mb = atof(e);
        h->l2_chunk_bytes = mb > 0 ? mb * 1e6 : 0.0;
    }
    int c = kInitFeatures, hw = H / 4;
    for (int b = 0; b < kNumBlocks; ++b) {
        h->geom[b].hw = hw;
        h->geom[b].c_in = c;
        h->geom[b].c_tot = c + kBlockLayers[b] * kGrowth;
        c = h->geom[b].c_tot / 2;
        hw /= 2;
    }
    if (max_samples == 0) {
        // a handle for the stateless kernels only (heightmap, NMS, argmax, Adam): no trunk workspace
        *out = h;
        return SMG_OK;
    }
    const size_t S = max_samples;
    // stats arena layout (double2 per sample)
    size_t off = 0;
    h->st_conv0 = off; off += 64;
    for (int b = 0; b < kNumBlocks; ++b) { h->st_block[b] = off; off += h->geom[b].c_tot; }
    h->st_bott = off; off += (size_t)58 * kBottleneck;
    h->stats_doubles_per_sample = 2 * off;
    h->stats_bytes = S * 2 * off * sizeof(double);

    ArenaPlanner p;
    const size_t o_in = p.take(S * 3 * H * H * 4);
    const size_t o_c0 = p.take(S * (size_t)(H / 2) * (H / 2) * 64 * 4);
    size_t o_blk[kNumBlocks];
    for (int b = 0; b < kNumBlocks; ++b) o_blk[b] = p.take(S * (size_t)h->geom[b].hw * h->geom[b].hw * h->geom[b].c_tot * 4);
    const size_t o_bott = p.take(S * (size_t)(H / 4) * (H / 4) * kBottleneck * 4);
    const size_t o_stats = p.take(h->stats_bytes);
    const size_t o_hs = p.take(S * kFeatC * 4), o_hh = p.take(S * kFeatC * 4);
    const size_t o_hp = p.take(S * (size_t)h->geom[3].hw * h->geom[3].hw * kHeadMid * 4);
    const size_t o_tmp = p.take((S > 3 ? S : 3) * (size_t)H * H * 4);
    const size_t o_hm = p.take((1 + S) * (size_t)(H / 2) * (H / 2) * 8);
    const size_t o_hb = p.take(S * S * 128 * 4);
    const size_t o_q = p.take(S * S * 4 * 4);
    uint8_t* base = nullptr;
    cudaError_t e = cudaMalloc(&base, p.off);
    if (e != cudaSuccess) {
        set_error("smg_create: cudaMalloc(%zu bytes) failed: %s", p.off, cudaGetErrorString(e));
        delete h;
        return SMG_ERR_CUDA;
    }
    h->workspace_bytes = (int64_t)p.off;
    h->input = reinterpret_cast<float*>(base + o_in);
    h->conv0 = reinterpret_cast<float*>(base + o_c0);
    for (int b = 0; b < kNumBlocks; ++b) h->block[b] = reinterpret_cast<float*>(base + o_blk[b]);
    h->bott = reinterpret_cast<float*>(base + o_bott);
    h->stats = reinterpret_cast<double*>(base + o_stats);
    h->head_scale = reinterpret_cast<float*>(base + o_hs);
    h->head_shift = reinterpret_cast<float*>(base + o_hh);
    h->head_p = reinterpret_cast<float*>(base + o_hp);
    h->scene_tmp = reinterpret_cast<float*>(base + o_tmp);
    h->hm_stage = reinterpret_cast<double*>(base + o_hm);
    h->head_bn1 = reinterpret_cast<float*>(base + o_hb);
    h->head_bn1_floats = S * S * 128;
    h->q_stage = reinterpret_cast<float*>(base + o_q);
    if (build_bn_regions(h) != SMG_OK) {
        cudaFree(base);
        delete h;
        return SMG_ERR_CUDA;
    }
    {
        // the graph stream carries the critical chain: highest priority, so that its CTAs are scheduled ahead of the
        // side-stream weight gradients whenever SMs free up
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        cudaStreamCreateWithPriority(&h->gstream, cudaStreamNonBlocking, hi);
    }
    cudaEventCreateWithFlags(&h->g_in, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->g_out, cudaEventDisableTiming);
    if (const char* e = getenv("SMG_NO_GRAPHS")) h->use_graphs = atoi(e) == 0;
    if (const char* e = getenv("SMG_ASYNC")) h->force_async = atoi(e);
    if (const char* e = getenv("SMG_TMA")) h->use_tma = atoi(e);
    if (const char* e = getenv("SMG_FP32_TC")) h->fp32_tc = atoi(e) != 0;
    if (const char* e = getenv("SMG_PDL")) h->use_pdl = atoi(e) != 0;
    *out = h;
    return SMG_OK;
}

int smg_destroy(smg_handle* h) {
    if (!h) return SMG_OK;
    DeviceGuard guard(h->device);
    cudaDeviceSynchronize();
    for (auto& g : h->graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    for (auto& g : h->step.graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    if (h->step.tables) cudaFree(h->step.tables);
    if (h->step.adam_tables) cudaFree(h->step.adam_tables);
    if (h->bn_regions_dev) cudaFree(h->bn_regions_dev);
    if (h->geo_out) cudaFree(h->geo_out);
    if (h->bn_stage) cudaFree(h->bn_stage);
    if (h->train.arena) cudaFree(h->train.arena);
    if (h->train.wstream) cudaStreamDestroy(h->train.wstream);
    for (auto& e : h->train.ev)
        if (e) cudaEventDestroy(e);
    if (h->gstream) cudaStreamDestroy(h->gstream);
    if (h->g_in) cudaEventDestroy(h->g_in);
    if (h->g_out) cudaEventDestroy(h->g_out);
    if (h->input) cudaFree(h->input);  // base of the workspace arena
    for (int t = 0; t < SMG_NUM_TRUNKS; ++t) {
        if (h->trunks[t].arena) cudaFree(h->trunks[t].arena);
        if (h->trunks[t].jobs_dev) cudaFree(h->trunks[t].jobs_dev);
    }
    for (int t = 0; t < SMG_NUM_HEADS; ++t)
        if (h->heads[t].arena) cudaFree(h->heads[t].arena);
    delete h;
    return SMG_OK;
}

int smg_set_precision(smg_handle* h, int precision) {
    SMG_CHECK(h != nullptr, SMG_ERR_INVALID, "NULL handle");
    SMG_CHECK(precision >= SMG_PREC_FP32 && precision <= SMG_PREC_BF16, SMG_ERR_INVALID, "precision %d", precision);
    h->precision = precision;
    return SMG_OK;
}
int smg_get_precision(smg_handle* h) { return h ? h->precision : SMG_ERR_INVALID; }
int64_t smg_workspace_bytes(smg_handle* h) { return h ? h->workspace_bytes : 0; }
int64_t smg_launch_count(smg_handle* h) { return h ? h->launches : 0; }

int smg_set_trunk_weights(smg_handle* h, int trunk_id, const float* const* dev_params, int n, void* stream) {
    SMG_CHECK(h != nullptr && dev_params != nullptr, SMG_ERR_INVALID, "NULL argument");
    SMG_CHECK(trunk_id >= 0 && trunk_id < SMG_NUM_TRUNKS, SMG_ERR_INVALID, "trunk_id %d", trunk_id);
    SMG_CHECK(n == SMG_TRUNK_NUM_PARAMS, SMG_ERR_INVALID, "expected %d trunk tensors, got %d", SMG_TRUNK_NUM_PARAMS, n);
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    TrunkW& T = h->trunks[trunk_id];
    if (!T.arena) {
        T.arena_bytes = plan_trunk(h, T, nullptr);
        SMG_CUDA(cudaMalloc(&T.arena, T.arena_bytes));
        plan_trunk(h, T, reinterpret_cast<uint8_t*>(T.arena));
    }
    // the job tables live in the trunk (host copy + device copy): they are rebuilt only when a source pointer changes, so
    // that repack_trunk() can re-run the packing kernels alone - e.g. inside the captured training step
    std::vector<const float*> src(dev_params, dev_params + n);
    int i = SMG_TRUNK_NUM_PARAMS;
    if (src != T.src) {
        i = 0;
        T.pack_jobs.clear();
        T.copy_jobs.clear();
        auto copy_bn = [&](BnP& b) {
            T.copy_jobs.push_back(CopyJob{dev_params[i++], b.gamma, b.c});
            T.copy_jobs.push_back(CopyJob{dev_params[i++], b.beta, b.c});
        };
        i++;  // conv0.weight: packed by its own two kernels
        copy_bn(T.norm0);
        for (int b = 0; b < kNumBlocks; ++b) {
            for (int l = 0; l < kBlockLayers[b]; ++l) {
                DenseLayerW& L = T.layers[b][l];
                copy_bn(L.norm1);
                T.pack_jobs.push_back(make_pack_job(dev_params[i++], L.conv1, 0, L.conv1.cin));
                copy_bn(L.norm2);
                T.pack_jobs.push_back(make_pack_job(dev_params[i++], L.conv2, 0, L.conv2.cin));
            }
            if (b < kNumBlocks - 1) {
                copy_bn(T.trans[b].norm);
                T.pack_jobs.push_back(make_pack_job(dev_params[i++], T.trans[b].conv, 0, T.trans[b].conv.cin));
            }
        }
        copy_bn(T.norm5);
        const size_t pb = T.pack_jobs.size() * sizeof(PackJob), cb = T.copy_jobs.size() * sizeof(CopyJob);
        if (!T.jobs_dev) SMG_CUDA(cudaMalloc(&T.jobs_dev, pb + cb));
        SMG_CUDA(cudaMemcpy(T.jobs_dev, T.pack_jobs.data(), pb, cudaMemcpyHostToDevice));
        SMG_CUDA(cudaMemcpy(reinterpret_cast<uint8_t*>(T.jobs_dev) + pb, T.copy_jobs.data(), cb, cudaMemcpyHostToDevice));
        T.src = src;
    }
    SMG_TRY(repack_trunk(h, trunk_id, st));
    SMG_CHECK(i == SMG_TRUNK_NUM_PARAMS, SMG_ERR_STATE, "consumed %d trunk tensors", i);
    SMG_CUDA(cudaGetLastError());
    T.set = true;
    T.packed = h->pack_mask;
    return SMG_OK;
}

int smg_set_head_weights(smg_handle* h, int head_id, const float* const* dev_params, int n_out, void* stream) {
    SMG_CHECK(h != nullptr && dev_params != nullptr, SMG_ERR_INVALID, "NULL argument");
    SMG_CHECK(head_id >= 0 && head_id < SMG_NUM_HEADS, SMG_ERR_INVALID, "head_id %d", head_id);
    SMG_CHECK(n_out >= 1 && n_out <= 4, SMG_ERR_INVALID, "n_out %d", n_out);
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    HeadW& Hd = h->heads[head_id];
    const int npix = kHeadK * kHeadK;
    if (!Hd.arena || Hd.n_out != n_out) {
        if (Hd.arena) { cudaFree(Hd.arena); Hd.arena = nullptr; }
        for (int pass = 0; pass < 2; ++pass) {
            ArenaPlanner p;
            uint8_t* base = reinterpret_cast<uint8_t*>(Hd.arena);
            plan_bn(p, Hd.norm0, 2 * kFeatC, base);
            plan_conv(p, Hd.conv0[0], kFeatC, kHeadMid, 1, base);
            plan_conv(p, Hd.conv0[1], kFeatC, kHeadMid, 1, base);
            plan_bn(p, Hd.norm1, kHeadMid, base);
            const size_t o = p.take((size_t)n_out * npix * kHeadMid * 4);
            if (base) Hd.conv1 = reinterpret_cast<float*>(base + o);
            if (pass == 0) {
                Hd.arena_bytes = p.off;
                SMG_CUDA(cudaMalloc(&Hd.arena, Hd.arena_bytes));
            }
        }
        Hd.n_out = n_out;
    }
    Hd.src.assign(dev_params, dev_params + SMG_HEAD_NUM_PARAMS);
    SMG_TRY(repack_head(h, head_id, st));
    Hd.set = true;
    Hd.packed = SMG_PACK_ALL;  // the head's few tensors are always packed in every layout
    return SMG_OK;
}

int smg_prep(smg_handle* h, const double* dev_heightmaps, int n, int hm_size, double mean, double stddev,
             float* dev_out, void* stream) {
    SMG_CHECK(h && dev_heightmaps && dev_out && n >= 1, SMG_ERR_INVALID, "smg_prep: bad argument");
    SMG_CHECK(stddev != 0.0, SMG_ERR_INVALID, "smg_prep: stddev is 0 (the reference's published literal gives NaN)");
    DeviceGuard guard(h->device);
    return launch_prep(h, dev_heightmaps, n, hm_size, mean, stddev, dev_out, 3, (cudaStream_t)stream);
}

int smg_rotate(smg_handle* h, const float* dev_in, const int* host_rot_idx, int n_rot, int num_rotations,
               float* dev_out, void* stream) {
    SMG_CHECK(h && dev_in && dev_out && host_rot_idx && n_rot >= 1 && num_rotations >= 1, SMG_ERR_INVALID,
              "smg_rotate: bad argument");
    DeviceGuard guard(h->device);
    return launch_rotate(h, dev_in, host_rot_idx, n_rot, num_rotations, dev_out, 3, (cudaStream_t)stream);
}

int smg_rotate_index_map(smg_handle* h, int rot_idx, int num_rotations, int32_t* dev_out, void* stream) {
    SMG_CHECK(h && dev_out && num_rotations >= 1, SMG_ERR_INVALID, "smg_rotate_index_map: bad argument");
    DeviceGuard guard(h->device);
    return launch_rotate_index_map(h, rot_idx, num_rotations, dev_out, (cudaStream_t)stream);
}

int smg_trunk_forward(smg_handle* h, int trunk_id, const float* dev_in, int n, float* dev_feat, float* dev_bn_mean,
                      float* dev_bn_var, void* stream) {
    SMG_CHECK(h && dev_in, SMG_ERR_INVALID, "smg_trunk_forward: bad argument");
    SMG_CHECK(trunk_id >= 0 && trunk_id < SMG_NUM_TRUNKS, SMG_ERR_INVALID, "trunk_id %d", trunk_id);
    SMG_CHECK(n >= 1 && n <= h->max_samples, SMG_ERR_INVALID, "n=%d outside [1,%d]", n, h->max_samples);
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    SMG_CUDA(cudaMemcpyAsync(h->input, dev_in, (size_t)n * 3 * h->H * h->H * 4, cudaMemcpyDeviceToDevice, st));
    SMG_TRY(trunk_forward(h, trunk_id, n, 3, st));
    if (dev_feat)
        SMG_TRY(launch_norm5_export(h, n, h->block[3], stats_ptr(h, h->st_block[3]), h->geom[3].c_tot,
                                    h->trunks[trunk_id].norm5, dev_feat, st));
    if (dev_bn_mean && dev_bn_var) SMG_TRY(export_bn_stats(h, n, dev_bn_mean, dev_bn_var, st));
    return SMG_OK;
}

static int qforward_common(smg_handle* h, int trunk_id, int head_id, int n_masks, int n_rot, int in_channels,
                           float* dev_q, float* dev_bn_mean, float* dev_bn_var, cudaStream_t st, int groups = 1) {
    SMG_TRY(trunk_forward(h, trunk_id, groups * (n_rot + n_masks), in_channels, st));
    SMG_TRY(heads_forward(h, trunk_id, head_id, n_rot, n_masks, dev_q, st, groups));
    if (dev_bn_mean && dev_bn_var) SMG_TRY(export_bn_stats(h, groups * (n_rot + n_masks), dev_bn_mean, dev_bn_var, st));
    return SMG_OK;
}

int smg_qforward(smg_handle* h, int trunk_id, int head_id, const float* dev_scene, const float* dev_masks,
                 int n_masks, const int* host_rot_idx, int n_rot, int num_rotations, float* dev_q, float* dev_bn_mean,
                 float* dev_bn_var, void* stream) {
    SMG_CHECK(h && dev_scene && dev_masks && host_rot_idx && dev_q, SMG_ERR_INVALID, "smg_qforward: NULL argument");
    SMG_CHECK(trunk_id >= 0 && trunk_id < SMG_NUM_TRUNKS && head_id >= 0 && head_id < SMG_NUM_HEADS, SMG_ERR_INVALID,
              "smg_qforward: trunk %d / head %d", trunk_id, head_id);
    SMG_CHECK(n_masks >= 1 && n_rot >= 1 && n_masks + n_rot <= h->max_samples, SMG_ERR_INVALID,
              "smg_qforward: %d rotations + %d masks exceed max_samples %d", n_rot, n_masks, h->max_samples);
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t img = (size_t)3 * h->H * h->H;
    SMG_TRY(launch_rotate(h, dev_scene, host_rot_idx, n_rot, num_rotations, h->input, 3, st));
    SMG_CUDA(cudaMemcpyAsync(h->input + (size_t)n_rot * img, dev_masks, (size_t)n_masks * img * 4, cudaMemcpyDeviceToDevice, st));
    return qforward_common(h, trunk_id, head_id, n_masks, n_rot, 3, dev_q, dev_bn_mean, dev_bn_var, st);
}

// `groups` independent units (scene g with its n_masks masked scenes) evaluated as ONE batch: samples are laid out as
// [groups x n_rot] rotated scenes followed by [groups x n_masks] masked scenes, so the trunk and the head GEMMs see
// one long sample axis (more CTAs per launch for the small late layers) and only the pairing in head_tail is per group.
static int qforward_maps_body(smg_handle* h, int trunk_id, int head_id, const double* dev_scene_hm, const double* dev_mask_hms,
                              int n_masks, int hm_size, double mean, double stddev, const int* host_rot_idx, int n_rot,
                              int num_rotations, float* dev_q, float* dev_bn_mean, float* dev_bn_var, cudaStream_t st,
                              int groups = 1) {
    // Trainer.forward feeds three identical channels (code/trainer.py:178-181): keep ONE plane per sample and use
    // the channel-folded conv0 weights (K = 49 instead of 147)
    {
        const double n_px = (double)groups * (n_rot + n_masks) * h->H * h->H;
        ProfScope ps(h, st, 3, 0.0, 4.0 * n_px);   // K1: one float written per network-input pixel (the heightmaps stay in L2)
        if (n_rot <= 32) {
            SMG_TRY(launch_prep_rotate(h, dev_scene_hm, groups, host_rot_idx, n_rot, num_rotations, dev_mask_hms, groups * n_masks,
                                       hm_size, mean, stddev, h->input, st));
        } else {
            const size_t img = (size_t)h->H * h->H;
            SMG_TRY(launch_prep(h, dev_scene_hm, groups, hm_size, mean, stddev, h->scene_tmp, 1, st));
            for (int g = 0; g < groups; ++g)
                SMG_TRY(launch_rotate(h, h->scene_tmp + (size_t)g * img, host_rot_idx, n_rot, num_rotations,
                                      h->input + (size_t)g * n_rot * img, 1, st));
            SMG_TRY(launch_prep(h, dev_mask_hms, groups * n_masks, hm_size, mean, stddev, h->input + (size_t)groups * n_rot * img, 1, st));
        }
    }
    return qforward_common(h, trunk_id, head_id, n_masks, n_rot, 1, dev_q, dev_bn_mean, dev_bn_var, st, groups);
}

int smg_qpartials(smg_handle* h, int trunk_id, int head_id, const double* dev_scene_hm, const int* host_rot_idx, int n_rot,
                  int num_rotations, const double* dev_mask_hms, int n_masks, int hm_size, double mean, double stddev,
                  float* dev_p, void* stream) {
    SMG_CHECK(h && dev_p && (n_rot == 0 || (dev_scene_hm && host_rot_idx)) && (n_masks == 0 || dev_mask_hms), SMG_ERR_INVALID,
              "smg_qpartials: NULL argument");
    SMG_CHECK(trunk_id >= 0 && trunk_id < SMG_NUM_TRUNKS && head_id >= 0 && head_id < SMG_NUM_HEADS, SMG_ERR_INVALID,
              "smg_qpartials: trunk %d / head %d", trunk_id, head_id);
    SMG_CHECK(n_rot >= 0 && n_masks >= 0 && n_rot + n_masks >= 1 && n_rot + n_masks <= h->max_samples, SMG_ERR_INVALID,
              "smg_qpartials: %d rotations + %d masks outside [1, %d]", n_rot, n_masks, h->max_samples);
    SMG_CHECK(stddev != 0.0 && 2 * hm_size <= h->H, SMG_ERR_INVALID, "smg_qpartials: stddev %g / hm_size %d", stddev, hm_size);
    SMG_CHECK(h->heads[head_id].set, SMG_ERR_STATE, "head %d: weights not set", head_id);
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t img = (size_t)h->H * h->H;
    if (n_rot > 0) {
        SMG_TRY(launch_prep(h, dev_scene_hm, 1, hm_size, mean, stddev, h->scene_tmp, 1, st));
        SMG_TRY(launch_rotate(h, h->scene_tmp, host_rot_idx, n_rot, num_rotations, h->input, 1, st));
    }
    if (n_masks > 0) SMG_TRY(launch_prep(h, dev_mask_hms, n_masks, hm_size, mean, stddev, h->input + (size_t)n_rot * img, 1, st));
    SMG_TRY(trunk_forward(h, trunk_id, n_rot + n_masks, 1, st));
    SMG_TRY(head_partials(h, trunk_id, head_id, n_rot, n_masks, st, 1));
    const size_t per = (size_t)kHeadK * kHeadK * kHeadMid;
    SMG_CUDA(cudaMemcpyAsync(dev_p, h->head_p, (size_t)(n_rot + n_masks) * per * 4, cudaMemcpyDeviceToDevice, st));
    return SMG_OK;
}

int smg_qcombine(smg_handle* h, int head_id, const float* dev_p_scene, int n_rot, const float* dev_p_mask, int n_masks,
                 float* dev_q, void* stream) {
    SMG_CHECK(h && dev_p_scene && dev_p_mask && dev_q && n_rot >= 1 && n_masks >= 1, SMG_ERR_INVALID, "smg_qcombine: bad argument");
    SMG_CHECK(head_id >= 0 && head_id < SMG_NUM_HEADS && h->heads[head_id].set, SMG_ERR_STATE, "smg_qcombine: head %d not set", head_id);
    DeviceGuard guard(h->device);
    return launch_head_tail(h, dev_p_scene, dev_p_mask, n_rot, n_masks, h->heads[head_id], dev_q, (cudaStream_t)stream, 1);
}

int smg_qforward_maps(smg_handle* h, int trunk_id, int head_id, const double* dev_scene_hm, const double* dev_mask_hms,
                      int n_masks, int hm_size, double mean, double stddev, const int* host_rot_idx, int n_rot,
                      int num_rotations, float* dev_q, float* dev_bn_mean, float* dev_bn_var, void* stream) {
    return smg_qforward_maps_batch(h, trunk_id, head_id, dev_scene_hm, dev_mask_hms, 1, n_masks, hm_size, mean, stddev,
                                   host_rot_idx, n_rot, num_rotations, dev_q, dev_bn_mean, dev_bn_var, stream);
}

int smg_qforward_maps_batch(smg_handle* h, int trunk_id, int head_id, const double* dev_scene_hm, const double* dev_mask_hms,
                            int groups, int n_masks, int hm_size, double mean, double stddev, const int* host_rot_idx,
                            int n_rot, int num_rotations, float* dev_q, float* dev_bn_mean, float* dev_bn_var, void* stream) {
    SMG_CHECK(h && dev_scene_hm && dev_mask_hms && host_rot_idx && dev_q, SMG_ERR_INVALID, "smg_qforward_maps: NULL argument");
    SMG_CHECK(trunk_id >= 0 && trunk_id < SMG_NUM_TRUNKS && head_id >= 0 && head_id < SMG_NUM_HEADS, SMG_ERR_INVALID,
              "smg_qforward_maps: trunk %d / head %d", trunk_id, head_id);
    SMG_CHECK(groups >= 1 && n_masks >= 1 && n_rot >= 1 && groups * (n_masks + n_rot) <= h->max_samples, SMG_ERR_INVALID,
              "smg_qforward_maps: %d x (%d rotations + %d masks) exceed max_samples %d", groups, n_rot, n_masks, h->max_samples);
    SMG_CHECK(stddev != 0.0, SMG_ERR_INVALID, "smg_qforward_maps: stddev is 0");
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t hm_elems = (size_t)hm_size * hm_size;
    SMG_CHECK(2 * hm_size <= h->H, SMG_ERR_INVALID, "smg_qforward_maps: hm_size %d too large for H %d", hm_size, h->H);
    h->train.valid = false;   // also on the graph-replay path, which does not go through trunk_forward
    const bool want_stats = dev_bn_mean != nullptr && dev_bn_var != nullptr;
    const bool graphable = h->use_graphs && !h->profile && (want_stats || (!dev_bn_mean && !dev_bn_var));
    if (!graphable) return qforward_maps_body(h, trunk_id, head_id, dev_scene_hm, dev_mask_hms, n_masks, hm_size, mean, stddev,
                                              host_rot_idx, n_rot, num_rotations, dev_q, dev_bn_mean, dev_bn_var, st, groups);
    // the BatchNorm running-statistics side effect needs the per-sample batch statistics of the pass: exported inside the
    // captured graph into a staging buffer of the handle (allocated on first use, outside any capture), copied out below
    const size_t bn_floats = (size_t)h->max_samples * SMG_TRUNK_BN_CHANNELS;
    if (want_stats && h->bn_stage == nullptr) SMG_CUDA(cudaMalloc(&h->bn_stage, 2 * bn_floats * sizeof(float)));
    float* st_mean = want_stats ? h->bn_stage : nullptr;
    float* st_var = want_stats ? h->bn_stage + bn_floats : nullptr;
    const size_t bn_bytes = (size_t)groups * (n_rot + n_masks) * SMG_TRUNK_BN_CHANNELS * sizeof(float);
    // stage the inputs at fixed addresses: [groups] scenes, then [groups x n_masks] masks
    double* stage_masks = h->hm_stage + (size_t)groups * hm_elems;
    SMG_CUDA(cudaMemcpyAsync(h->hm_stage, dev_scene_hm, (size_t)groups * hm_elems * 8, cudaMemcpyDeviceToDevice, st));
    SMG_CUDA(cudaMemcpyAsync(stage_masks, dev_mask_hms, (size_t)groups * n_masks * hm_elems * 8, cudaMemcpyDeviceToDevice, st));
    smg_handle::QGraph* G = nullptr;
    for (auto& g : h->graphs)
        if (g.trunk_id == trunk_id && g.head_id == head_id && g.n_masks == n_masks && g.n_rot == n_rot && g.groups == groups &&
            g.num_rot == num_rotations && g.hm_size == hm_size && g.precision == h->precision && g.mean == mean &&
            g.stddev == stddev && g.stats == (want_stats ? 1 : 0) && g.rots == std::vector<int>(host_rot_idx, host_rot_idx + n_rot)) {
            G = &g;
            break;
        }
    if (!G) {
        smg_handle::QGraph g;
        g.trunk_id = trunk_id; g.head_id = head_id; g.n_masks = n_masks; g.n_rot = n_rot; g.num_rot = num_rotations; g.groups = groups;
        g.hm_size = hm_size; g.precision = h->precision; g.mean = mean; g.stddev = stddev; g.stats = want_stats ? 1 : 0;
        g.rots.assign(host_rot_idx, host_rot_idx + n_rot);
        h->graphs.push_back(g);
        G = &h->graphs.back();
    }
    const size_t q_bytes = (size_t)groups * n_masks * n_rot * h->heads[head_id].n_out * 4;
    if (G->seen == 0 || h->graphs.size() > 64) {
        // first sighting: run eagerly (also performs the one-time cudaFuncSetAttribute calls)
        G->seen = 1;
        SMG_TRY(qforward_maps_body(h, trunk_id, head_id, h->hm_stage, stage_masks, n_masks, hm_size, mean, stddev,
                                   host_rot_idx, n_rot, num_rotations, h->q_stage, st_mean, st_var, st, groups));
        SMG_CUDA(cudaMemcpyAsync(dev_q, h->q_stage, q_bytes, cudaMemcpyDeviceToDevice, st));
        if (want_stats) {
            SMG_CUDA(cudaMemcpyAsync(dev_bn_mean, st_mean, bn_bytes, cudaMemcpyDeviceToDevice, st));
            SMG_CUDA(cudaMemcpyAsync(dev_bn_var, st_var, bn_bytes, cudaMemcpyDeviceToDevice, st));
        }
        return SMG_OK;
    }
    SMG_CUDA(cudaEventRecord(h->g_in, st));
    SMG_CUDA(cudaStreamWaitEvent(h->gstream, h->g_in, 0));
    if (!G->exec) {
        cudaGraph_t graph = nullptr;
        const int64_t launches_before = h->launches;
        SMG_CUDA(cudaStreamBeginCapture(h->gstream, cudaStreamCaptureModeRelaxed));
        const int status = qforward_maps_body(h, trunk_id, head_id, h->hm_stage, stage_masks, n_masks, hm_size, mean,
                                              stddev, host_rot_idx, n_rot, num_rotations, h->q_stage, st_mean, st_var, h->gstream,
                                              groups);
        cudaError_t e = cudaStreamEndCapture(h->gstream, &graph);
        const int64_t captured = h->launches - launches_before;
        h->launches = launches_before;  // capturing enqueues nothing
        if (status != SMG_OK || e != cudaSuccess) {
            if (graph) cudaGraphDestroy(graph);
            if (status == SMG_OK) set_error("smg_qforward_maps: graph capture failed: %s", cudaGetErrorString(e));
            return status != SMG_OK ? status : SMG_ERR_CUDA;
        }
        e = cudaGraphInstantiate(&G->exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) {
            set_error("smg_qforward_maps: cudaGraphInstantiate: %s", cudaGetErrorString(e));
            return SMG_ERR_CUDA;
        }
        G->n_launches = captured;
        G->head_pairs = h->head_bn1_pairs;
        G->seen = 2;
    }
    h->head_bn1_pairs = G->head_pairs;   // launch_head_tail's host-side bookkeeping does not run on a replay
    SMG_CUDA(cudaGraphLaunch(G->exec, h->gstream));
    h->launches += G->n_launches;
    SMG_CUDA(cudaEventRecord(h->g_out, h->gstream));
    SMG_CUDA(cudaStreamWaitEvent(st, h->g_out, 0));
    SMG_CUDA(cudaMemcpyAsync(dev_q, h->q_stage, q_bytes, cudaMemcpyDeviceToDevice, st));
    if (want_stats) {
        SMG_CUDA(cudaMemcpyAsync(dev_bn_mean, st_mean, bn_bytes, cudaMemcpyDeviceToDevice, st));
        SMG_CUDA(cudaMemcpyAsync(dev_bn_var, st_var, bn_bytes, cudaMemcpyDeviceToDevice, st));
    }
    return SMG_OK;
}

}  // extern "C" (reopened below)

namespace smg {

// ---------------------------------------------------------------------------------------
// training: forward that keeps what the backward needs, backward, Adam
// ---------------------------------------------------------------------------------------
int ensure_train_workspace(smg_handle* h) {
    smg_handle::TrainWs& W = h->train;
    if (W.arena) return SMG_OK;
    const size_t S = 2;
    for (int pass = 0; pass < 2; ++pass) {
        ArenaPlanner p;
        uint8_t* base = reinterpret_cast<uint8_t*>(W.arena);
        int li = 0;
        for (int b = 0; b < kNumBlocks; ++b)
            for (int l = 0; l < kBlockLayers[b]; ++l, ++li) {
                const size_t o = p.take(S * h->geom[b].hw * h->geom[b].hw * kBottleneck * 4);
                if (base) W.bott_saved[li] = reinterpret_cast<float*>(base + o);
            }
        li = 0;
        for (int b = 0; b < kNumBlocks; ++b)
            for (int l = 0; l < kBlockLayers[b]; ++l, ++li) {
                const size_t o = p.take(S * h->geom[b].hw * h->geom[b].hw * kBottleneck * 4);
                if (base) W.dy1_saved[li] = reinterpret_cast<float*>(base + o);
            }
        for (int b = 0; b < kNumBlocks; ++b) {
            const size_t o = p.take(S * h->geom[b].hw * h->geom[b].hw * h->geom[b].c_tot * 4);
            if (base) W.dblk[b] = reinterpret_cast<float*>(base + o);
        }
        const size_t hc = (size_t)(h->H / 2) * (h->H / 2), hq = (size_t)(h->H / 4) * (h->H / 4);
        const size_t o_c0 = p.take(S * hc * 64 * 4);
        const size_t o_a = p.take(S * hq * kBottleneck * 4), o_b = p.take(S * hq * kBottleneck * 4);
        const size_t o_c = p.take(S * hq * h->geom[0].c_tot * 4);
        const size_t o_p = p.take((size_t)kHeadK * kHeadK * kHeadMid * 4);
        const size_t o_w3 = p.take((size_t)58 * 288 * 128 * 4);
        const size_t o_wj = p.take((size_t)58 * 2 * sizeof(void*));
        W.sums_bytes = S * (size_t)SMG_TRUNK_BN_CHANNELS * 2 * sizeof(double);
        const size_t o_s = p.take(W.sums_bytes);
        if (base) {
            W.dconv0 = reinterpret_cast<float*>(base + o_c0);
            W.t_a = reinterpret_cast<float*>(base + o_a);
            W.t_b = reinterpret_cast<float*>(base + o_b);
            W.t_c = reinterpret_cast<float*>(base + o_c);
            W.dP = reinterpret_cast<float*>(base + o_p);
            W.wg3_scratch = reinterpret_cast<float*>(base + o_w3);
            W.wg3_jobs_dev = base + o_wj;
            W.wg3_jobs.assign(58 * 2, nullptr);
            W.sums = reinterpret_cast<double*>(base + o_s);
        } else {
            W.bytes = p.off;
            SMG_CUDA(cudaMalloc(&W.arena, W.bytes));
            h->workspace_bytes += (int64_t)W.bytes;
            SMG_CUDA(cudaStreamCreateWithPriority(&W.wstream, cudaStreamNonBlocking, 0));   // lower priority than the main chain's
            for (auto& e : W.ev) SMG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        }
    }
    return SMG_OK;
}

// gradient pointer cursor over the smg_set_trunk_weights parameter order
struct TrunkGradMap {
    float* conv0;
    float* norm0[2];
    struct L { float* n1[2]; float* c1; float* n2[2]; float* c2; };
    std::vector<L> layers[kNumBlocks];
    struct T { float* n[2]; float* c; } trans[kNumBlocks - 1];
    float* norm5[2];
    void fill(float* const* g) {
        int i = 0;
        conv0 = g[i++]; norm0[0] = g[i++]; norm0[1] = g[i++];
        for (int b = 0; b < kNumBlocks; ++b) {
            layers[b].resize(kBlockLayers[b]);
            for (int l = 0; l < kBlockLayers[b]; ++l) {
                L& x = layers[b][l];
                x.n1[0] = g[i++]; x.n1[1] = g[i++]; x.c1 = g[i++];
                x.n2[0] = g[i++]; x.n2[1] = g[i++]; x.c2 = g[i++];
            }
            if (b < kNumBlocks - 1) { trans[b].n[0] = g[i++]; trans[b].n[1] = g[i++]; trans[b].c = g[i++]; }
        }
        norm5[0] = g[i++]; norm5[1] = g[i++];
    }
};

// a data-gradient convolution: identity prologue, flipped / transposed weights.  tf32 mode: the tcgen05 kernel of
// conv_umma.cu on the w_dgrad_tf32 stage image; fp32 mode: CUDA cores on w_dgrad.
static int dgrad_conv(smg_handle* h, ConvArgs a, const ConvW& cw, cudaStream_t st) {
    a.prologue_mode = 2;
    a.relu = 0;
    a.w = nullptr;
    const double px = (double)a.n * a.hin * a.hin;
    ProfScope ps(h, st, 3, 2.0 * px * a.cout * a.cin * a.taps, 4.0 * px * (a.cin + a.cout));
    if (h->precision == SMG_PREC_TF32 && cw.w_dgrad_tf32 != nullptr) {
        a.w_umma = cw.w_dgrad_tf32;
        return launch_conv_umma(h, a, SMG_PREC_TF32, st);
    }
    a.w_raw = cw.w_dgrad;
    return launch_conv_ffma(h, a, st);
}

// one BatchNorm(+ReLU) backward: reduce -> parameter grads -> apply
static int bn_backward(smg_handle* h, BnBwd a, int S, double*& sums_cursor, float* dgamma, float* dbeta, cudaStream_t st,
                       bool reduce_done = false) {
    a.sums = sums_cursor;
    sums_cursor += (size_t)2 * S * a.C;
    if (!reduce_done) SMG_TRY(launch_bn_bwd(h, a, S, false, nullptr, nullptr, st));   // else: fused into the producing dgrad
    return launch_bn_bwd(h, a, S, true, dgamma, dbeta, st);
}

// lets the epilogue of a tensor-core data-gradient convolution do the reduction of the BatchNorm backward that consumes its
// output (tf32 mode only: the CUDA-core kernel has no such epilogue); `sums` = the region bn_backward will be given next
static bool fuse_bn_reduce(const smg_handle* h, ConvArgs& a, const BnBwd& bb, double* sums) {
    static const bool on = getenv("SMG_BN_FUSE") == nullptr || atoi(getenv("SMG_BN_FUSE")) != 0;
    if (!on || h->precision != SMG_PREC_TF32 || bb.da_pooled || !bb.relu) return false;
    a.bnr_x = bb.x; a.bnr_x_cstride = bb.x_cstride; a.bnr_stats = bb.stats; a.bnr_stats_stride = bb.stats_stride;
    a.bnr_gamma = bb.gamma; a.bnr_beta = bb.beta; a.bnr_sums = sums;
    return true;
}

int qbackward_impl(smg_handle* h, const float* dq, float* const* tg, float* const* hg, cudaStream_t st) {
    smg_handle::TrainWs& W = h->train;
    TrunkW& T = h->trunks[W.trunk_id];
    HeadW& Hd = h->heads[W.head_id];
    const int S = 2;
    TrunkGradMap G;
    G.fill(tg);
    SMG_CUDA(cudaMemsetAsync(W.sums, 0, W.sums_bytes, st));
    double* sums = W.sums;
    // tf32 mode: the dense layers' weight gradients run on the tensor cores (wgrad_umma.cu; SMG_WGRAD_TC=0 keeps CUDA cores)
    static const bool wgrad_tc_env = getenv("SMG_WGRAD_TC") == nullptr || atoi(getenv("SMG_WGRAD_TC")) != 0;
    const bool wgrad_tc = wgrad_tc_env && h->precision == SMG_PREC_TF32;
    int n_wg3 = 0;
    // The weight gradients feed nothing downstream: with per-layer copies of d(conv1 output) they run on a second stream
    // next to the dgrad / BatchNorm chain (fork here, one dependency per layer, join before the stem).  Their grids are
    // capped so that the critical chain always finds free SMs.  SMG_WGRAD_ASYNC=0 keeps everything on one stream.
    static const bool async_env = getenv("SMG_WGRAD_ASYNC") == nullptr || atoi(getenv("SMG_WGRAD_ASYNC")) != 0;
    static const int cap_env = getenv("SMG_WGRAD_CTAS") ? atoi(getenv("SMG_WGRAD_CTAS")) : 64;
    const bool async_w = wgrad_tc && async_env;
    cudaStream_t ws = async_w ? W.wstream : st;
    h->wgrad_cta_cap = async_w ? cap_env : 0;
    if (async_w) {
        SMG_CUDA(cudaEventRecord(W.ev[58], st));
        SMG_CUDA(cudaStreamWaitEvent(ws, W.ev[58], 0));
    }
    if (wgrad_tc) SMG_CUDA(cudaMemsetAsync(W.wg3_scratch, 0, (size_t)58 * 288 * 128 * 4, ws));
    const BlockGeom& g4 = h->geom[3];
    const int npix4 = g4.hw * g4.hw;

    // ---- head: BN(64)+ReLU+20x20 conv, then the two halves of the 1x1 conv, then norm0 o norm5
    SMG_TRY(launch_head_tail_bwd(h, h->head_p, Hd, dq, W.dP, hg[3], hg[4], hg[5], st));
    SMG_CUDA(cudaMemsetAsync(hg[2], 0, (size_t)kHeadMid * 2 * kFeatC * 4, st));
    for (int half = 0; half < 2; ++half) {
        const int s = half;  // sample 0 = rotated scene -> channels [0,1024), sample 1 = mask -> [1024,2048)
        Wgrad wg{};
        wg.g = W.dP; wg.g_cstride = kHeadMid; wg.g_coff = 0; wg.cout = kHeadMid;
        wg.x = h->block[3] + (size_t)s * npix4 * g4.c_tot; wg.x_cstride = g4.c_tot; wg.cin = kFeatC; wg.hin = g4.hw; wg.hout = g4.hw;
        wg.prologue_mode = 1; wg.scale = h->head_scale + (size_t)s * kFeatC; wg.shift = h->head_shift + (size_t)s * kFeatC;
        wg.relu = 1; wg.dw = hg[2]; wg.k_total = 2 * kFeatC; wg.k_off = half * kFeatC;
        SMG_TRY(launch_wgrad(h, wg, 1, 1, 0, st));
        ConvArgs a;
        a.in = W.dP; a.in_cstride = kHeadMid; a.cin = kHeadMid; a.hin = g4.hw;
        a.taps = 1;
        a.out = W.t_c + (size_t)s * npix4 * kFeatC; a.out_cstride = kFeatC; a.out_coff = 0; a.cout = kFeatC;
        a.n = 1;
        SMG_TRY(dgrad_conv(h, a, Hd.conv0[half], st));
    }
    SMG_TRY(launch_head_norm_bwd(h, W.t_c, h->block[3], stats_ptr(h, h->st_block[3]), g4.c_tot, T.norm5, Hd.norm0,
                                 W.dblk[3], G.norm5[0], G.norm5[1], hg[0], hg[1], st));

    // ---- dense blocks in reverse
    int layer_index = 58;
    for (int b = kNumBlocks - 1; b >= 0; --b) {
        const BlockGeom& g = h->geom[b];
        const int npix = g.hw * g.hw;
        double* st_blk = stats_ptr(h, h->st_block[b]);
        for (int l = kBlockLayers[b] - 1; l >= 0; --l) {
            --layer_index;
            const DenseLayerW& L = T.layers[b][l];
            const TrunkGradMap::L& GL = G.layers[b][l];
            const int cin = g.c_in + l * kGrowth;
            double* st_bott = stats_ptr(h, h->st_bott + (size_t)layer_index * kBottleneck);
            float* y1 = W.bott_saved[layer_index];
            float* dy1 = async_w ? W.dy1_saved[layer_index] : W.t_b;   // d(conv1 output): kept per layer for the side stream
            // (1) 3x3 dgrad: d relu(bn2(y1)) = conv3x3(dX[:, cin:cin+32], flipped W2); its epilogue reduces BN2's backward
            BnBwd bb2{};
            bb2.da = W.t_a; bb2.da_cstride = kBottleneck; bb2.x = y1; bb2.x_cstride = kBottleneck;
            bb2.stats = st_bott; bb2.stats_stride = kBottleneck; bb2.gamma = L.norm2.gamma; bb2.beta = L.norm2.beta;
            bb2.C = kBottleneck; bb2.hw = g.hw; bb2.relu = 1; bb2.dst = dy1; bb2.dst_cstride = kBottleneck; bb2.accumulate = 0;
            bool fused2;
            {
                ConvArgs a;
                a.in = W.dblk[b] + cin; a.in_cstride = g.c_tot; a.cin = kGrowth; a.hin = g.hw;
                a.taps = 9;
                a.out = W.t_a; a.out_cstride = kBottleneck; a.out_coff = 0; a.cout = kBottleneck; a.n = S;
                fused2 = fuse_bn_reduce(h, a, bb2, sums);
                SMG_TRY(dgrad_conv(h, a, L.conv2, st));
            }
            // (3) BN2 + ReLU backward -> d y1
            SMG_TRY(bn_backward(h, bb2, S, sums, GL.n2[0], GL.n2[1], st, fused2));
            if (async_w) {   // both weight gradients of this layer may start: its gradient slice and d y1 are final
                SMG_CUDA(cudaEventRecord(W.ev[layer_index], st));
                SMG_CUDA(cudaStreamWaitEvent(ws, W.ev[layer_index], 0));
            }
            // (2) 3x3 wgrad
            int wg3_status = SMG_ERR_UNSUPPORTED;
            if (wgrad_tc) {
                float* scratch = W.wg3_scratch + (size_t)n_wg3 * 288 * 128;
                wg3_status = launch_wgrad3_umma(h, W.dblk[b], g.c_tot, cin, y1, g.hw, S, st_bott, kBottleneck, L.norm2.gamma,
                                                L.norm2.beta, scratch, ws);
                if (wg3_status == SMG_OK) {
                    W.wg3_jobs[2 * n_wg3] = scratch;
                    W.wg3_jobs[2 * n_wg3 + 1] = GL.c2;
                    ++n_wg3;
                } else if (wg3_status != SMG_ERR_UNSUPPORTED) {
                    return wg3_status;
                }
            }
            if (wg3_status == SMG_ERR_UNSUPPORTED) {
                SMG_CUDA(cudaMemsetAsync(GL.c2, 0, (size_t)kGrowth * kBottleneck * 9 * 4, ws));
                Wgrad wg{};
                wg.g = W.dblk[b]; wg.g_cstride = g.c_tot; wg.g_coff = cin; wg.cout = kGrowth;
                wg.x = y1; wg.x_cstride = kBottleneck; wg.cin = kBottleneck; wg.hin = g.hw; wg.hout = g.hw;
                wg.prologue_mode = 0; wg.stats = st_bott; wg.stats_stride = kBottleneck; wg.gamma = L.norm2.gamma; wg.beta = L.norm2.beta;
                wg.relu = 1; wg.dw = GL.c2; wg.k_total = kBottleneck; wg.k_off = 0;
                SMG_TRY(launch_wgrad(h, wg, S, 9, 0, ws));
            }
            // (4) 1x1 wgrad
            SMG_CUDA(cudaMemsetAsync(GL.c1, 0, (size_t)kBottleneck * cin * 4, ws));
            int wg1_status = SMG_ERR_UNSUPPORTED;
            if (wgrad_tc) {
                wg1_status = launch_wgrad1_umma(h, dy1, h->block[b], g.c_tot, cin, g.hw, S, st_blk, g.c_tot, L.norm1.gamma,
                                                L.norm1.beta, GL.c1, ws);
                if (wg1_status != SMG_OK && wg1_status != SMG_ERR_UNSUPPORTED) return wg1_status;
            }
            if (wg1_status == SMG_ERR_UNSUPPORTED) {
                Wgrad wg{};
                wg.g = dy1; wg.g_cstride = kBottleneck; wg.g_coff = 0; wg.cout = kBottleneck;
                wg.x = h->block[b]; wg.x_cstride = g.c_tot; wg.cin = cin; wg.hin = g.hw; wg.hout = g.hw;
                wg.prologue_mode = 0; wg.stats = st_blk; wg.stats_stride = g.c_tot; wg.gamma = L.norm1.gamma; wg.beta = L.norm1.beta;
                wg.relu = 1; wg.dw = GL.c1; wg.k_total = cin; wg.k_off = 0;
                SMG_TRY(launch_wgrad(h, wg, S, 1, 0, ws));
            }
            // (5) 1x1 dgrad: d relu(bn1(X[:, :cin])) = dy1 . W1
            // (6) BN1 + ReLU backward, accumulated into the block gradient; its reduction rides in the 1x1 dgrad's epilogue
            {
                BnBwd bb{};
                bb.da = W.t_c; bb.da_cstride = cin; bb.x = h->block[b]; bb.x_cstride = g.c_tot;
                bb.stats = st_blk; bb.stats_stride = g.c_tot; bb.gamma = L.norm1.gamma; bb.beta = L.norm1.beta;
                bb.C = cin; bb.hw = g.hw; bb.relu = 1; bb.dst = W.dblk[b]; bb.dst_cstride = g.c_tot; bb.accumulate = 1;
                ConvArgs a;
                a.in = dy1; a.in_cstride = kBottleneck; a.cin = kBottleneck; a.hin = g.hw;
                a.taps = 1;
                a.out = W.t_c; a.out_cstride = cin; a.out_coff = 0; a.cout = cin; a.n = S;
                const bool fused1 = fuse_bn_reduce(h, a, bb, sums);
                SMG_TRY(dgrad_conv(h, a, L.conv1, st));
                SMG_TRY(bn_backward(h, bb, S, sums, GL.n1[0], GL.n1[1], st, fused1));
            }
        }
        if (b > 0) {
            // transition b-1: X_b[:, :C/2] = conv1x1(avgpool(relu(bn(X_{b-1}))))
            const BlockGeom& gp = h->geom[b - 1];
            const TransitionW& R = T.trans[b - 1];
            const int C = gp.c_tot, Co = C / 2;
            double* st_prev = stats_ptr(h, h->st_block[b - 1]);
            SMG_CUDA(cudaMemsetAsync(G.trans[b - 1].c, 0, (size_t)Co * C * 4, st));
            {
                Wgrad wg{};
                wg.g = W.dblk[b]; wg.g_cstride = g.c_tot; wg.g_coff = 0; wg.cout = Co;
                wg.x = h->block[b - 1]; wg.x_cstride = C; wg.cin = C; wg.hin = gp.hw; wg.hout = g.hw;
                wg.prologue_mode = 0; wg.stats = st_prev; wg.stats_stride = C; wg.gamma = R.norm.gamma; wg.beta = R.norm.beta;
                wg.relu = 1; wg.dw = G.trans[b - 1].c; wg.k_total = C; wg.k_off = 0;
                SMG_TRY(launch_wgrad(h, wg, S, 1, 1, st));
            }
            {
                ConvArgs a;
                a.in = W.dblk[b]; a.in_cstride = g.c_tot; a.cin = Co; a.hin = g.hw;
                a.taps = 1;
                a.out = W.t_c; a.out_cstride = C; a.out_coff = 0; a.cout = C; a.n = S;
                SMG_TRY(dgrad_conv(h, a, R.conv, st));
            }
            {
                BnBwd bb{};
                bb.da = W.t_c; bb.da_cstride = C; bb.da_pooled = 1; bb.x = h->block[b - 1]; bb.x_cstride = C;
                bb.stats = st_prev; bb.stats_stride = C; bb.gamma = R.norm.gamma; bb.beta = R.norm.beta;
                bb.C = C; bb.hw = gp.hw; bb.relu = 1; bb.dst = W.dblk[b - 1]; bb.dst_cstride = C; bb.accumulate = 0;
                SMG_TRY(bn_backward(h, bb, S, sums, G.trans[b - 1].n[0], G.trans[b - 1].n[1], st));
            }
        }
    }
    if (n_wg3 > 0) {
        // the table lives in the handle (stable host address): a captured copy node re-reads it at every replay
        SMG_CUDA(cudaMemcpyAsync(W.wg3_jobs_dev, W.wg3_jobs.data(), (size_t)n_wg3 * 2 * sizeof(void*), cudaMemcpyHostToDevice, ws));
        SMG_TRY(launch_wgrad3_finish(h, W.wg3_jobs_dev, n_wg3, ws));
    }
    if (async_w) {   // join: everything after this point (and the caller) sees the weight gradients
        SMG_CUDA(cudaEventRecord(W.ev[59], ws));
        SMG_CUDA(cudaStreamWaitEvent(st, W.ev[59], 0));
    }
    h->wgrad_cta_cap = 0;
    // ---- stem: maxpool -> relu/bn0 -> conv0 wgrad (no data gradient is needed for the input image)
    const int Hc = h->H / 2;
    SMG_CUDA(cudaMemsetAsync(W.dconv0, 0, (size_t)S * Hc * Hc * 64 * 4, st));
    SMG_TRY(launch_pool0_bwd(h, S, W.dblk[0], h->geom[0].c_tot, h->conv0, stats_ptr(h, h->st_conv0), T.norm0.gamma,
                             T.norm0.beta, W.dconv0, st));
    {
        BnBwd bb{};
        bb.da = W.dconv0; bb.da_cstride = 64; bb.x = h->conv0; bb.x_cstride = 64;
        bb.stats = stats_ptr(h, h->st_conv0); bb.stats_stride = 64; bb.gamma = T.norm0.gamma; bb.beta = T.norm0.beta;
        bb.C = 64; bb.hw = Hc; bb.relu = 1; bb.dst = W.dconv0; bb.dst_cstride = 64; bb.accumulate = 0;
        SMG_TRY(bn_backward(h, bb, S, sums, G.norm0[0], G.norm0[1], st));
    }
    if (W.in_channels == 1) {
        // Trainer.forward feeds three identical channels: one 49-tap gradient, replicated into [64][3][7][7]
        float* g1 = W.t_a;
        SMG_CUDA(cudaMemsetAsync(g1, 0, (size_t)64 * 49 * 4, st));
        SMG_TRY(launch_conv0_wgrad(h, S, W.dconv0, h->input, 1, g1, st));
        SMG_TRY(launch_replicate_conv0_grad(h, g1, G.conv0, st));
    } else {
        SMG_CUDA(cudaMemsetAsync(G.conv0, 0, (size_t)64 * 147 * 4, st));
        SMG_TRY(launch_conv0_wgrad(h, S, W.dconv0, h->input, 3, G.conv0, st));
    }
    return SMG_OK;
}

}  // namespace smg

extern "C" {

int smg_qforward_train(smg_handle* h, int trunk_id, int head_id, const float* dev_scene, const float* dev_mask,
                       int rot_idx, int num_rotations, float* dev_q, float* dev_bn_mean, float* dev_bn_var, void* stream) {
    SMG_CHECK(h && dev_scene && dev_mask && dev_q, SMG_ERR_INVALID, "smg_qforward_train: NULL argument");
    SMG_CHECK(trunk_id >= 0 && trunk_id < SMG_NUM_TRUNKS && head_id >= 0 && head_id < SMG_NUM_HEADS, SMG_ERR_INVALID,
              "smg_qforward_train: trunk %d / head %d", trunk_id, head_id);
    SMG_CHECK(h->max_samples >= 2, SMG_ERR_STATE, "smg_qforward_train needs max_samples >= 2");
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    SMG_TRY(ensure_train_workspace(h));
    h->train.valid = false;
    const size_t img = (size_t)3 * h->H * h->H;
    SMG_TRY(launch_rotate(h, dev_scene, &rot_idx, 1, num_rotations, h->input, 3, st));
    SMG_CUDA(cudaMemcpyAsync(h->input + img, dev_mask, img * 4, cudaMemcpyDeviceToDevice, st));
    SMG_TRY(trunk_forward(h, trunk_id, 2, 3, st, true));
    SMG_TRY(heads_forward(h, trunk_id, head_id, 1, 1, dev_q, st));
    if (dev_bn_mean && dev_bn_var) SMG_TRY(export_bn_stats(h, 2, dev_bn_mean, dev_bn_var, st));
    h->train.trunk_id = trunk_id;
    h->train.head_id = head_id;
    h->train.in_channels = 3;
    h->train.valid = true;
    h->train.pass_id++;
    return SMG_OK;
}

int smg_qbackward(smg_handle* h, const float* dev_dq, float* const* dev_trunk_grads, float* const* dev_head_grads,
                  void* stream) {
    SMG_CHECK(h && dev_dq && dev_trunk_grads && dev_head_grads, SMG_ERR_INVALID, "smg_qbackward: NULL argument");
    SMG_CHECK(h->train.valid, SMG_ERR_STATE, "smg_qbackward: no smg_qforward_train result is pending on this handle");
    SMG_CHECK(h->trunks[h->train.trunk_id].packed & SMG_PACK_DGRAD, SMG_ERR_STATE,
              "smg_qbackward: the trunk weights were packed without the data-gradient layout (smg_set_pack_layouts)");
    DeviceGuard guard(h->device);
    const int status = qbackward_impl(h, dev_dq, dev_trunk_grads, dev_head_grads, (cudaStream_t)stream);
    h->train.valid = false;
    return status;
}

int64_t smg_train_pass_id(smg_handle* h) { return (h && h->train.valid) ? h->train.pass_id : -1; }

int smg_adam_step(smg_handle* h, float* const* dev_params, const float* const* dev_grads, float* const* dev_m,
                  float* const* dev_v, const int64_t* host_numel, int n_tensors, int step, float lr, float beta1,
                  float beta2, float eps, void* stream) {
    SMG_CHECK(h && dev_params && dev_grads && dev_m && dev_v && host_numel && n_tensors >= 0 && step >= 1, SMG_ERR_INVALID,
              "smg_adam_step: bad argument");
    DeviceGuard guard(h->device);
    return adam_multi_tensor(h, dev_params, dev_grads, dev_m, dev_v, host_numel, n_tensors, step, lr, beta1, beta2, eps,
                             (cudaStream_t)stream);
}

int smg_argmax(smg_handle* h, const float* dev_q, int n, float* dev_out, int32_t* dev_out_idx, void* stream) {
    SMG_CHECK(h && dev_q && dev_out && dev_out_idx, SMG_ERR_INVALID, "smg_argmax: NULL argument");
    DeviceGuard guard(h->device);
    ProfScope ps(h, (cudaStream_t)stream, 3, 0.0, 4.0 * n);
    return launch_argmax(h, dev_q, n, dev_out, dev_out_idx, (cudaStream_t)stream);
}

int smg_heightmap_color(smg_handle* h, const uint8_t* dev_color, uint8_t* dev_out224, uint8_t* dev_out448, void* stream) {
    SMG_CHECK(h && dev_color && dev_out224 && dev_out448, SMG_ERR_INVALID, "smg_heightmap_color: NULL argument");
    DeviceGuard guard(h->device);
    return launch_heightmap_color(h, dev_color, dev_out224, dev_out448, (cudaStream_t)stream);
}

int smg_heightmap(smg_handle* h, const double* dev_depth, const double* host_K, const double* host_pose,
                  double* dev_out224, double* dev_out448, double* host_A_htor, void* stream) {
    SMG_CHECK(h && dev_depth && host_K && host_pose && dev_out224 && dev_out448, SMG_ERR_INVALID, "smg_heightmap: NULL argument");
    DeviceGuard guard(h->device);
    return launch_heightmap(h, dev_depth, host_K, host_pose, dev_out224, dev_out448, host_A_htor, (cudaStream_t)stream);
}

int smg_resize_masks(smg_handle* h, const float* dev_masks, int n, int size_in, int size_out, float* dev_out, void* stream) {
    SMG_CHECK(h && (n == 0 || (dev_masks && dev_out)) && n >= 0 && size_in >= 2 && size_out >= 1, SMG_ERR_INVALID,
              "smg_resize_masks: bad argument");
    DeviceGuard guard(h->device);
    return launch_resize_masks(h, dev_masks, n, size_in, size_out, dev_out, (cudaStream_t)stream);
}

int smg_geometry(smg_handle* h, int mode, const double* dev_depth, int img_h, int img_w, const double* host_A_htor,
                 const double* host_K, const double* host_pose, const double* host_boxes, const double* host_centers,
                 int n_objects, int best_id, int flag, const double* host_pix, double* host_out, void* stream) {
    SMG_CHECK(h && dev_depth && host_A_htor && host_K && host_pose && host_out, SMG_ERR_INVALID, "smg_geometry: NULL argument");
    SMG_CHECK(mode >= 0 && mode <= 2 && (mode == 0 ? host_pix != nullptr : host_boxes != nullptr) && (mode != 2 || !flag || host_centers),
              SMG_ERR_INVALID, "smg_geometry: mode %d with missing inputs", mode);
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (!h->geo_out) SMG_CUDA(cudaMalloc(&h->geo_out, 8 * sizeof(double)));
    SMG_TRY(launch_geometry(h, mode, dev_depth, img_h, img_w, host_A_htor, host_K, host_pose, host_boxes, host_centers, n_objects,
                            best_id, flag, host_pix, h->geo_out, st));
    double out[6];
    SMG_CUDA(cudaMemcpyAsync(out, h->geo_out, sizeof(out), cudaMemcpyDeviceToHost, st));
    SMG_CUDA(cudaStreamSynchronize(st));
    for (int i = 0; i < 5; ++i) host_out[i] = out[i];
    const int status = (int)out[5];
    SMG_CHECK(!(status & 1), SMG_ERR_INVALID, "smg_geometry: a heightmap pixel maps outside the %d x %d camera image", img_w, img_h);
    SMG_CHECK(!(status & 2), SMG_ERR_STATE, "smg_geometry: no free suction direction was found");
    return SMG_OK;
}

int smg_nms(smg_handle* h, const float* dev_boxes, int n, float co_thresh, float min_area, float max_area,
            int32_t* dev_keep, int32_t* dev_n_keep, void* stream) {
    SMG_CHECK(h && dev_keep && dev_n_keep && (dev_boxes || n == 0), SMG_ERR_INVALID, "smg_nms: NULL argument");
    DeviceGuard guard(h->device);
    return launch_nms(h, dev_boxes, n, co_thresh, min_area, max_area, dev_keep, dev_n_keep, (cudaStream_t)stream);
}

int smg_debug_read(smg_handle* h, const char* what, int sample, float* dev_out_nchw, int64_t capacity_floats,
                   void* stream) {
    SMG_CHECK(h && what && dev_out_nchw, SMG_ERR_INVALID, "smg_debug_read: NULL argument");
    SMG_CHECK(sample >= 0 && sample < h->last_n, SMG_ERR_INVALID, "smg_debug_read: sample %d of %d", sample, h->last_n);
    DeviceGuard guard(h->device);
    const float* src = nullptr;
    int hw = 0, c = 0, cstride = 0;
    if (!strcmp(what, "conv0")) {
        hw = h->H / 2; c = 64; cstride = 64; src = h->conv0;
    } else if (!strcmp(what, "pool0")) {
        hw = h->geom[0].hw; c = 64; cstride = h->geom[0].c_tot; src = h->block[0];
    } else if (!strncmp(what, "block", 5) && what[5] >= '1' && what[5] <= '4') {
        const int b = what[5] - '1';
        hw = h->geom[b].hw; c = h->geom[b].c_tot; cstride = c; src = h->block[b];
    } else if (!strncmp(what, "trans", 5) && what[5] >= '1' && what[5] <= '3') {
        const int b = what[5] - '1' + 1;
        hw = h->geom[b].hw; c = h->geom[b].c_in; cstride = h->geom[b].c_tot; src = h->block[b];
    } else if (!strcmp(what, "bott")) {
        hw = h->geom[3].hw; c = kBottleneck; cstride = kBottleneck; src = h->bott;
    } else {
        set_error("smg_debug_read: unknown activation '%s'", what);
        return SMG_ERR_INVALID;
    }
    SMG_CHECK((int64_t)hw * hw * c <= capacity_floats, SMG_ERR_INVALID, "smg_debug_read: need %lld floats",
              (long long)hw * hw * c);
    src += (size_t)sample * hw * hw * cstride;
    return launch_nhwc_to_nchw(h, src, hw, c, cstride, dev_out_nchw, (cudaStream_t)stream);
}

int smg_head_bn_stats(smg_handle* h, float* dev_out, int n_pairs, void* stream) {
    SMG_CHECK(h && dev_out, SMG_ERR_INVALID, "smg_head_bn_stats: NULL argument");
    SMG_CHECK(n_pairs == h->head_bn1_pairs && n_pairs > 0, SMG_ERR_STATE, "smg_head_bn_stats: the last head pass had %d pairs, not %d",
              h->head_bn1_pairs, n_pairs);
    DeviceGuard guard(h->device);
    SMG_CUDA(cudaMemcpyAsync(dev_out, h->head_bn1, (size_t)n_pairs * 128 * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return SMG_OK;
}

int smg_set_pack_layouts(smg_handle* h, int mask) {
    SMG_CHECK(h != nullptr && mask > 0 && mask <= SMG_PACK_ALL, SMG_ERR_INVALID, "smg_set_pack_layouts: mask %d", mask);
    h->pack_mask = mask;
    return SMG_OK;
}

int smg_profile_enable(smg_handle* h, int enable) {
    SMG_CHECK(h != nullptr, SMG_ERR_INVALID, "NULL handle");
    DeviceGuard guard(h->device);
    for (auto& r : h->prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    h->prof.clear();
    h->profile = enable != 0;
    return SMG_OK;
}

int smg_profile_read(smg_handle* h, double* host_ms, int64_t* host_launches, double* host_flops, double* host_bytes) {
    SMG_CHECK(h && host_ms && host_launches && host_flops && host_bytes, SMG_ERR_INVALID, "smg_profile_read: NULL argument");
    DeviceGuard guard(h->device);
    SMG_CUDA(cudaDeviceSynchronize());
    for (int i = 0; i < SMG_PROFILE_CLASSES; ++i) { host_ms[i] = 0; host_launches[i] = 0; host_flops[i] = 0; host_bytes[i] = 0; }
    for (auto& r : h->prof) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.e0, r.e1);
        host_ms[r.cls] += ms;
        host_launches[r.cls] += 1;
        host_flops[r.cls] += r.flops;
        host_bytes[r.cls] += r.bytes;
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    h->prof.clear();
    return SMG_OK;
}

int smg_debug_conv(smg_handle* h, int precision, const float* dev_in, int n, int hin, int cin, int in_cstride,
                   const float* dev_scale, const float* dev_shift, int relu, int pool, int taps,
                   const float* dev_w_oihw, int cout, float* dev_out, int out_cstride, int out_coff,
                   double* dev_out_stats, void* stream) {
    SMG_CHECK(h && dev_in && dev_scale && dev_shift && dev_w_oihw && dev_out, SMG_ERR_INVALID, "smg_debug_conv: NULL argument");
    SMG_CHECK(taps == 1 || taps == 9, SMG_ERR_INVALID, "smg_debug_conv: taps %d", taps);
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    ConvW cw;
    ArenaPlanner p;
    plan_conv(p, cw, cin, cout, taps, nullptr);
    uint8_t* base = nullptr;
    SMG_CUDA(cudaMalloc(&base, p.off));
    ArenaPlanner p2;
    plan_conv(p2, cw, cin, cout, taps, base);
    int status = pack_conv_weights(h, dev_w_oihw, cw, 0, cin, st);
    if (status == SMG_OK) {
        ConvArgs a;
        a.in = dev_in; a.in_cstride = in_cstride; a.cin = cin; a.hin = hin;
        a.prologue_mode = 1; a.scale = dev_scale; a.shift = dev_shift; a.relu = relu;
        a.pool = pool; a.taps = taps; a.w = &cw;
        a.out = dev_out; a.out_cstride = out_cstride; a.out_coff = out_coff; a.cout = cout;
        a.out_stats = dev_out_stats; a.out_stats_stride = out_cstride;
        a.n = n;
        const int saved = h->precision;
        h->precision = precision;
        status = conv_dispatch(h, a, st);  // same kernel selection (one-tile / multi-tile) as the trunk schedule
        h->precision = saved;
    }
    cudaError_t e = cudaStreamSynchronize(st);
    cudaFree(base);
    if (status == SMG_OK && e != cudaSuccess) {
        set_error("smg_debug_conv: %s", cudaGetErrorString(e));
        return SMG_ERR_CUDA;
    }
    return status;
}

int smg_debug_dgrad(smg_handle* h, int precision, const float* dev_g, int n, int hin, int cout, int g_cstride, int g_coff,
                    int taps, const float* dev_w_oihw, int cin, float* dev_dx, void* stream) {
    SMG_CHECK(h && dev_g && dev_w_oihw && dev_dx, SMG_ERR_INVALID, "smg_debug_dgrad: NULL argument");
    SMG_CHECK(taps == 1 || taps == 9, SMG_ERR_INVALID, "smg_debug_dgrad: taps %d", taps);
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    ConvW cw;
    ArenaPlanner p;
    plan_conv(p, cw, cin, cout, taps, nullptr);
    uint8_t* base = nullptr;
    SMG_CUDA(cudaMalloc(&base, p.off));
    ArenaPlanner p2;
    plan_conv(p2, cw, cin, cout, taps, base);
    int status = pack_conv_weights(h, dev_w_oihw, cw, 0, cin, st);
    if (status == SMG_OK) {
        ConvArgs a;
        a.in = dev_g + g_coff; a.in_cstride = g_cstride; a.cin = cout; a.hin = hin;
        a.taps = taps;
        a.out = dev_dx; a.out_cstride = cin; a.out_coff = 0; a.cout = cin;
        a.n = n;
        const int saved = h->precision;
        h->precision = precision;
        status = dgrad_conv(h, a, cw, st);
        h->precision = saved;
    }
    cudaError_t e = cudaStreamSynchronize(st);
    cudaFree(base);
    if (status == SMG_OK && e != cudaSuccess) {
        set_error("smg_debug_dgrad: %s", cudaGetErrorString(e));
        return SMG_ERR_CUDA;
    }
    return status;
}

int smg_debug_wgrad(smg_handle* h, int taps, const float* dev_g, int g_cstride, int g_coff, const float* dev_x, int x_cstride,
                    int cin, int hw, int S, const double* dev_stats, int stats_stride, const float* dev_gamma,
                    const float* dev_beta, float* dev_dw, void* stream) {
    SMG_CHECK(h && dev_g && dev_x && dev_stats && dev_gamma && dev_beta && dev_dw, SMG_ERR_INVALID, "smg_debug_wgrad: NULL argument");
    SMG_CHECK(taps == 1 || taps == 9, SMG_ERR_INVALID, "smg_debug_wgrad: taps %d", taps);
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    int status;
    if (taps == 1) {
        SMG_CHECK(g_cstride == 128 && g_coff == 0, SMG_ERR_INVALID, "smg_debug_wgrad: the 1x1 gradient operand is dense [..,128]");
        SMG_CUDA(cudaMemsetAsync(dev_dw, 0, (size_t)128 * cin * 4, st));
        status = launch_wgrad1_umma(h, dev_g, dev_x, x_cstride, cin, hw, S, dev_stats, stats_stride, dev_gamma, dev_beta, dev_dw, st);
    } else {
        SMG_CHECK(x_cstride == 128 && cin == 128, SMG_ERR_INVALID, "smg_debug_wgrad: the 3x3 activation operand is dense [..,128]");
        uint8_t* base = nullptr;
        const size_t sb = (size_t)288 * 128 * 4;
        SMG_CUDA(cudaMalloc(&base, sb + 64));
        SMG_CUDA(cudaMemsetAsync(base, 0, sb, st));
        float* scratch = reinterpret_cast<float*>(base);
        status = launch_wgrad3_umma(h, dev_g, g_cstride, g_coff, dev_x, hw, S, dev_stats, stats_stride, dev_gamma, dev_beta,
                                    scratch, st);
        if (status == SMG_OK) {
            void* job[2] = {scratch, dev_dw};
            SMG_CUDA(cudaMemcpyAsync(base + sb, job, sizeof(job), cudaMemcpyHostToDevice, st));
            status = launch_wgrad3_finish(h, base + sb, 1, st);
        }
        cudaStreamSynchronize(st);
        cudaFree(base);
    }
    if (status == SMG_ERR_UNSUPPORTED) set_error("smg_debug_wgrad: shape not served by the tensor-core kernels");
    if (status != SMG_OK) return status;
    SMG_CUDA(cudaStreamSynchronize(st));
    return SMG_OK;
}

int smg_debug_bn_bwd(smg_handle* h, const float* dev_da, int da_cstride, int da_pooled, const float* dev_x,
                     int x_cstride, const double* dev_stats, int stats_stride, const float* dev_gamma,
                     const float* dev_beta, int C, int hw, int relu, int S, double* dev_sums, float* dev_dst,
                     int dst_cstride, int accumulate, float* dev_dgamma, float* dev_dbeta, void* stream) {
    SMG_CHECK(h && dev_da && dev_x && dev_stats && dev_gamma && dev_beta && dev_sums && dev_dst && dev_dgamma && dev_dbeta,
              SMG_ERR_INVALID, "smg_debug_bn_bwd: NULL argument");
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    BnBwd bb{};
    bb.da = dev_da; bb.da_cstride = da_cstride; bb.da_pooled = da_pooled; bb.x = dev_x; bb.x_cstride = x_cstride;
    bb.stats = dev_stats; bb.stats_stride = stats_stride; bb.gamma = dev_gamma; bb.beta = dev_beta;
    bb.C = C; bb.hw = hw; bb.relu = relu; bb.sums = dev_sums; bb.dst = dev_dst; bb.dst_cstride = dst_cstride;
    bb.accumulate = accumulate;
    SMG_TRY(launch_bn_bwd(h, bb, S, false, nullptr, nullptr, st));
    SMG_TRY(launch_bn_bwd(h, bb, S, true, dev_dgamma, dev_dbeta, st));
    SMG_CUDA(cudaStreamSynchronize(st));
    return SMG_OK;
}

}  // extern "C"
