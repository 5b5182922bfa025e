// smg_internal.cuh - shared declarations of libsmg_b200.so (not part of the public ABI).
//
// Data layout in HBM (per handle, sized for max_samples S and input size H):
//   input      [S,3,H,H]            fp32 planar (what the reference feeds the net)
//   conv0 raw  [S,H/2,H/2,64]       fp32 NHWC, pre-BN
//   block b    [S,Hb,Hb,Ctot_b]     fp32 NHWC, pre-BN ("dense block buffer": every dense layer
//                                   writes its 32 new channels into its slice, so torch.cat
//                                   (densenet `bn_function`) never copies)
//   bottleneck [S,Hb,Hb,128]        fp32 NHWC scratch (conv1 output of the layer in flight)
//   stats      double2 per (sample, channel): (sum x, sum x^2) over H*W, accumulated by the
//              kernel that PRODUCES the channel; every consumer BN derives mean/var from it.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/smg_b200.h"

namespace smg {

constexpr int kNumBlocks = 4;
constexpr int kBlockLayers[kNumBlocks] = {6, 12, 24, 16};
constexpr int kGrowth = 32;
constexpr int kBottleneck = 128;
constexpr int kInitFeatures = 64;
constexpr float kBnEps = 1e-5f;
constexpr int kHeadMid = 64;
constexpr int kHeadK = 20;  // 20x20 valid conv (code/models.py:322)
constexpr int kFeatC = 1024;

// thread-local error string
void set_error(const char* fmt, ...);
const char* get_error();

#define SMG_CUDA(call)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (call);                                                               \
        if (_e != cudaSuccess) {                                                               \
            smg::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
            return SMG_ERR_CUDA;                                                               \
        }                                                                                      \
    } while (0)

#define SMG_CHECK(cond, code, ...)          \
    do {                                    \
        if (!(cond)) {                      \
            smg::set_error(__VA_ARGS__);    \
            return (code);                  \
        }                                   \
    } while (0)

#define SMG_TRY(expr)            \
    do {                         \
        int _s = (expr);         \
        if (_s != SMG_OK) return _s; \
    } while (0)

// ---------------------------------------------------------------------------------------
// one convolution of the trunk as the implicit-GEMM kernels see it
// ---------------------------------------------------------------------------------------
struct ConvW {
    // fp32 CUDA-core layout: [taps][Cin][Cout]
    float* w_ffma = nullptr;
    // tcgen05 layouts: a sequence of ready-to-copy shared-memory stage images
    // (no-swizzle K-major core matrices, see conv_umma.cu), tf32 (4-byte) and bf16
    uint8_t* w_tf32 = nullptr;
    // tf32, weights as the A operand in tensor memory: 1x1 with cout = 128 (conv1_t.cu): [cin/16][128][16] = w[row][16 c16 + c];
    // 3x3 (conv3_wt.cu): [column block of 16][dx*32 + cout][16], column = ((g*3 + dy)*4 + k)*8 + e
    uint8_t* w_tf32_t = nullptr;   // same with 32-channel groups: [kgroup][dy][chunk (8)][dx*32 + cout][4]
    uint8_t* w_bf16 = nullptr;
    // data-gradient weights for the fp32 kernels, in the same [taps][K][N] layout with the roles swapped:
    // 1x1: [cout][cin] (the torch layout); 3x3: [flipped tap][cout][cin]
    float* w_dgrad = nullptr;
    // the same data-gradient weights as a tf32 stage image for conv_umma.cu (N = 128 tiles, zero-padded to a whole tile):
    // 1x1: [ntile over cin/128][kgroup over cout/32][chunk][n][4], value w[kgroup*32 + ..][ntile*128 + n];
    // 3x3: [kgroup over cout/32][flipped tap][chunk][n = cin][4]
    uint8_t* w_dgrad_tf32 = nullptr;
    // fp32 mode on the tensor cores ("3xTF32"): 1x1: the w_tf32 stage sequence with every stage stored as [hi image][lo image]
    // (hi = nearest tf32 value, lo = weight - hi); 3x3: stages (kgroup, dy) of [hi | lo] x [chunk][dx*cout + co][4] for the
    // dx-merged form of conv_umma.cu (TAPS == 3)
    uint8_t* w_split = nullptr;
    int cin = 0, cout = 0, taps = 1;
};

// weight-packing job tables (pack.cu)
struct PackJob {
    const float* src;
    float* ffma;
    float* tf32;
    float* tf32_t;
    float* dgrad_tf32;
    float* split;
    __nv_bfloat16* bf16;
    float* dgrad;
    int cin, cout, taps, k_off, k_total, bn;
};
struct CopyJob {
    const float* src;
    float* dst;
    int n;
};
struct BnP {
    float* gamma = nullptr;
    float* beta = nullptr;
    int c = 0;
};

struct DenseLayerW {
    BnP norm1;
    ConvW conv1;  // 1x1 Cin -> 128
    BnP norm2;
    ConvW conv2;  // 3x3 128 -> 32
};

struct TransitionW {
    BnP norm;
    ConvW conv;  // 1x1 C -> C/2 (applied after the 2x2 average pool, which commutes with it)
};

struct TrunkW {
    bool set = false;
    int packed = 0;  // SMG_PACK_* layouts that are current
    float* conv0 = nullptr;  // [147][64], k = (c*7+kh)*7+kw
    float* conv0_folded = nullptr;  // [49][64]: weights summed over the input channel (identical channels)
    float* conv0_umma = nullptr;    // the folded weights as a tensor-memory A image [4 column blocks][128 rows = w_hi | w_lo][16], k = kh*8 + kw (stem_umma.cu)
    BnP norm0;
    std::vector<DenseLayerW> layers[kNumBlocks];
    TransitionW trans[kNumBlocks - 1];
    BnP norm5;
    void* arena = nullptr;
    size_t arena_bytes = 0;
    // packing tables of the last smg_set_trunk_weights (sources = the caller's parameter tensors)
    std::vector<const float*> src;
    std::vector<PackJob> pack_jobs;
    std::vector<CopyJob> copy_jobs;
    void* jobs_dev = nullptr;
};

struct HeadW {
    bool set = false;
    int packed = 0;
    int n_out = 0;
    BnP norm0;        // 2048
    ConvW conv0[2];   // 1x1 2048 -> 64 split into the scene half [0] and the mask half [1] (K = 1024 each)
    BnP norm1;        // 64
    float* conv1 = nullptr;  // [n_out][400][64]  (pixel-major, channel fastest)
    void* arena = nullptr;
    size_t arena_bytes = 0;
    std::vector<const float*> src;   // parameter tensors of the last smg_set_head_weights
};

// geometry of one dense block
struct BlockGeom {
    int hw;     // spatial size (square)
    int c_in;   // channels entering the block
    int c_tot;  // channels at the end of the block
};

}  // namespace smg

// the public opaque handle
struct smg_handle {
    int device = 0;
    int max_samples = 0;
    int H = 0;
    int precision = SMG_PREC_FP32;
    int num_sms = 148;
    int64_t launches = 0;
    int64_t workspace_bytes = 0;
    double l2_chunk_bytes = 0.0;   // >0: run each dense block over sample chunks of about this footprint (L2 residency)
    int pack_mask = 15;            // weight layouts written by smg_set_*_weights (SMG_PACK_*)
    int use_tma = 240;             // A/B bit mask of the persistent TMA-fed tf32 kernels (SMG_TMA): 16 = tensor-core 7x7 stem for
                                   // identical input channels (stem_umma.cu), 64 = 3x3 with the weights resident in tensor memory
                                   // (conv3_wt.cu), 128 = persistent 1x1 with swapped operand roles (conv1_t.cu); a cleared bit
                                   // routes the layer to the register-producer kernel of conv_umma.cu
    bool use_pdl = true;           // programmatic dependent launch between the persistent dense-layer kernels (SMG_PDL=0: off)
    bool fp32_exact_pass = false;  // set while a grad-enabled fp32 pass runs: CUDA-core FFMA convolutions (see trunk_forward)
    bool fp32_tc = true;           // fp32 mode: convolutions on the tensor cores with hi/lo split operands (SMG_FP32_TC=0: CUDA cores)
    int wgrad_cta_cap = 0;         // > 0 while the weight gradients share the GPU with the dgrad chain: max CTAs per wgrad launch
    int force_async = 0;           // tuning: -1 auto by grid size, 0 register producers (default: measured fastest), 1 cp.async producers
    std::vector<const void*> smem_opt_in;   // kernels whose >48 KB dynamic shared memory opt-in was set on this handle's device

    smg::BlockGeom geom[smg::kNumBlocks];
    smg::TrunkW trunks[SMG_NUM_TRUNKS];
    smg::HeadW heads[SMG_NUM_HEADS];

    // workspace
    float* input = nullptr;       // [S,3,H,H]
    float* conv0 = nullptr;       // [S,H/2,H/2,64]
    float* block[smg::kNumBlocks] = {nullptr, nullptr, nullptr, nullptr};
    float* bott = nullptr;        // [S,H/4,H/4,128]
    double* stats = nullptr;      // stats arena (zeroed once per forward)
    size_t stats_doubles_per_sample = 0;
    size_t stats_bytes = 0;
    // offsets (in double2 units, per sample) into the stats arena
    size_t st_conv0 = 0;
    size_t st_block[smg::kNumBlocks] = {0, 0, 0, 0};
    size_t st_bott = 0;  // + layer_index*128, layer_index over all 58 dense layers
    // head workspace
    float* head_scale = nullptr;  // [S,1024]
    float* head_shift = nullptr;  // [S,1024]
    float* head_p = nullptr;      // [S,400,64]
    float* scene_tmp = nullptr;   // [3,H,H] staging for smg_qforward_maps
    float* head_bn1 = nullptr;    // [pairs][2][64] batch mean / biased var of the head's BN(64), last head pass
    size_t head_bn1_floats = 0;
    int head_bn1_pairs = 0;
    // CUDA-graph replay of smg_qforward_maps: the pass is a fixed sequence of ~250 launches on fixed workspace
    // addresses, so it is captured once per (trunk, head, shapes, rotations, precision) and replayed
    double* hm_stage = nullptr;   // [1 + S][(H/2)^2] heightmap staging (scene, then masks)
    float* q_stage = nullptr;     // [S*S*4] result staging
    float* bn_stage = nullptr;    // [2][max_samples][SMG_TRUNK_BN_CHANNELS] statistics staging of the captured passes (lazy)
    cudaStream_t gstream = nullptr;
    cudaEvent_t g_in = nullptr, g_out = nullptr;
    bool use_graphs = true;
    struct QGraph {
        int trunk_id, head_id, n_masks, n_rot, num_rot, hm_size, precision, groups;
        double mean, stddev;
        std::vector<int> rots;
        int stats = 0;            // 1: the captured pass also exports the per-sample BatchNorm statistics (bn_stage)
        int head_pairs = 0;       // head_bn1_pairs of the captured pass (host-side state the replay must restore)
        int seen = 0;
        int64_t n_launches = 0;   // kernels inside the captured graph (for smg_launch_count)
        cudaGraphExec_t exec = nullptr;
    };
    std::vector<QGraph> graphs;
    int last_n = 0;               // samples of the last trunk forward (for smg_debug_read)
    double* geo_out = nullptr;        // result staging of smg_geometry
    void* bn_regions_dev = nullptr;   // the 121 BatchNorm statistics regions in module order (export_bn_stats)
    int bn_regions = 0;

    // training workspace (2 samples: rotated scene + masked scene), allocated by the first smg_qforward_train
    struct TrainWs {
        void* arena = nullptr;
        size_t bytes = 0;
        float* bott_saved[58] = {nullptr};   // conv1 output of every dense layer [2,Hb,Hb,128]
        float* dblk[smg::kNumBlocks] = {nullptr, nullptr, nullptr, nullptr};  // gradient w.r.t. the raw block buffers
        float* dconv0 = nullptr;             // [2,H/2,H/2,64]
        float* t_a = nullptr;                // [2,H/4,H/4,128] scratch (d relu(bn2))
        float* t_b = nullptr;                // [2,H/4,H/4,128] scratch (d conv1 output)
        float* t_c = nullptr;                // [2,H/4,H/4,256] scratch (d relu(bn1) / transition dgrad)
        float* dy1_saved[58] = {nullptr};    // gradient w.r.t. conv1's output of every dense layer [2,Hb,Hb,128] (side-stream wgrad)
        cudaStream_t wstream = nullptr;      // the weight gradients' stream
        cudaEvent_t ev[60] = {nullptr};      // [0,58) per-layer dependencies, 58 fork, 59 join
        float* dP = nullptr;                 // [400,64]
        float* wg3_scratch = nullptr;        // [58][(tap, co) 288][ci 128]: 3x3 weight gradients before the final transpose
        std::vector<void*> wg3_jobs;         // host copy of the {scratch, gradient tensor} table of wgrad3_finish_kernel
        void* wg3_jobs_dev = nullptr;
        double* sums = nullptr;              // BN backward reductions, zeroed once per backward
        size_t sums_bytes = 0;
        bool valid = false;                  // a training forward is waiting for its backward
        int64_t pass_id = 0;                 // stamp of the last smg_qforward_train (smg_train_pass_id)
        int trunk_id = -1, head_id = -1, in_channels = 3;
    } train;

    // smg_train_step: device tables of the caller's tensors, per-call scalars, staged outputs, captured steps
    struct StepGraph {
        uint64_t sig;
        int seen;
        int64_t n_launches;
        cudaGraphExec_t exec;
    };
    struct StepState {
        void* tables = nullptr;
        size_t tables_bytes = 0;
        uint64_t tables_sig = 0;
        int tables_n_out = 0;
        float** d_params = nullptr;
        float** d_grads = nullptr;
        float** d_m = nullptr;
        float** d_v = nullptr;
        uint8_t* d_chunks = nullptr;
        int n_chunks = 0;
        float* dyn = nullptr;      // [0] label, [1] 1 - beta1^t, [2] sqrt(1 - beta2^t)
        float* out = nullptr;      // [0..3] Q / logits, [4] loss, [8..11] dLoss/dQ
        float* bn_mean = nullptr;  // [2][SMG_TRUNK_BN_CHANNELS] batch statistics of the two passes
        float* bn_var = nullptr;
        std::vector<float*> host_grads;
        std::vector<StepGraph> graphs;
        // smg_adam_step's own tables (any tensor list)
        void* adam_tables = nullptr;
        size_t adam_tables_bytes = 0;
        uint64_t adam_sig = 0;
        int adam_chunks = 0;
    } step;

    // optional per-kernel-class timing (bench.py roofline): CUDA events around every launch of a class
    bool profile = false;
    struct ProfRec {
        int cls;
        cudaEvent_t e0, e1;
        double flops, bytes;
    };
    std::vector<ProfRec> prof;
};

namespace smg {

inline double* stats_ptr(const smg_handle* h, size_t off_double2) {
    return h->stats + 2 * off_double2 * (size_t)h->max_samples;
}
// stats layout: for a region with C channels: [S][C] double2, region base = off * S

// makes the handle's device current for the lifetime of an ABI call
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// ---- kernels (defined in the .cu files) ---------------------------------------------
// K1
int launch_prep(smg_handle* h, const double* hm, int n, int hm_size, double mean, double stddev, float* out,
                int channels, cudaStream_t st);
int launch_rotate(smg_handle* h, const float* in, const int* host_rot, int n_rot, int num_rot, float* out,
                  int channels, cudaStream_t st);
int launch_rotate_index_map(smg_handle* h, int rot, int num_rot, int32_t* out, cudaStream_t st);
// the whole input stage of a heightmap pass in one launch: [groups x n_rot] rotated scenes then n_mask_samples masked scenes
int launch_prep_rotate(smg_handle* h, const double* scene_hm, int groups, const int* host_rot, int n_rot, int num_rot,
                       const double* mask_hm, int n_mask_samples, int hm_size, double mean, double stddev, float* out, cudaStream_t st);
int launch_resize_masks(smg_handle* h, const float* in, int n, int hin, int hout, float* out, cudaStream_t st);
// stem
int launch_conv0(smg_handle* h, const float* in, int cin, int n, const float* w, float* out, double* stats,
                 cudaStream_t st);
int launch_conv0_umma(smg_handle* h, const float* in, int n, const float* w_umma, float* out, double* stats, cudaStream_t st);
int pack_conv0_umma(smg_handle* h, const float* folded, float* out, cudaStream_t st);
int launch_pool0(smg_handle* h, int n, const float* conv0, const double* stats_in, const float* gamma,
                 const float* beta, float* out, int out_cstride, double* stats_out, cudaStream_t st);

// generic convolution (1x1 / 3x3 / pooled 1x1) with BN-ReLU prologue and stats epilogue
struct ConvArgs {
    const float* in = nullptr;  // NHWC, pre-BN
    int in_cstride = 0;         // floats per pixel
    int cin = 0;
    int hin = 0;                // input spatial size (square)
    // prologue: mode 0 = derive scale/shift from (stats, gamma, beta); mode 1 = read scale/shift [S][cin]
    int prologue_mode = 0;             // 2 = identity (no scale/shift; used by the data-gradient convolutions)
    const double* in_stats = nullptr;  // [S][in_stats_stride] double2
    int in_stats_stride = 0;
    const float* gamma = nullptr;
    const float* beta = nullptr;
    const float* scale = nullptr;
    const float* shift = nullptr;
    int relu = 1;
    int pool = 0;  // 1: average the 2x2 window of prologue outputs (transition); output spatial = hin/2
    int taps = 1;  // 1 or 9 (3x3, pad 1)
    const ConvW* w = nullptr;
    const float* w_raw = nullptr;  // conv_ffma only: overrides w->w_ffma ([taps][cin][cout] fp32)
    const uint8_t* w_umma = nullptr;  // conv_umma only (tf32): overrides w->w_tf32 with another packed stage image (w_dgrad_tf32)
    // conv_umma only: fused BatchNorm-backward reduction in the epilogue of a data-gradient convolution (see UmmaDev)
    const float* bnr_x = nullptr;
    int bnr_x_cstride = 0;
    const double* bnr_stats = nullptr;
    int bnr_stats_stride = 0;
    const float* bnr_gamma = nullptr;
    const float* bnr_beta = nullptr;
    double* bnr_sums = nullptr;
    float* out = nullptr;
    int out_cstride = 0;
    int out_coff = 0;
    int cout = 0;
    double* out_stats = nullptr;  // [S][out_stats_stride] double2, at channel out_coff; may be null
    int out_stats_stride = 0;
    int n = 0;  // samples
};
int launch_conv_ffma(smg_handle* h, const ConvArgs& a, cudaStream_t st);
int launch_conv_umma(smg_handle* h, const ConvArgs& a, int precision, cudaStream_t st);
int ensure_dyn_smem(smg_handle* h, const void* kernel, int bytes);   // tma_common.cu
int launch_conv3_wt(smg_handle* h, const ConvArgs& a, cudaStream_t st);
int launch_conv1_t(smg_handle* h, const ConvArgs& a, cudaStream_t st);
int launch_trans_t(smg_handle* h, const ConvArgs& a, cudaStream_t st);

// head
int launch_head_prepare(smg_handle* h, int n, const double* stats4, int stats_stride, const BnP& norm5,
                        const BnP& hnorm0, int half, float* scale, float* shift, cudaStream_t st);
int launch_norm5_export(smg_handle* h, int n, const float* block4, const double* stats4, int stats_stride,
                        const BnP& norm5, float* out_nchw, cudaStream_t st);
int launch_head_tail(smg_handle* h, const float* p_scene, const float* p_mask, int n_rot, int n_masks, const HeadW& hw, float* q,
                     cudaStream_t st, int groups = 1);
// one BatchNorm layer's slice of the statistics arena: `count` channels of a [S][stride] double2 region -> columns
// [out_off, out_off + count) of the exported [S][total] mean / biased-variance tables
struct BnRegion {
    const double* stats;
    int stride, count;
    double cnt;
    int out_off;
};
int launch_bn_export_all(smg_handle* h, int n, const void* dev_regions, int n_regions, int total, float* mean, float* var,
                         cudaStream_t st);
int launch_argmax(smg_handle* h, const float* q, int n, float* out, int32_t* out_idx, cudaStream_t st);
int launch_nhwc_to_nchw(smg_handle* h, const float* in, int hw, int c, int cstride, float* out, cudaStream_t st);

// PE / OO geometry (geometry.cu)
int launch_geometry(smg_handle* h, int mode, const double* dev_depth, int img_h, int img_w, const double* A, const double* K,
                    const double* P, const double* boxes, const double* centers, int n, int best, int flag, const double* pix,
                    double* dev_out, cudaStream_t st);
// K11 / K12
int launch_heightmap_color(smg_handle* h, const uint8_t* color, uint8_t* out224, uint8_t* out448, cudaStream_t st);
int launch_heightmap(smg_handle* h, const double* depth, const double* K, const double* pose, double* out224,
                     double* out448, double* host_A_htor, cudaStream_t st);
int launch_nms(smg_handle* h, const float* boxes, int n, float thr, float amin, float amax, int32_t* keep,
               int32_t* n_keep, cudaStream_t st);

// ---- backward (backward.cu) ------------------------------------------------------------
struct BnBwd {
    const float* da;       // gradient w.r.t. the BN(+ReLU) output, NHWC [S, hw_da^2, da_cstride]
    int da_cstride;
    int da_pooled;         // 1: da lives at half resolution and is spread over the 2x2 window (x 0.25)
    const float* x;        // raw BN input, NHWC [S, hw^2, x_cstride]
    int x_cstride;
    const double* stats;   // (sum, sumsq) of x: [S, stats_stride] double2
    int stats_stride;
    const float* gamma;
    const float* beta;
    int C, hw, relu;
    double* sums;          // [S, C] double2 (S1, S2)
    float* dst;            // apply: gradient w.r.t. x, NHWC [S, hw^2, dst_cstride]
    int dst_cstride;
    int accumulate;
    int pix_per_cta;
};

struct Wgrad {
    const float* g;        // NHWC [S, hout^2, g_cstride], channels [g_coff, g_coff+cout)
    int g_cstride, g_coff, cout;
    const float* x;        // raw activation source NHWC [S, hin^2, x_cstride], channels [0, cin)
    int x_cstride, cin, hin, hout;
    int prologue_mode;     // 0 stats, 1 scale/shift arrays
    const double* stats;
    int stats_stride;
    const float* gamma;
    const float* beta;
    const float* scale;
    const float* shift;
    int relu;
    float* dw;             // torch OIHW [cout][k_total][taps], this conv occupies input channels [k_off, k_off+cin)
    int k_total, k_off;
    int pix_per_cta, chunks_per_sample;
};

// reduce (apply = false) then apply (apply = true; also writes dgamma / dbeta [C] from the completed reductions)
int launch_bn_bwd(smg_handle* h, BnBwd a, int S, bool apply, float* dgamma, float* dbeta, cudaStream_t st);
int launch_wgrad(smg_handle* h, Wgrad a, int S, int taps, int pool, cudaStream_t st);
int launch_pool0_bwd(smg_handle* h, int S, const float* g, int g_cstride, const float* conv0, const double* stats,
                     const float* gamma, const float* beta, float* da0, cudaStream_t st);
int launch_conv0_wgrad(smg_handle* h, int S, const float* d, const float* in, int cin, float* dw, cudaStream_t st);
int launch_replicate_conv0_grad(smg_handle* h, const float* g1, float* g3, cudaStream_t st);   // [64][49] -> [64][3][49]
int launch_head_tail_bwd(smg_handle* h, const float* p, const HeadW& hw, const float* dq, float* dP, float* dg1,
                         float* db1, float* dw1, cudaStream_t st);
int launch_head_norm_bwd(smg_handle* h, const float* da0, const float* x4, const double* stats, int stats_stride,
                         const BnP& norm5, const BnP& hnorm0, float* dx4, float* dg5, float* db5, float* dgh, float* dbh,
                         cudaStream_t st);
// tensor-core weight gradients of the dense layers (wgrad_umma.cu); SMG_ERR_UNSUPPORTED for shapes they do not serve
int launch_wgrad1_umma(smg_handle* h, const float* g, const float* x, int x_cstride, int cin, int hw, int S, const double* stats,
                       int stats_stride, const float* gamma, const float* beta, float* dw, cudaStream_t st);
int launch_wgrad3_umma(smg_handle* h, const float* g, int g_cstride, int g_coff, const float* y, int hw, int S,
                       const double* stats, int stats_stride, const float* gamma, const float* beta, float* scratch,
                       cudaStream_t st);
int launch_wgrad3_finish(smg_handle* h, const void* dev_jobs, int n_jobs, cudaStream_t st);

// schedule pieces shared by api.cu and train.cu
int conv_dispatch(smg_handle* h, const ConvArgs& a, cudaStream_t st);
int trunk_forward(smg_handle* h, int trunk_id, int n, int in_channels, cudaStream_t st, bool save_bott = false);
int heads_forward(smg_handle* h, int trunk_id, int head_id, int n_rot, int n_masks, float* dev_q, cudaStream_t st,
                  int groups = 1);
int head_partials(smg_handle* h, int trunk_id, int head_id, int n_rot, int n_masks, cudaStream_t st, int groups = 1);
int export_bn_stats(smg_handle* h, int n, float* mean, float* var, cudaStream_t st);
// ONE launch of torch.optim.Adam over a list of tensors (train.cu); the device tables are cached per pointer set
int adam_multi_tensor(smg_handle* h, float* const* params, const float* const* grads, float* const* m, float* const* v,
                      const int64_t* numel, int n, int step, float lr, float b1, float b2, float eps, cudaStream_t st);
int ensure_train_workspace(smg_handle* h);
int qbackward_impl(smg_handle* h, const float* dq, float* const* tg, float* const* hg, cudaStream_t st);
int repack_trunk(smg_handle* h, int trunk_id, cudaStream_t st);   // kernels only (graph-capturable), tables from the last set
int repack_head(smg_handle* h, int head_id, cudaStream_t st);

// weight packing
enum { SMG_PACK_FFMA = 1, SMG_PACK_TF32 = 2, SMG_PACK_BF16 = 4, SMG_PACK_DGRAD = 8, SMG_PACK_ALL = 15 };
__host__ __device__ inline int dgrad_cin_padded(int cin, int taps) { return taps == 9 ? cin : (cin + 127) / 128 * 128; }
PackJob make_pack_job(const float* w_oihw, const ConvW& cw, int k_offset, int k_total);
int launch_pack_tables(smg_handle* h, const PackJob* dev_pj, int n_pj, const CopyJob* dev_cj, int n_cj, cudaStream_t st);
int pack_conv_weights(smg_handle* h, const float* w_oihw, ConvW& cw, int k_offset, int k_total, cudaStream_t st);
size_t conv_packed_bytes_ffma(int cin, int cout, int taps);
size_t conv_packed_bytes_umma(int cin, int cout, int taps, int elt_bytes);

}  // namespace smg
