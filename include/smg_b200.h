/*
 * smg_b200.h - C ABI of libsmg_b200.so, the B200 (sm_100a) implementation of the
 * SMG-multimodal-grasping grasp-affordance hot path.
 *
 * The reference (fukangl/SMG-multimodal-grasping) is pure Python and has no FFI or
 * plugin interface; its boundary for this path is the Python surface of
 * code/models.py, code/trainer.py, code/utils.py and code/NMS.py (SURVEY.md
 * section 8(b)).  The entry points below are what a ctypes binding of that surface
 * calls; each one cites the reference code it replaces.  INTEGRATION.md shows the
 * binding.
 *
 * Conventions
 *   - every function returns 0 on success, a negative smg_status otherwise;
 *     smg_last_error() returns a thread-local message.  No exceptions cross the ABI.
 *   - all `dev_*` pointers are DEVICE pointers owned by the caller (torch tensors'
 *     data_ptr()); `host_*` pointers are host memory.  `stream` is a cudaStream_t
 *     (NULL = default stream).  Work is enqueued asynchronously on `stream` unless
 *     the function is documented as synchronous.
 *   - a handle is bound to one GPU; it owns only workspace and re-packed weights.
 *     Not thread-safe per handle.  One handle per GPU per process.
 *   - "sample" = one H x H network input (a rotated scene or an object-masked
 *     scene); BatchNorm statistics are always per sample (reference: batch 1,
 *     train mode, code/trainer.py:95).
 */
#ifndef SMG_B200_H
#define SMG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct smg_handle smg_handle;

typedef enum smg_status {
    SMG_OK = 0,
    SMG_ERR_INVALID = -1, /* bad argument */
    SMG_ERR_CUDA = -2,    /* CUDA runtime error (message in smg_last_error) */
    SMG_ERR_STATE = -3,   /* weights not set, workspace too small, ... */
    SMG_ERR_UNSUPPORTED = -4
} smg_status;

/* arithmetic mode of the trunk / head convolutions */
typedef enum smg_precision {
    SMG_PREC_FP32 = 0, /* fp32 accuracy (parity <= 1e-4 vs the fp32 reference): tcgen05 kind::tf32 with hi/lo split operands
                          ("3xTF32": a_lo b_hi + a_hi b_lo + a_hi b_hi, fp32 accumulate); SMG_FP32_TC=0 selects CUDA-core FFMA */
    SMG_PREC_TF32 = 1, /* tcgen05 kind::tf32, fp32 accumulate in TMEM */
    SMG_PREC_BF16 = 2  /* tcgen05 kind::f16 (bf16 operands), fp32 accumulate in TMEM */
} smg_precision;

/* which DenseNet-121 trunk / which head (code/models.py:308-310, :316-343) */
enum { SMG_TRUNK_SUCTION = 0, SMG_TRUNK_GRASP = 1, SMG_TRUNK_GS = 2, SMG_NUM_TRUNKS = 3 };
enum { SMG_HEAD_SUCTION = 0, SMG_HEAD_GRASP = 1, SMG_HEAD_GS = 2, SMG_NUM_HEADS = 3 };

#define SMG_TRUNK_NUM_PARAMS 362 /* conv/BN affine tensors of densenet121().features, state_dict order */
#define SMG_HEAD_NUM_PARAMS 6    /* norm0.w, norm0.b, conv0.w, norm1.w, norm1.b, conv1.w */
#define SMG_TRUNK_BN_CHANNELS 41824 /* sum of num_features over the 121 BatchNorm2d of one trunk */

int smg_version(void);
const char* smg_last_error(void);

/* ---- lifetime ------------------------------------------------------------------ */
/* Allocate workspace for up to `max_samples` samples of H x H input (H multiple of 32;
 * the heads need H = 640, code/models.py:322).  max_samples = 0 creates a handle without
 * trunk workspace for the stateless kernels (smg_heightmap*, smg_nms, smg_argmax,
 * smg_adam_step, smg_geometry_*).  The training workspace is allocated by the first
 * smg_qforward_train.                                                              */
int smg_create(int device, int max_samples, int H, smg_handle** out);
int smg_destroy(smg_handle* h);
int smg_set_precision(smg_handle* h, int precision);
int smg_get_precision(smg_handle* h);
/* bytes of device workspace held by the handle */
int64_t smg_workspace_bytes(smg_handle* h);

/* ---- weights --------------------------------------------------------------------
 * Re-pack fp32 parameters (device pointers, contiguous, torch layouts) into the
 * kernel layouts.  Replaces what `model.cuda()` / `load_state_dict` do for the
 * reference (code/trainer.py:85-92); call again after every optimizer step.
 * Trunk parameter order (n = SMG_TRUNK_NUM_PARAMS): conv0.weight, norm0.weight,
 * norm0.bias, then per dense layer norm1.w, norm1.b, conv1.w, norm2.w, norm2.b,
 * conv2.w, per transition norm.w, norm.b, conv.w (in state_dict order), norm5.w, norm5.b. */
int smg_set_trunk_weights(smg_handle* h, int trunk_id, const float* const* dev_params, int n, void* stream);
/* Head parameter order (SMG_HEAD_NUM_PARAMS): norm0.w[2048], norm0.b, conv0.w[64,2048,1,1],
 * norm1.w[64], norm1.b, conv1.w[n_out,64,20,20]; n_out = 1 (reinforcement) or 3 (reactive). */
int smg_set_head_weights(smg_handle* h, int head_id, const float* const* dev_params, int n_out, void* stream);

/* Which kernel layouts smg_set_*_weights writes (bit 0 fp32 FFMA, 1 tf32 tcgen05, 2 bf16 tcgen05, 3 data-gradient);
 * default all.  Restricting it to the active precision (+ bit 3 when training) makes the per-step re-pack cheaper.
 * Weights must be set again after the mask or the precision changes.                                            */
int smg_set_pack_layouts(smg_handle* h, int mask);

/* ---- K1: input stage ------------------------------------------------------------
 * smg_prep: code/trainer.py:165-191.  224x224 float64 heightmaps -> float32 [n,3,H,H]
 * (nearest zoom x2, zero pad to H, (x-mean)/std, 3 identical channels).             */
int smg_prep(smg_handle* h, const double* dev_heightmaps, int n, int hm_size, double mean, double stddev,
             float* dev_out, void* stream);
/* smg_rotate: code/models.py:371-382.  F.affine_grid(align_corners=True) +
 * F.grid_sample(mode='nearest', zero padding) of one [3,H,H] image for each listed
 * rotation index (angle = idx * 360/num_rotations degrees) -> [n_rot,3,H,H].        */
int smg_rotate(smg_handle* h, const float* dev_in, const int* host_rot_idx, int n_rot, int num_rotations,
               float* dev_out, void* stream);
/* same index arithmetic, returned as source linear indices (-1 = outside): test hook  */
int smg_rotate_index_map(smg_handle* h, int rot_idx, int num_rotations, int32_t* dev_out, void* stream);

/* ---- trunk / head ---------------------------------------------------------------
 * smg_trunk_forward: `trunk.features(x)` (code/models.py:384-385) for n samples
 * [n,3,H,H] -> dev_feat [n,1024,H/32,H/32] (NCHW float32, after norm5).  If
 * dev_bn_mean/dev_bn_var are non-NULL they receive the per-sample batch statistics
 * of all 121 BatchNorm layers, [n, SMG_TRUNK_BN_CHANNELS] each, in module order
 * (biased variance), from which the caller applies the running-stat EMA.            */
int smg_trunk_forward(smg_handle* h, int trunk_id, const float* dev_in, int n, float* dev_feat,
                      float* dev_bn_mean, float* dev_bn_var, void* stream);

/* smg_qforward: the fused, de-duplicated Q pass.  One scene image and n_masks masked
 * images ([3,H,H] each); evaluates head(cat(trunk(rotate(scene, r)), trunk(mask_k))) for
 * every listed rotation r and every mask k (code/models.py:371-389; the reference
 * recomputes the mask pass per rotation and the scene passes per object).
 * dev_q receives [n_masks, n_rot, n_out] float32.  Optional dev_bn_mean/var:
 * [n_rot + n_masks, SMG_TRUNK_BN_CHANNELS] (rotations first, then masks).            */
int smg_qforward(smg_handle* h, int trunk_id, int head_id, const float* dev_scene, const float* dev_masks,
                 int n_masks, const int* host_rot_idx, int n_rot, int num_rotations, float* dev_q,
                 float* dev_bn_mean, float* dev_bn_var, void* stream);
/* ---- the Q pass split for multi-GPU decisions -----------
 * The head's BN(2048)+ReLU+1x1 conv acts on cat(scene features, mask features): its first 1024 input channels depend only
 * on the scene sample, the last 1024 only on the mask sample.  smg_qpartials runs pre-processing, rotation and the trunk for
 * the LISTED rotations of the scene (n_rot may be 0) and the given masked heightmaps (n_masks may be 0) and returns the
 * per-sample partial products P [n_rot + n_masks, 400, 64] (rotations first); any rank can then form Q for every
 * (mask, rotation) pair from gathered partials with smg_qcombine: dev_q [n_masks, n_rot, n_out].  Results are identical to
 * smg_qforward_maps (BatchNorm is per sample).  The gather between the two is the path's data exchange (NCCL).       */
int smg_qpartials(smg_handle* h, int trunk_id, int head_id, const double* dev_scene_hm, const int* host_rot_idx, int n_rot,
                  int num_rotations, const double* dev_mask_hms, int n_masks, int hm_size, double mean, double stddev,
                  float* dev_p, void* stream);
int smg_qcombine(smg_handle* h, int head_id, const float* dev_p_scene, int n_rot, const float* dev_p_mask, int n_masks,
                 float* dev_q, void* stream);

/* same, starting from 224x224 float64 heightmaps (fuses smg_prep): Trainer.forward,
 * code/trainer.py:162-207.                                                          */
int smg_qforward_maps(smg_handle* h, int trunk_id, int head_id, const double* dev_scene_hm,
                      const double* dev_mask_hms, int n_masks, int hm_size, double mean, double stddev,
                      const int* host_rot_idx, int n_rot, int num_rotations, float* dev_q, float* dev_bn_mean,
                      float* dev_bn_var, void* stream);

/* `groups` independent units in one batch: scene g [hm,hm] with its n_masks masked heightmaps [groups, n_masks, hm, hm];
 * dev_q receives [groups, n_masks, n_rot, n_out].  Every unit is evaluated exactly as by smg_qforward_maps (BatchNorm
 * statistics are per sample), the batch only gives the late, small layers more CTAs per launch.                   */
int smg_qforward_maps_batch(smg_handle* h, int trunk_id, int head_id, const double* dev_scene_hms,
                            const double* dev_mask_hms, int groups, int n_masks, int hm_size, double mean, double stddev,
                            const int* host_rot_idx, int n_rot, int num_rotations, float* dev_q, float* dev_bn_mean,
                            float* dev_bn_var, void* stream);

/* Batch statistics of the head's BatchNorm2d(64) from the LAST smg_qforward* / smg_qforward_train call on this handle:
 * dev_out [n_pairs, 2, 64] float32 = (mean, biased variance) per (mask, rotation) pair in q order.  Together with the
 * trunk statistics (norm5's mean/var determine the statistics of the head's BatchNorm2d(2048): mean = norm5.bias,
 * var = gamma5^2 var/(var+eps)) this lets the caller apply the reference's running-stat updates to the head.      */
int smg_head_bn_stats(smg_handle* h, float* dev_out, int n_pairs, void* stream);

/* ---- training (code/trainer.py:278-384) -----------------------------------------
 * smg_qforward with n_masks = n_rot = 1 and save_for_backward; then smg_qbackward
 * with dq = dLoss/dQ [n_out] produces gradients for every trunk / head parameter in
 * the smg_set_*_weights order (device pointers to caller-owned float32 buffers of
 * the parameter shapes; gradients are WRITTEN, not accumulated).                     */
int smg_qforward_train(smg_handle* h, int trunk_id, int head_id, const float* dev_scene, const float* dev_mask,
                       int rot_idx, int num_rotations, float* dev_q, float* dev_bn_mean, float* dev_bn_var,
                       void* stream);
int smg_qbackward(smg_handle* h, const float* dev_dq, float* const* dev_trunk_grads, float* const* dev_head_grads,
                  void* stream);
/* ---- the whole training step in one call (code/trainer.py:338-383) ----------------
 * Trainer.backprop's device work: Trainer.forward pre-processing of the two 224x224 float64 heightmaps (scene and masked
 * scene), the grad-enabled Q pass at rotation rot_idx, the loss (0: hand-written Huber, delta 1, on Q - label,
 * trainer.py:345-348; 1: CrossEntropyLoss2d with class weights on the 3 logits, trainer.py:284-299), backward, Adam
 * (torch.optim.Adam semantics, trainer.py:99) and the re-pack of the updated weights.  The whole sequence is captured in a
 * CUDA graph per configuration and replayed.
 * dev_params / dev_grads / dev_exp_avg / dev_exp_avg_sq: n_tensors = SMG_TRUNK_NUM_PARAMS + SMG_HEAD_NUM_PARAMS device
 * pointers (trunk tensors in smg_set_trunk_weights order, then the head's) to caller-owned float32 buffers of the parameter
 * shapes; dev_params must be the tensors the packed weights were set from.  Gradients are WRITTEN, parameters and moments
 * updated in place.  dev_loss [1], dev_q [n_out]; dev_bn_mean / dev_bn_var optional [2, SMG_TRUNK_BN_CHANNELS]
 * (scene pass, then mask pass) for the caller's running-statistics update.                                              */
typedef struct smg_train_step_args {
    int32_t trunk_id, head_id;      /* routing of the sample's primitive (code/models.py:513-586) */
    int32_t rot_idx, num_rotations; /* rotation of the scene pass: angle = rot_idx * 360 / num_rotations */
    int32_t hm_size;                /* the heightmaps are hm_size x hm_size float64 */
    int32_t loss_kind;              /* 0 Huber on the scalar Q, 1 weighted cross-entropy on 3 logits */
    int32_t adam_step;              /* 1-based step count of this update (bias corrections) */
    float label;                    /* target value (loss 0) or class index (loss 1) */
    double mean, stddev;            /* (x - mean) / stddev of Trainer.forward */
    float class_weight[3];          /* loss 1 only */
    float lr, beta1, beta2, eps;    /* Adam hyper-parameters */
    int32_t flags;                  /* SMG_STEP_* */
} smg_train_step_args;
/* gradients only: no Adam update, no re-pack (data-parallel training accumulates / all-reduces the gradients of many
 * samples first and then calls smg_adam_step once) */
#define SMG_STEP_GRADS_ONLY 1
int smg_train_step(smg_handle* h, const smg_train_step_args* args, const double* dev_scene_hm, const double* dev_mask_hm,
                   float* const* dev_params, float* const* dev_grads, float* const* dev_exp_avg,
                   float* const* dev_exp_avg_sq, int n_tensors, float* dev_loss, float* dev_q, float* dev_bn_mean,
                   float* dev_bn_var, void* stream);

/* Stamp of the pending smg_qforward_train result (incremented by every such call), or -1 if there is none - any
 * other forward on the handle overwrites the saved activations and invalidates it; smg_qbackward then fails with
 * SMG_ERR_STATE instead of differentiating another pass.  Only ONE grad-enabled pass may be in flight per handle. */
int64_t smg_train_pass_id(smg_handle* h);
/* torch.optim.Adam (code/trainer.py:99,383) over a list of tensors in ONE launch; `step` is the 1-based update count */
int smg_adam_step(smg_handle* h, float* const* dev_params, const float* const* dev_grads, float* const* dev_m,
                  float* const* dev_v, const int64_t* host_numel, int n_tensors, int step, float lr, float beta1,
                  float beta2, float eps, void* stream);

/* ---- K9: action argmax (code/main.py:167-173,194-195) ----------------------------
 * dev_q [n] float32 -> dev_out[0] = max value, dev_out_idx[0] = first index of the max
 * (np.argmax first-max-wins).                                                       */
int smg_argmax(smg_handle* h, const float* dev_q, int n, float* dev_out, int32_t* dev_out_idx, void* stream);

/* ---- K11: heightmap (code/utils.py:12-68, depth path) -----------------------------
 * depth [480,640] float64 (metres), K 3x3 and pose 4x4 float64 row-major on the HOST;
 * writes depth_heightmap [224,224] and depth_mask [448,448] float64, bit-exact with
 * the reference's numpy + cv2.warpPerspective result; host_A_htor (optional, 9 doubles)
 * receives cv2.getPerspectiveTransform(dst_heightmap, src).                          */
int smg_heightmap(smg_handle* h, const double* dev_depth, const double* host_K, const double* host_pose,
                  double* dev_out224, double* dev_out448, double* host_A_htor, void* stream);

/* Colour outputs of the same function (code/utils.py:62,64): cv2.warpPerspective of the uint8 camera image [480,640,3]
 * with the two transforms above -> color_heightmap [224,224,3] and color_mask [448,448,3] uint8, bit-exact with cv2's
 * 15-bit fixed-point bilinear remap (consumed by Mask R-CNN / logging, not by the Q pass).                            */
int smg_heightmap_color(smg_handle* h, const uint8_t* dev_color, uint8_t* dev_out224, uint8_t* dev_out448, void* stream);

/* ---- detector post-processing in front of the Q pass (code/masks.py:51) -----------
 * F.interpolate(masks, size=[size_out, size_out], mode="bilinear", align_corners=True) of n float32 soft masks
 * [n, size_in, size_in] (the reference resizes Mask R-CNN's 448x448 masks to the 224x224 heightmap grid).          */
int smg_resize_masks(smg_handle* h, const float* dev_masks, int n, int size_in, int size_out, float* dev_out, void* stream);

/* ---- PE / OO geometry after the argmax (code/utils.py:70-81, :316-366, :370-612) ----
 * mode 0 global_position(host_pix = (_, row, col)) -> host_out[0..2] robot xyz;
 * mode 1 get_best_grasp_angle(is_pe = flag, box_mask_cors, bestg_id[0] = best_id) -> xyz, [3] jaw angle (rad), [4] opening (m);
 * mode 2 get_best_suction_angle(is_oo = flag, objects_number = n_objects, masks_cter, box_mask_cors, bests_id[0] = best_id)
 *        -> xyz, [3] np.deg2rad(chosen direction).
 * dev_depth: the camera depth image [img_h, img_w] float64 on the device; host_A_htor 3x3, host_K 3x3, host_pose 4x4 row-major;
 * host_boxes [n,4,2] (x, y) min-area box corners, host_centers [n,2] (x, y); host_out 5 doubles.  Synchronous.           */
int smg_geometry(smg_handle* h, int mode, const double* dev_depth, int img_h, int img_w, const double* host_A_htor,
                 const double* host_K, const double* host_pose, const double* host_boxes, const double* host_centers,
                 int n_objects, int best_id, int flag, const double* host_pix, double* host_out, void* stream);

/* ---- K12: box NMS (code/NMS.py:8-59) ---------------------------------------------
 * boxes [n,2,2] float32 ((x1,y1),(x2,y2)); keeps index-order greedy survivors.
 * dev_keep [n] int32 receives kept indices, dev_n_keep[0] their count.  n <= 1024.   */
int smg_nms(smg_handle* h, const float* dev_boxes, int n, float co_thresh, float min_area, float max_area,
            int32_t* dev_keep, int32_t* dev_n_keep, void* stream);

/* ---- introspection --------------------------------------------------------------- */
/* number of kernels this library has launched on behalf of the handle since creation */
int64_t smg_launch_count(smg_handle* h);
/* copy an internal activation for tests: what = "conv0", "pool0", "block1".."block4",
 * "trans1".."trans3" (NHWC float32 of the last forward, sample `sample`) -> NCHW.     */
int smg_debug_read(smg_handle* h, const char* what, int sample, float* dev_out_nchw, int64_t capacity_floats,
                   void* stream);

/* Per-kernel-class device timing for roofline reporting.  While enabled, every launch of a class is
 * bracketed by CUDA events on the launching stream.  Classes: 0 stem (conv0, norm0+pool0), 1 conv 1x1
 * (dense-layer bottlenecks, transitions, head), 2 conv 3x3, 3 everything else.  smg_profile_read
 * synchronises, then returns per class the summed milliseconds, launch count, algorithmic FLOPs and
 * algorithmic bytes since smg_profile_enable(h, 1), and clears the records.                          */
#define SMG_PROFILE_CLASSES 4
int smg_profile_enable(smg_handle* h, int enable);
int smg_profile_read(smg_handle* h, double* host_ms, int64_t* host_launches, double* host_flops, double* host_bytes);

/* run ONE convolution of the generic conv kernel on caller tensors (unit-test hook for conv_ffma.cu /
 * conv_umma.cu): in NHWC [n,hin,hin,in_cstride] (channels [0,cin) used), prologue a = relu?(x*scale+shift)
 * with scale/shift [n,cin], optional 2x2 average pool of the prologue output, taps 1 (1x1) or 9 (3x3 pad 1),
 * torch OIHW weights [cout,cin,k,k]; writes channels [out_coff,out_coff+cout) of out NHWC
 * [n,hout,hout,out_cstride] and, if non-NULL, ACCUMULATES (sum, sumsq) into dev_out_stats [n,out_cstride,2].
 * Synchronous.                                                                                          */
int smg_debug_conv(smg_handle* h, int precision, const float* dev_in, int n, int hin, int cin, int in_cstride,
                   const float* dev_scale, const float* dev_shift, int relu, int pool, int taps,
                   const float* dev_w_oihw, int cout, float* dev_out, int out_cstride, int out_coff,
                   double* dev_out_stats, void* stream);

/* unit-test hook for the data-gradient convolutions of the backward pass: g NHWC [n,hin,hin,g_cstride] (channels
 * [g_coff, g_coff+cout) used) is the gradient w.r.t. the OUTPUT of a convolution with torch OIHW weights [cout,cin,k,k]
 * (k = 1 or 3, pad k/2); writes the gradient w.r.t. its input, NHWC [n,hin,hin,cin].  tf32: tcgen05 kernel on the
 * w_dgrad_tf32 image; fp32: CUDA cores.  Synchronous.                                                              */
int smg_debug_dgrad(smg_handle* h, int precision, const float* dev_g, int n, int hin, int cout, int g_cstride, int g_coff,
                    int taps, const float* dev_w_oihw, int cin, float* dev_dx, void* stream);

/* unit-test hook for the tensor-core weight gradients of the dense layers (wgrad_umma.cu, tf32): g NHWC
 * [S,hw,hw,g_cstride] (channels [g_coff, g_coff+cout) used) is the gradient w.r.t. the convolution's output, x the RAW
 * input activation NHWC [S,hw,hw,x_cstride] (channels [0,cin)), normalised as relu(bn(x)) from dev_stats (sum, sumsq)
 * [S,stats_stride,2] doubles and gamma/beta.  taps = 1: cout = 128, dev_dw [128,cin]; taps = 9: cin = 128, cout = 32, dev_dw
 * [32,128,3,3].  Synchronous.                                                                                        */
int smg_debug_wgrad(smg_handle* h, int taps, const float* dev_g, int g_cstride, int g_coff, const float* dev_x, int x_cstride,
                    int cin, int hw, int S, const double* dev_stats, int stats_stride, const float* dev_gamma,
                    const float* dev_beta, float* dev_dw, void* stream);

/* unit-test hook for the BatchNorm(+ReLU) backward kernels (backward.cu): x NHWC [S,hw,hw,x_cstride] is the raw
 * BN input, dev_stats its (sum,sumsq) [S,stats_stride,2] doubles, da the gradient w.r.t. relu(bn(x)) (at half
 * resolution and spread x0.25 if da_pooled).  Writes dx into dev_dst (accumulating if requested), the two
 * reductions into dev_sums [S,C,2] doubles (must be zero on entry) and dgamma/dbeta [C].  Synchronous.        */
int smg_debug_bn_bwd(smg_handle* h, const float* dev_da, int da_cstride, int da_pooled, const float* dev_x,
                     int x_cstride, const double* dev_stats, int stats_stride, const float* dev_gamma,
                     const float* dev_beta, int C, int hw, int relu, int S, double* dev_sums, float* dev_dst,
                     int dst_cstride, int accumulate, float* dev_dgamma, float* dev_dbeta, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SMG_B200_H */
