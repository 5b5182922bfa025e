// train.cu - the whole training step of Trainer.backprop (/root/reference/code/trainer.py:338-383) as ONE device-side
// sequence, captured in a CUDA graph: pre-processing + rotation of the two heightmaps, the grad-enabled Q pass (rotated
// scene + masked scene through one trunk, one head), the loss on the scalar / the three logits, the backward pass
// (backward.cu), Adam on the 368 tensors the sample touches and the re-pack of the updated weights into the kernel layouts.
//
// Reference semantics kept: hand-written Huber with delta = 1 on Q - label (trainer.py:345-348) or the class-weighted
// cross-entropy of CrossEntropyLoss2d on [1,3,1,1] logits (trainer.py:284-299, utils.py:306-313); torch.optim.Adam with
// lr 1e-4, betas (0.9, 0.999), eps 1e-8, no weight decay (trainer.py:99), bias corrections from the 1-based step count;
// gradients are written to the caller's .grad buffers, parameters and moments are updated in place in the caller's tensors.
#include "smg_internal.cuh"

namespace smg {

namespace {

// dyn[0] = label / class index, dyn[1] = 1 - beta1^t, dyn[2] = sqrt(1 - beta2^t)
__global__ void loss_kernel(const float* __restrict__ q, int n_out, int kind, const float* __restrict__ dyn, float w0, float w1,
                            float w2, float* __restrict__ loss, float* __restrict__ dq) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const float label = dyn[0];
    if (kind == 0) {
        const float d = q[0] - label;
        if (fabsf(d) < 1.f) {
            loss[0] = 0.5f * d * d;
            dq[0] = d;
        } else {
            loss[0] = fabsf(d) - 0.5f;
            dq[0] = d > 0.f ? 1.f : -1.f;
        }
        return;
    }
    // F.nll_loss(log_softmax(logits), target, weight=w, reduction='mean'): the weighted mean over ONE pixel is
    // -w[t] * logp[t] / w[t]; d/dlogit_c = softmax_c - [c == t]   (w[t] = 0 gives 0/0 = NaN exactly like torch)
    const int t = (int)label;
    const float w[3] = {w0, w1, w2};
    float mx = q[0];
    for (int c = 1; c < n_out; ++c) mx = fmaxf(mx, q[c]);
    float se = 0.f;
    for (int c = 0; c < n_out; ++c) se += expf(q[c] - mx);
    const float lse = mx + logf(se);
    const float ratio = w[t] / w[t];
    loss[0] = -(q[t] - lse) * ratio;
    for (int c = 0; c < n_out; ++c) dq[c] = (expf(q[c] - lse) - (c == t ? 1.f : 0.f)) * ratio;
}

struct AdamChunk {
    int tensor, offset, count;
};

// torch.optim.Adam (no amsgrad, no weight decay) over a table of tensors, one CTA per chunk of <= 4096 elements
__global__ void __launch_bounds__(256)
adam_multi_kernel(float* const* __restrict__ params, const float* const* __restrict__ grads, float* const* __restrict__ exp_avg,
                  float* const* __restrict__ exp_avg_sq, const AdamChunk* __restrict__ chunks, const float* __restrict__ dyn,
                  float lr, float b1, float b2, float eps) {
    const AdamChunk c = chunks[blockIdx.x];
    float* p = params[c.tensor] + c.offset;
    const float* g = grads[c.tensor] + c.offset;
    float* m = exp_avg[c.tensor] + c.offset;
    float* v = exp_avg_sq[c.tensor] + c.offset;
    const float bc1 = dyn[1], bc2_sqrt = dyn[2];
    const float step_size = lr / bc1;
    for (int i = threadIdx.x; i < c.count; i += 256) {
        const float gi = g[i];
        const float mi = b1 * m[i] + (1.f - b1) * gi;
        const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        p[i] -= step_size * (mi / (sqrtf(vi) / bc2_sqrt + eps));
    }
}

uint64_t fnv(uint64_t hsh, const void* data, size_t bytes) {
    const uint8_t* p = reinterpret_cast<const uint8_t*>(data);
    for (size_t i = 0; i < bytes; ++i) hsh = (hsh ^ p[i]) * 1099511628211ull;
    return hsh;
}

const int64_t* trunk_param_numel(const smg_handle* h, std::vector<int64_t>& out, int n_out) {
    // element counts in the smg_set_trunk_weights / smg_set_head_weights order
    out.clear();
    out.push_back(64 * 147);
    out.push_back(64); out.push_back(64);
    for (int b = 0; b < kNumBlocks; ++b) {
        for (int l = 0; l < kBlockLayers[b]; ++l) {
            const int cin = h->geom[b].c_in + l * kGrowth;
            out.push_back(cin); out.push_back(cin); out.push_back((int64_t)kBottleneck * cin);
            out.push_back(kBottleneck); out.push_back(kBottleneck); out.push_back((int64_t)kGrowth * kBottleneck * 9);
        }
        if (b < kNumBlocks - 1) {
            const int C = h->geom[b].c_tot;
            out.push_back(C); out.push_back(C); out.push_back((int64_t)(C / 2) * C);
        }
    }
    out.push_back(kFeatC); out.push_back(kFeatC);
    out.push_back(2 * kFeatC); out.push_back(2 * kFeatC); out.push_back((int64_t)kHeadMid * 2 * kFeatC);
    out.push_back(kHeadMid); out.push_back(kHeadMid); out.push_back((int64_t)n_out * kHeadMid * kHeadK * kHeadK);
    return out.data();
}

}  // namespace

// everything between the staged inputs and the staged outputs; enqueues on `st` only (graph-capturable)
static int train_step_body(smg_handle* h, const smg_train_step_args& a, cudaStream_t st) {
    smg_handle::StepState& S = h->step;
    const size_t img = (size_t)h->H * h->H;
    const size_t hm_elems = (size_t)a.hm_size * a.hm_size;
    // Trainer.forward feeds three identical channels (code/trainer.py:178-181): one plane per sample, channel-folded conv0
    SMG_TRY(launch_prep(h, h->hm_stage, 1, a.hm_size, a.mean, a.stddev, h->scene_tmp, 1, st));
    SMG_TRY(launch_rotate(h, h->scene_tmp, &a.rot_idx, 1, a.num_rotations, h->input, 1, st));
    SMG_TRY(launch_prep(h, h->hm_stage + hm_elems, 1, a.hm_size, a.mean, a.stddev, h->input + img, 1, st));
    SMG_TRY(trunk_forward(h, a.trunk_id, 2, 1, st, true));
    SMG_TRY(heads_forward(h, a.trunk_id, a.head_id, 1, 1, S.out, st));
    h->train.trunk_id = a.trunk_id;
    h->train.head_id = a.head_id;
    h->train.in_channels = 1;
    const int n_out = h->heads[a.head_id].n_out;
    loss_kernel<<<1, 32, 0, st>>>(S.out, n_out, a.loss_kind, S.dyn, a.class_weight[0], a.class_weight[1], a.class_weight[2],
                                  S.out + 4, S.out + 8);
    h->launches++;
    SMG_TRY(qbackward_impl(h, S.out + 8, S.host_grads.data(), S.host_grads.data() + SMG_TRUNK_NUM_PARAMS, st));
    if (!(a.flags & SMG_STEP_GRADS_ONLY)) {
        adam_multi_kernel<<<S.n_chunks, 256, 0, st>>>(S.d_params, S.d_grads, S.d_m, S.d_v,
                                                      reinterpret_cast<const AdamChunk*>(S.d_chunks), S.dyn, a.lr, a.beta1, a.beta2,
                                                      a.eps);
        h->launches++;
        SMG_CUDA(cudaGetLastError());
        SMG_TRY(repack_trunk(h, a.trunk_id, st));
        SMG_TRY(repack_head(h, a.head_id, st));
    }
    SMG_TRY(export_bn_stats(h, 2, S.bn_mean, S.bn_var, st));
    return SMG_OK;
}

int adam_multi_tensor(smg_handle* h, float* const* params, const float* const* grads, float* const* m, float* const* v,
                      const int64_t* numel, int n, int step, float lr, float b1, float b2, float eps, cudaStream_t st) {
    if (n == 0) return SMG_OK;
    smg_handle::StepState& S = h->step;
    const size_t ptr_bytes = (size_t)n * sizeof(void*);
    std::vector<AdamChunk> chunks;
    for (int t = 0; t < n; ++t)
        for (int64_t o = 0; o < numel[t]; o += 4096) chunks.push_back(AdamChunk{t, (int)o, (int)(numel[t] - o < 4096 ? numel[t] - o : 4096)});
    const size_t need = 4 * ptr_bytes + chunks.size() * sizeof(AdamChunk) + 64;
    uint64_t sig = fnv(1469598103934665603ull, params, ptr_bytes);
    sig = fnv(sig, grads, ptr_bytes);
    sig = fnv(sig, m, ptr_bytes);
    sig = fnv(sig, v, ptr_bytes);
    sig = fnv(sig, numel, (size_t)n * sizeof(int64_t));
    if (need > S.adam_tables_bytes) {
        SMG_CUDA(cudaStreamSynchronize(st));
        if (S.adam_tables) cudaFree(S.adam_tables);
        S.adam_tables_bytes = need * 2;
        SMG_CUDA(cudaMalloc(&S.adam_tables, S.adam_tables_bytes));
        S.adam_sig = 0;
    }
    uint8_t* b = reinterpret_cast<uint8_t*>(S.adam_tables);
    float* dyn = reinterpret_cast<float*>(b + 4 * ptr_bytes + chunks.size() * sizeof(AdamChunk));
    if (sig != S.adam_sig) {
        SMG_CUDA(cudaStreamSynchronize(st));
        SMG_CUDA(cudaMemcpy(b, params, ptr_bytes, cudaMemcpyHostToDevice));
        SMG_CUDA(cudaMemcpy(b + ptr_bytes, grads, ptr_bytes, cudaMemcpyHostToDevice));
        SMG_CUDA(cudaMemcpy(b + 2 * ptr_bytes, m, ptr_bytes, cudaMemcpyHostToDevice));
        SMG_CUDA(cudaMemcpy(b + 3 * ptr_bytes, v, ptr_bytes, cudaMemcpyHostToDevice));
        SMG_CUDA(cudaMemcpy(b + 4 * ptr_bytes, chunks.data(), chunks.size() * sizeof(AdamChunk), cudaMemcpyHostToDevice));
        S.adam_sig = sig;
        S.adam_chunks = (int)chunks.size();
    }
    const float hd[3] = {0.f, 1.f - powf(b1, (float)step), sqrtf(1.f - powf(b2, (float)step))};
    SMG_CUDA(cudaMemcpyAsync(dyn, hd, sizeof(hd), cudaMemcpyHostToDevice, st));
    adam_multi_kernel<<<S.adam_chunks, 256, 0, st>>>(reinterpret_cast<float* const*>(b), reinterpret_cast<const float* const*>(b + ptr_bytes),
                                                     reinterpret_cast<float* const*>(b + 2 * ptr_bytes),
                                                     reinterpret_cast<float* const*>(b + 3 * ptr_bytes),
                                                     reinterpret_cast<const AdamChunk*>(b + 4 * ptr_bytes), dyn, lr, b1, b2, eps);
    h->launches++;
    SMG_CUDA(cudaGetLastError());
    return SMG_OK;
}

}  // namespace smg

using namespace smg;

extern "C" int smg_train_step(smg_handle* h, const smg_train_step_args* args, const double* dev_scene_hm, const double* dev_mask_hm,
                              float* const* dev_params, float* const* dev_grads, float* const* dev_exp_avg,
                              float* const* dev_exp_avg_sq, int n_tensors, float* dev_loss, float* dev_q, float* dev_bn_mean,
                              float* dev_bn_var, void* stream) {
    SMG_CHECK(h && args && dev_scene_hm && dev_mask_hm && dev_params && dev_grads && dev_exp_avg && dev_exp_avg_sq && dev_loss && dev_q,
              SMG_ERR_INVALID, "smg_train_step: NULL argument");
    const smg_train_step_args& a = *args;
    SMG_CHECK(n_tensors == SMG_TRUNK_NUM_PARAMS + SMG_HEAD_NUM_PARAMS, SMG_ERR_INVALID, "smg_train_step: expected %d tensors, got %d",
              SMG_TRUNK_NUM_PARAMS + SMG_HEAD_NUM_PARAMS, n_tensors);
    SMG_CHECK(a.trunk_id >= 0 && a.trunk_id < SMG_NUM_TRUNKS && a.head_id >= 0 && a.head_id < SMG_NUM_HEADS, SMG_ERR_INVALID,
              "smg_train_step: trunk %d / head %d", a.trunk_id, a.head_id);
    SMG_CHECK(h->max_samples >= 2 && 2 * a.hm_size <= h->H && a.stddev != 0.0 && a.num_rotations >= 1 && a.adam_step >= 1 &&
                  (a.loss_kind == 0 || a.loss_kind == 1),
              SMG_ERR_INVALID, "smg_train_step: bad argument (max_samples %d, hm_size %d, stddev %g, step %d)", h->max_samples,
              a.hm_size, a.stddev, a.adam_step);
    TrunkW& T = h->trunks[a.trunk_id];
    HeadW& Hd = h->heads[a.head_id];
    SMG_CHECK(T.set && Hd.set, SMG_ERR_STATE, "smg_train_step: weights of trunk %d / head %d not set", a.trunk_id, a.head_id);
    SMG_CHECK((T.packed & SMG_PACK_DGRAD) && (h->pack_mask & SMG_PACK_DGRAD), SMG_ERR_STATE,
              "smg_train_step: the data-gradient layout is not packed (smg_set_pack_layouts)");
    // the step updates the caller's parameter tensors in place and re-packs from them: they must be the tensors the packed
    // weights came from
    for (int i = 0; i < SMG_TRUNK_NUM_PARAMS; ++i)
        SMG_CHECK(T.src.size() == SMG_TRUNK_NUM_PARAMS && T.src[i] == dev_params[i], SMG_ERR_STATE,
                  "smg_train_step: parameter %d is not the tensor smg_set_trunk_weights packed", i);
    for (int i = 0; i < SMG_HEAD_NUM_PARAMS; ++i)
        SMG_CHECK(Hd.src.size() == SMG_HEAD_NUM_PARAMS && Hd.src[i] == dev_params[SMG_TRUNK_NUM_PARAMS + i], SMG_ERR_STATE,
                  "smg_train_step: head parameter %d is not the tensor smg_set_head_weights packed", i);
    if (a.loss_kind == 1) SMG_CHECK(Hd.n_out == 3 && a.label >= 0.f && a.label <= 2.f, SMG_ERR_INVALID, "smg_train_step: CE needs 3 logits");
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    SMG_TRY(ensure_train_workspace(h));
    smg_handle::StepState& S = h->step;
    const size_t ptr_bytes = (size_t)n_tensors * sizeof(void*);

    // ---- device tables of the caller's tensors (rebuilt when a pointer changes)
    uint64_t sig = 1469598103934665603ull;
    sig = fnv(sig, dev_params, ptr_bytes);
    sig = fnv(sig, dev_grads, ptr_bytes);
    sig = fnv(sig, dev_exp_avg, ptr_bytes);
    sig = fnv(sig, dev_exp_avg_sq, ptr_bytes);
    if (!S.tables) {
        // 4 pointer arrays + chunk list (<= 7.2 M parameters / 4096 + one partial chunk per tensor) + dyn + out + BN staging
        const size_t chunk_cap = 4096;
        S.tables_bytes = 4 * ptr_bytes + chunk_cap * sizeof(AdamChunk) + 256 + 256 + 2 * (size_t)2 * SMG_TRUNK_BN_CHANNELS * 4;
        SMG_CUDA(cudaMalloc(&S.tables, S.tables_bytes));
        uint8_t* b = reinterpret_cast<uint8_t*>(S.tables);
        S.d_params = reinterpret_cast<float**>(b);
        S.d_grads = reinterpret_cast<float**>(b + ptr_bytes);
        S.d_m = reinterpret_cast<float**>(b + 2 * ptr_bytes);
        S.d_v = reinterpret_cast<float**>(b + 3 * ptr_bytes);
        S.d_chunks = b + 4 * ptr_bytes;
        S.dyn = reinterpret_cast<float*>(b + 4 * ptr_bytes + chunk_cap * sizeof(AdamChunk));
        S.out = S.dyn + 64;
        S.bn_mean = S.out + 64;
        S.bn_var = S.bn_mean + (size_t)2 * SMG_TRUNK_BN_CHANNELS;
        h->workspace_bytes += (int64_t)S.tables_bytes;
    }
    if (sig != S.tables_sig || Hd.n_out != S.tables_n_out) {
        std::vector<int64_t> numel;
        trunk_param_numel(h, numel, Hd.n_out);
        SMG_CHECK((int)numel.size() == n_tensors, SMG_ERR_STATE, "smg_train_step: %zu tensor sizes", numel.size());
        std::vector<AdamChunk> chunks;
        for (int t = 0; t < n_tensors; ++t)
            for (int64_t o = 0; o < numel[t]; o += 4096)
                chunks.push_back(AdamChunk{t, (int)o, (int)(numel[t] - o < 4096 ? numel[t] - o : 4096)});
        SMG_CHECK(chunks.size() <= 4096, SMG_ERR_STATE, "smg_train_step: %zu Adam chunks", chunks.size());
        SMG_CUDA(cudaStreamSynchronize(st));   // a previous step may still read the tables
        SMG_CUDA(cudaMemcpy(S.d_params, dev_params, ptr_bytes, cudaMemcpyHostToDevice));
        SMG_CUDA(cudaMemcpy(S.d_grads, dev_grads, ptr_bytes, cudaMemcpyHostToDevice));
        SMG_CUDA(cudaMemcpy(S.d_m, dev_exp_avg, ptr_bytes, cudaMemcpyHostToDevice));
        SMG_CUDA(cudaMemcpy(S.d_v, dev_exp_avg_sq, ptr_bytes, cudaMemcpyHostToDevice));
        SMG_CUDA(cudaMemcpy(S.d_chunks, chunks.data(), chunks.size() * sizeof(AdamChunk), cudaMemcpyHostToDevice));
        S.n_chunks = (int)chunks.size();
        S.host_grads.assign(dev_grads, dev_grads + n_tensors);
        S.tables_sig = sig;
        S.tables_n_out = Hd.n_out;
    }

    // ---- per-call scalars and inputs at fixed addresses
    const float dyn[3] = {a.label, 1.f - powf(a.beta1, (float)a.adam_step), sqrtf(1.f - powf(a.beta2, (float)a.adam_step))};
    SMG_CUDA(cudaMemcpyAsync(S.dyn, dyn, sizeof(dyn), cudaMemcpyHostToDevice, st));   // pageable source: staged before the call returns
    const size_t hm_elems = (size_t)a.hm_size * a.hm_size;
    SMG_CUDA(cudaMemcpyAsync(h->hm_stage, dev_scene_hm, hm_elems * 8, cudaMemcpyDeviceToDevice, st));
    SMG_CUDA(cudaMemcpyAsync(h->hm_stage + hm_elems, dev_mask_hm, hm_elems * 8, cudaMemcpyDeviceToDevice, st));

    // ---- the step itself: eager on first sight of a configuration, captured on the second, replayed afterwards
    uint64_t gsig = fnv(sig, &a.trunk_id, sizeof(int));
    gsig = fnv(gsig, &a.head_id, sizeof(int));
    gsig = fnv(gsig, &a.rot_idx, sizeof(int));
    gsig = fnv(gsig, &a.num_rotations, sizeof(int));
    gsig = fnv(gsig, &a.hm_size, sizeof(int));
    gsig = fnv(gsig, &a.mean, sizeof(double));
    gsig = fnv(gsig, &a.stddev, sizeof(double));
    gsig = fnv(gsig, &a.loss_kind, sizeof(int));
    gsig = fnv(gsig, a.class_weight, sizeof(a.class_weight));
    gsig = fnv(gsig, &a.lr, 4 * sizeof(float));
    gsig = fnv(gsig, &a.flags, sizeof(int));
    gsig = fnv(gsig, &h->precision, sizeof(int));
    gsig = fnv(gsig, &h->pack_mask, sizeof(int));
    smg_handle::StepGraph* G = nullptr;
    for (auto& g : S.graphs)
        if (g.sig == gsig) { G = &g; break; }
    if (!G) {
        S.graphs.push_back(smg_handle::StepGraph{gsig, 0, 0, nullptr});
        G = &S.graphs.back();
    }
    const bool graphable = h->use_graphs && !h->profile && S.graphs.size() <= 128;
    if (!graphable || G->seen == 0) {
        G->seen = 1;
        SMG_TRY(train_step_body(h, a, st));
    } else {
        SMG_CUDA(cudaEventRecord(h->g_in, st));
        SMG_CUDA(cudaStreamWaitEvent(h->gstream, h->g_in, 0));
        if (!G->exec) {
            cudaGraph_t graph = nullptr;
            const int64_t before = h->launches;
            SMG_CUDA(cudaStreamBeginCapture(h->gstream, cudaStreamCaptureModeRelaxed));
            const int status = train_step_body(h, a, h->gstream);
            cudaError_t e = cudaStreamEndCapture(h->gstream, &graph);
            G->n_launches = h->launches - before;
            h->launches = before;   // capturing enqueues nothing
            if (status != SMG_OK || e != cudaSuccess) {
                if (graph) cudaGraphDestroy(graph);
                if (status == SMG_OK) set_error("smg_train_step: graph capture failed: %s", cudaGetErrorString(e));
                return status != SMG_OK ? status : SMG_ERR_CUDA;
            }
            e = cudaGraphInstantiate(&G->exec, graph, 0);
            cudaGraphDestroy(graph);
            if (e != cudaSuccess) {
                set_error("smg_train_step: cudaGraphInstantiate: %s", cudaGetErrorString(e));
                return SMG_ERR_CUDA;
            }
        }
        SMG_CUDA(cudaGraphLaunch(G->exec, h->gstream));
        h->launches += G->n_launches;
        SMG_CUDA(cudaEventRecord(h->g_out, h->gstream));
        SMG_CUDA(cudaStreamWaitEvent(st, h->g_out, 0));
    }
    h->train.valid = false;   // the saved activations were consumed by this step's own backward
    const int n_out = Hd.n_out;
    SMG_CUDA(cudaMemcpyAsync(dev_q, S.out, (size_t)n_out * 4, cudaMemcpyDeviceToDevice, st));
    SMG_CUDA(cudaMemcpyAsync(dev_loss, S.out + 4, 4, cudaMemcpyDeviceToDevice, st));
    if (dev_bn_mean && dev_bn_var) {
        SMG_CUDA(cudaMemcpyAsync(dev_bn_mean, S.bn_mean, (size_t)2 * SMG_TRUNK_BN_CHANNELS * 4, cudaMemcpyDeviceToDevice, st));
        SMG_CUDA(cudaMemcpyAsync(dev_bn_var, S.bn_var, (size_t)2 * SMG_TRUNK_BN_CHANNELS * 4, cudaMemcpyDeviceToDevice, st));
    }
    return SMG_OK;
}
