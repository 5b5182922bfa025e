// nms.cu - K12: NMS.py_cpu_nms (/root/reference/code/NMS.py:8-59).
// Greedy IoU suppression in INDEX order (the reference does not sort; it relies on the
// detector's score-sorted output), +1 pixel area convention, strict area pre-filter.
// One CTA; boxes and the alive flags live in shared memory; every float32 operation is
// the same single rounding numpy performs, so the kept list is identical.
#include "smg_internal.cuh"

namespace smg {

constexpr int NMS_MAX = 1024;

__global__ void __launch_bounds__(256)
nms_kernel(const float* __restrict__ boxes, int n, float thr, float amin, float amax, int32_t* __restrict__ keep,
           int32_t* __restrict__ n_keep) {
    __shared__ float x1[NMS_MAX], y1[NMS_MAX], x2[NMS_MAX], y2[NMS_MAX], ar[NMS_MAX];
    __shared__ unsigned char alive[NMS_MAX];
    __shared__ int s_count;
    const int tid = threadIdx.x;
    for (int i = tid; i < n; i += blockDim.x) {
        const float a = boxes[i * 4 + 0], b = boxes[i * 4 + 1], c = boxes[i * 4 + 2], d = boxes[i * 4 + 3];
        x1[i] = a; y1[i] = b; x2[i] = c; y2[i] = d;
        const float area = __fmul_rn(__fsub_rn(c, a), __fsub_rn(d, b));                    // NMS.py:19
        alive[i] = (area > amin && area < amax) ? 1 : 0;                                   // NMS.py:20
        ar[i] = __fmul_rn(__fadd_rn(__fsub_rn(c, a), 1.0f), __fadd_rn(__fsub_rn(d, b), 1.0f));  // NMS.py:23
    }
    if (tid == 0) s_count = 0;
    __syncthreads();
    for (int i = 0; i < n; ++i) {
        if (alive[i]) {  // uniform across the CTA (read after the barrier)
            const float ax1 = x1[i], ay1 = y1[i], ax2 = x2[i], ay2 = y2[i], aa = ar[i];
            for (int j = i + 1 + tid; j < n; j += blockDim.x) {
                if (!alive[j]) continue;
                const float xx1 = fmaxf(ax1, x1[j]), yy1 = fmaxf(ay1, y1[j]);
                const float xx2 = fminf(ax2, x2[j]), yy2 = fminf(ay2, y2[j]);
                const float w = fmaxf(0.0f, __fadd_rn(__fsub_rn(xx2, xx1), 1.0f));
                const float hgt = fmaxf(0.0f, __fadd_rn(__fsub_rn(yy2, yy1), 1.0f));
                const float inter = __fmul_rn(w, hgt);
                const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(aa, ar[j]), inter));  // NMS.py:38
                if (!(ovr <= thr)) alive[j] = 0;                                            // NMS.py:39
            }
            if (tid == 0) keep[s_count++] = i;
        }
        __syncthreads();
    }
    if (tid == 0) n_keep[0] = s_count;
}

int launch_nms(smg_handle* h, const float* boxes, int n, float thr, float amin, float amax, int32_t* keep,
               int32_t* n_keep, cudaStream_t st) {
    SMG_CHECK(n >= 0 && n <= NMS_MAX, SMG_ERR_INVALID, "nms: n=%d outside [0,%d]", n, NMS_MAX);
    nms_kernel<<<1, 256, 0, st>>>(boxes, n, thr, amin, amax, keep, n_keep);
    h->launches++;
    SMG_CUDA(cudaGetLastError());
    return SMG_OK;
}

}  // namespace smg
