// pack.cu - re-pack torch OIHW fp32 convolution weights into the kernel layouts.
// Replaces what `model.cuda()` does for the reference (code/trainer.py:90-92): here the
// device copy of the weights is also a layout change.
//   w_ffma : [tap][cin][cout] fp32                                   (conv_ffma.cu)
//   w_tf32 / w_bf16 : stage images in the UMMA no-swizzle K-major layout (conv_umma.cu):
//       1x1 : [ntile][kgroup(32 ch)][chunk(16 B)][n (BN rows)][elements of the chunk]
//       3x3 : [kgroup][tap][chunk][n (32 rows)][elements of the chunk]
#include "smg_internal.cuh"

namespace smg {

__global__ void pack_ffma_kernel(const float* __restrict__ w, float* __restrict__ out, int cout, int cin, int taps,
                                 int k_offset, int k_total) {
    const int total = taps * cin * cout;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int co = i % cout;
        const int ci = (i / cout) % cin;
        const int t = i / (cout * cin);
        out[i] = w[((size_t)co * k_total + k_offset + ci) * taps + t];
    }
}

// data-gradient weights: out[tap'][co][ci] = w[co][ci][taps-1-tap']  (3x3: both kernel axes flipped)
__global__ void pack_dgrad_kernel(const float* __restrict__ w, float* __restrict__ out, int cout, int cin, int taps,
                                  int k_offset, int k_total) {
    const int total = taps * cin * cout;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int ci = i % cin;
        const int co = (i / cin) % cout;
        const int t = i / (cout * cin);
        out[i] = w[((size_t)co * k_total + k_offset + ci) * taps + (taps - 1 - t)];
    }
}

template <typename T>
__device__ __forceinline__ T cvt(float v);
template <>
__device__ __forceinline__ float cvt<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 cvt<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

template <typename T>
__global__ void pack_umma_kernel(const float* __restrict__ w, T* __restrict__ out, int cout, int cin, int taps,
                                 int k_offset, int k_total, int bn) {
    constexpr int EPC = 16 / sizeof(T);
    constexpr int CH = 32 / EPC;
    const int kgs = cin / 32;
    const size_t total = (size_t)taps * cin * cout;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        size_t r = i;
        const int e = (int)(r % EPC); r /= EPC;
        const int n = (int)(r % bn); r /= bn;
        const int c = (int)(r % CH); r /= CH;
        int ntile, kg, tap;
        if (taps == 1) {
            kg = (int)(r % kgs); r /= kgs;
            ntile = (int)r;
            tap = 0;
        } else {
            tap = (int)(r % taps); r /= taps;
            kg = (int)r;
            ntile = 0;
        }
        const int co = ntile * bn + n;
        const int ci = kg * 32 + c * EPC + e;
        out[i] = cvt<T>(w[((size_t)co * k_total + k_offset + ci) * taps + tap]);
    }
}

size_t conv_packed_bytes_ffma(int cin, int cout, int taps) { return (size_t)cin * cout * taps * sizeof(float); }
size_t conv_packed_bytes_umma(int cin, int cout, int taps, int elt_bytes) {
    return (size_t)cin * cout * taps * elt_bytes;
}

int pack_conv_weights(smg_handle* h, const float* w_oihw, ConvW& cw, int k_offset, int k_total, cudaStream_t st) {
    const int total = cw.cin * cw.cout * cw.taps;
    const int threads = 256;
    int blocks = (total + threads - 1) / threads;
    if (blocks > 1024) blocks = 1024;
    const int bn = cw.taps == 9 ? 32 : (cw.cout < 128 ? cw.cout : 128);
    pack_ffma_kernel<<<blocks, threads, 0, st>>>(w_oihw, cw.w_ffma, cw.cout, cw.cin, cw.taps, k_offset, k_total);
    pack_umma_kernel<float><<<blocks, threads, 0, st>>>(w_oihw, reinterpret_cast<float*>(cw.w_tf32), cw.cout, cw.cin,
                                                        cw.taps, k_offset, k_total, bn);
    pack_umma_kernel<__nv_bfloat16><<<blocks, threads, 0, st>>>(w_oihw, reinterpret_cast<__nv_bfloat16*>(cw.w_bf16),
                                                                cw.cout, cw.cin, cw.taps, k_offset, k_total, bn);
    pack_dgrad_kernel<<<blocks, threads, 0, st>>>(w_oihw, cw.w_dgrad, cw.cout, cw.cin, cw.taps, k_offset, k_total);
    h->launches += 4;
    SMG_CUDA(cudaGetLastError());
    return SMG_OK;
}

}  // namespace smg
