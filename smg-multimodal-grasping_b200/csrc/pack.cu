// pack.cu - re-pack torch OIHW fp32 convolution weights into the kernel layouts.
// Replaces what `model.cuda()` does for the reference (code/trainer.py:90-92): here the
// device copy of the weights is also a layout change.
//   w_ffma : [tap][cin][cout] fp32                                   (conv_ffma.cu)
//   w_tf32 / w_bf16 : stage images in the UMMA no-swizzle K-major layout (conv_umma.cu):
//       1x1 : [ntile][kgroup(32 ch)][chunk(16 B)][n (BN rows)][elements of the chunk]
//       3x3 : [kgroup][tap][chunk][n (32 rows)][elements of the chunk]
//   w_tf32_t : weights as the tensor-memory A operand of conv1_t.cu / conv3_wt.cu;  w_dgrad_tf32 : data-gradient stage image
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

// data-gradient weights as a tf32 stage image for conv_umma.cu (see ConvW::w_dgrad_tf32): element i of the image
__device__ __forceinline__ float dgrad_image_value(const float* __restrict__ w, int i, int cout, int cin, int taps, int k_offset,
                                                   int k_total) {
    int r = i;
    const int e = r % 4; r /= 4;
    const int n = r % 128; r /= 128;
    const int c = r % 8; r /= 8;
    int ntile, kg, tap;
    if (taps == 1) { kg = r % (cout / 32); ntile = r / (cout / 32); tap = 0; }
    else { tap = r % 9; kg = r / 9; ntile = 0; }
    const int oc = ntile * 128 + n;          // output channel of the data gradient = input channel of the convolution
    const int ic = kg * 32 + c * 4 + e;      // input channel of the data gradient = output channel of the convolution
    return oc < cin ? w[((size_t)ic * k_total + k_offset + oc) * taps + (taps - 1 - tap)] : 0.f;
}

__global__ void pack_dgrad_umma_kernel(const float* __restrict__ w, float* __restrict__ out, int cout, int cin, int taps,
                                       int k_offset, int k_total) {
    const int total = taps * cout * dgrad_cin_padded(cin, taps);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x)
        out[i] = dgrad_image_value(w, i, cout, cin, taps, k_offset, k_total);
}

// 3x3, tf32: [column block][dx*cout + co][16], column = ((g*3 + dy)*4 + k)*8 + e, for the weights-in-tensor-memory kernel
__global__ void pack_umma_t_kernel(const float* __restrict__ w, float* __restrict__ out, int cout, int cin, int k_offset,
                                   int k_total) {
    const int total = 9 * cin * cout;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int rows = 3 * cout;
        const int c_in = i % 16, row = (i / 16) % rows, c16 = i / (16 * rows);
        int r = c16 * 16 + c_in;
        const int e = r % 8; r /= 8;
        const int k = r % 4; r /= 4;
        const int dy = r % 3; r /= 3;
        const int g = r;
        const int dx = row / cout, co = row - dx * cout, ci = g * 32 + k * 8 + e;
        out[i] = w[((size_t)co * k_total + k_offset + ci) * 9 + dy * 3 + dx];
    }
}

// 1x1, cout = 128, tf32: [cin/16][128 rows][16] for conv1_t.cu
__global__ void pack_umma_t1_kernel(const float* __restrict__ w, float* __restrict__ out, int cin, int k_offset, int k_total) {
    const int total = cin * 128;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int c_in = i % 16, row = (i / 16) % 128, c16 = i / (16 * 128);
        out[i] = w[(size_t)row * k_total + k_offset + c16 * 16 + c_in];
    }
}

// w_tf32 stage images -> [hi image][lo image] per stage (3xTF32 operands of the fp32 mode)
__global__ void pack_split_kernel(const float* __restrict__ tf32, float* __restrict__ out, int total, int stage_elems) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int stage = i / stage_elems, within = i - stage * stage_elems;
        const float v = tf32[i];
        const float hi = __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u);   // nearest tf32
        out[(size_t)(2 * stage) * stage_elems + within] = hi;
        out[(size_t)(2 * stage + 1) * stage_elems + within] = v - hi;
    }
}

// 3x3 split weights for the dx-merged tensor-core form (conv_umma.cu, TAPS == 3): stage (kgroup, dy) = [hi | lo] images of
// [chunk (8)][n = dx*cout + co][4 floats]; element i of the hi/lo pair
__device__ __forceinline__ void split3_store(const float* __restrict__ w, float* __restrict__ out, int i, int cout, int k_off,
                                             int k_total) {
    int r = i;
    const int e = r % 4; r /= 4;
    const int n = r % (3 * cout); r /= 3 * cout;
    const int c = r % 8; r /= 8;
    const int dy = r % 3, kg = r / 3;
    const int dx = n / cout, co = n - dx * cout, ci = kg * 32 + c * 4 + e;
    const float v = w[((size_t)co * k_total + k_off + ci) * 9 + dy * 3 + dx];
    const int stage_elems = 32 * 3 * cout;
    const int stage = i / stage_elems, within = i - stage * stage_elems;
    const float hi = __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u);   // nearest tf32
    out[(size_t)(2 * stage) * stage_elems + within] = hi;
    out[(size_t)(2 * stage + 1) * stage_elems + within] = v - hi;
}

__global__ void pack_split3_kernel(const float* __restrict__ w, float* __restrict__ out, int cout, int cin, int k_off, int k_total) {
    const int total = 9 * cin * cout;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x)
        split3_store(w, out, i, cout, k_off, k_total);
}

// ---- batched variant: one launch packs every convolution of a trunk (blockIdx.y = job) ----
__global__ void pack_batch_kernel(const PackJob* __restrict__ jobs, int mask) {
    const PackJob j = jobs[blockIdx.y];
    const int total = j.taps * j.cin * j.cout;
    const int kgs = j.cin / 32;
    if ((mask & SMG_PACK_DGRAD) && j.dgrad_tf32 != nullptr) {
        const int total_pad = j.taps * j.cout * dgrad_cin_padded(j.cin, j.taps);
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total_pad; i += gridDim.x * blockDim.x)
            j.dgrad_tf32[i] = dgrad_image_value(j.src, i, j.cout, j.cin, j.taps, j.k_off, j.k_total);
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        if (mask & SMG_PACK_FFMA) {
            const int co = i % j.cout, ci = (i / j.cout) % j.cin, t = i / (j.cout * j.cin);
            j.ffma[i] = j.src[((size_t)co * j.k_total + j.k_off + ci) * j.taps + t];
        }
        if (mask & SMG_PACK_DGRAD) {
            const int ci = i % j.cin, co = (i / j.cin) % j.cout, t = i / (j.cout * j.cin);
            j.dgrad[i] = j.src[((size_t)co * j.k_total + j.k_off + ci) * j.taps + (j.taps - 1 - t)];
        }
        if ((mask & SMG_PACK_TF32) && j.taps == 9) {
            if (j.tf32_t != nullptr) {
                // [column block c16][row = dx*cout + co][16]: column ((g*3 + dy)*4 + k)*8 + e, cin = g*32 + k*8 + e (conv3_wt.cu:
                // weights in tensor memory; a warp reads 2 KB contiguous per column block)
                const int rows = 3 * j.cout;
                const int c_in = i % 16, row = (i / 16) % rows, c16 = i / (16 * rows);
                int r = c16 * 16 + c_in;
                const int e = r % 8; r /= 8;
                const int k = r % 4; r /= 4;
                const int dy = r % 3; r /= 3;
                const int g = r;
                const int dx = row / j.cout, co = row - dx * j.cout, ci = g * 32 + k * 8 + e;
                j.tf32_t[i] = j.src[((size_t)co * j.k_total + j.k_off + ci) * 9 + dy * 3 + dx];
            }
        }
        if ((mask & SMG_PACK_TF32) && j.taps == 1 && j.cout == 128 && j.tf32_t != nullptr) {
            // [cin/16][128 rows][16]: the [cout][cin] matrix in 16-column blocks (conv1_t.cu: weights as the A operand)
            const int c_in = i % 16, row = (i / 16) % 128, c16 = i / (16 * 128);
            j.tf32_t[i] = j.src[(size_t)row * j.k_total + j.k_off + c16 * 16 + c_in];
        }
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
            if (!(mask & (pass == 0 ? (SMG_PACK_TF32 | SMG_PACK_FFMA) : SMG_PACK_BF16))) continue;
            const int EPC = pass == 0 ? 4 : 8, CH = 32 / EPC;
            int r = i;
            const int e = r % EPC; r /= EPC;
            const int n = r % j.bn; r /= j.bn;
            const int c = r % CH; r /= CH;
            int ntile, kg, tap;
            if (j.taps == 1) { kg = r % kgs; r /= kgs; ntile = r; tap = 0; }
            else { tap = r % j.taps; r /= j.taps; kg = r; ntile = 0; }
            const int co = ntile * j.bn + n, ci = kg * 32 + c * EPC + e;
            const float v = j.src[((size_t)co * j.k_total + j.k_off + ci) * j.taps + tap];
            if (pass == 0) j.tf32[i] = v;
            else j.bf16[i] = __float2bfloat16_rn(v);
            if (pass == 0 && (mask & SMG_PACK_FFMA) && j.split != nullptr && j.taps == 9) {
                split3_store(j.src, j.split, i, j.cout, j.k_off, j.k_total);
            } else if (pass == 0 && (mask & SMG_PACK_FFMA) && j.split != nullptr) {
                // the same stage sequence with each stage (32 channels x bn rows) stored as [hi image][lo image]
                const int stage_elems = 32 * j.bn;
                const int stage = i / stage_elems, within = i - stage * stage_elems;
                const float hi = __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u);   // nearest tf32
                j.split[(size_t)(2 * stage) * stage_elems + within] = hi;
                j.split[(size_t)(2 * stage + 1) * stage_elems + within] = v - hi;
            }
        }
    }
}

__global__ void copy_batch_kernel(const CopyJob* __restrict__ jobs) {
    const CopyJob j = jobs[blockIdx.x];
    for (int i = threadIdx.x; i < j.n; i += blockDim.x) j.dst[i] = j.src[i];
}

PackJob make_pack_job(const float* w_oihw, const ConvW& cw, int k_offset, int k_total) {
    PackJob j;
    j.src = w_oihw; j.ffma = cw.w_ffma; j.tf32 = reinterpret_cast<float*>(cw.w_tf32);
    j.tf32_t = reinterpret_cast<float*>(cw.w_tf32_t);
    j.dgrad_tf32 = reinterpret_cast<float*>(cw.w_dgrad_tf32);
    j.split = reinterpret_cast<float*>(cw.w_split);
    j.bf16 = reinterpret_cast<__nv_bfloat16*>(cw.w_bf16); j.dgrad = cw.w_dgrad;
    j.cin = cw.cin; j.cout = cw.cout; j.taps = cw.taps; j.k_off = k_offset; j.k_total = k_total;
    j.bn = cw.taps == 9 ? 32 : (cw.cout < 128 ? cw.cout : 128);
    return j;
}

// packs every convolution / copies every BatchNorm affine of a trunk from job tables that already live on the device
int launch_pack_tables(smg_handle* h, const PackJob* dev_pj, int n_pj, const CopyJob* dev_cj, int n_cj, cudaStream_t st) {
    if (n_pj > 0) {
        pack_batch_kernel<<<dim3(48, (unsigned)n_pj), 256, 0, st>>>(dev_pj, h->pack_mask);
        h->launches++;
    }
    if (n_cj > 0) {
        copy_batch_kernel<<<(unsigned)n_cj, 256, 0, st>>>(dev_cj);
        h->launches++;
    }
    SMG_CUDA(cudaGetLastError());
    return SMG_OK;
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
    if (cw.taps == 1 && cw.cout == 128 && cw.w_tf32_t != nullptr) {
        pack_umma_t1_kernel<<<blocks, threads, 0, st>>>(w_oihw, reinterpret_cast<float*>(cw.w_tf32_t), cw.cin, k_offset, k_total);
        h->launches++;
    }
    if (cw.taps == 9 && cw.w_tf32_t != nullptr) {
        pack_umma_t_kernel<<<blocks, threads, 0, st>>>(w_oihw, reinterpret_cast<float*>(cw.w_tf32_t), cw.cout, cw.cin, k_offset,
                                                       k_total);
        h->launches++;
    }
    if (cw.w_split != nullptr) {
        if (cw.taps == 9)
            pack_split3_kernel<<<blocks, threads, 0, st>>>(w_oihw, reinterpret_cast<float*>(cw.w_split), cw.cout, cw.cin, k_offset, k_total);
        else
            pack_split_kernel<<<blocks, threads, 0, st>>>(reinterpret_cast<const float*>(cw.w_tf32), reinterpret_cast<float*>(cw.w_split),
                                                          total, 32 * bn);
        h->launches++;
    }
    if (cw.w_dgrad_tf32 != nullptr) {
        pack_dgrad_umma_kernel<<<blocks, threads, 0, st>>>(w_oihw, reinterpret_cast<float*>(cw.w_dgrad_tf32), cw.cout, cw.cin,
                                                           cw.taps, k_offset, k_total);
        h->launches++;
    }
    SMG_CUDA(cudaGetLastError());
    return SMG_OK;
}

}  // namespace smg
