// tma_rate.cu - microbenchmark: what rate can ONE loader thread per SM pull activations into shared memory with
// cp.async.bulk.tensor boxes of the shapes the convolution kernels use?  No consumer: a stage is re-issued as soon as its
// transaction barrier completes, NS stages in flight per CTA, one CTA per SM (grid = 148).
//   pattern 0: 1x1 conv operand   - tensor [17][25600][256] fp32, box 32 ch x 128 px (128 rows of 128 B, 1 KB apart), 6 K groups
//   pattern 1: 3x3 conv patch     - tensor [17][160][160][128] fp32, box 32 ch x 42 x 5 (210 rows of 128 B, 512 B apart), 4 groups
//   pattern 2: contiguous bulk    - cp.async.bulk of 16 KB linear chunks (upper bound of the copy engine)
//   pattern 3/4: stem patches     - tensor [72][640][640], box 40 (64) x 21 floats, no swizzle
//   pattern 5: pattern 0 + 96 threads per CTA writing the layer's 128-channel fp32 output (read : write = 768 : 512 B per pixel)
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
}

template <int PATTERN>
__global__ void __launch_bounds__(128, 1) tma_kernel(const __grid_constant__ CUtensorMap tm, const float* base, int ns, int total_boxes,
                                                      int slot_bytes) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bars[16];
    if (threadIdx.x == 0) {
        for (int i = 0; i < 16; ++i) mbar_init(&bars[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (PATTERN == 5 && threadIdx.x >= 32) {
        // writers: 96 threads stream the 1x1 layer's output (128 channels fp32 per pixel) of the tiles this CTA "processes"
        const int b0 = (int)(((long long)blockIdx.x * total_boxes) / gridDim.x);
        const int b1 = (int)(((long long)(blockIdx.x + 1) * total_boxes) / gridDim.x);
        float4* outp = reinterpret_cast<float4*>(const_cast<float*>(base)) + (size_t)17 * 25600 * 256 / 4;   // second half of the buffer
        for (int t = b0 / 6; t < b1 / 6; ++t)
            for (int i = threadIdx.x - 32; i < 128 * 32; i += 96) outp[(size_t)t * 128 * 32 + i] = make_float4(1.f, 2.f, 3.f, 4.f);
    }
    if (threadIdx.x == 0) {
        const int b0 = (int)(((long long)blockIdx.x * total_boxes) / gridDim.x);
        const int b1 = (int)(((long long)(blockIdx.x + 1) * total_boxes) / gridDim.x);
        for (int i = b0; i < b1; ++i) {
            const int n = i - b0, slot = n % ns;
            if (n >= ns) mbar_wait(&bars[slot], ((n / ns) - 1) & 1);
            uint8_t* dst = smem + (size_t)slot * slot_bytes;
            if (PATTERN == 0 || PATTERN == 5) {
                const int kg = i % 6, tile = i / 6, s = tile / 200, mt = tile % 200;
                mbar_expect(&bars[slot], 16384);
                asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::
                             "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(kg * 32), "r"(mt * 128), "r"(s), "r"(smem_u32(&bars[slot])) : "memory");
            } else if (PATTERN == 1) {
                const int g = i % 4, tile = i / 4, s = tile / 216, r = tile % 216, ty = r / 4, tx = r % 4;
                mbar_expect(&bars[slot], 210 * 128);
                asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::
                             "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(g * 32), "r"(tx * 40 - 1), "r"(ty * 3 - 1), "r"(s), "r"(smem_u32(&bars[slot])) : "memory");
            } else if (PATTERN == 3 || PATTERN == 4) {
                const int bw = PATTERN == 3 ? 40 : 64;
                const int tile = i, s = tile / 800, r = tile % 800, ty = r / 20, tx = r % 20;
                mbar_expect(&bars[slot], bw * 21 * 4);
                asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::
                             "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(tx * 32 - 4), "r"(ty * 16 - 3), "r"(s), "r"(smem_u32(&bars[slot])) : "memory");
            } else {
                mbar_expect(&bars[slot], 16384);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                             "r"(smem_u32(dst)), "l"(reinterpret_cast<const uint8_t*>(base) + (size_t)i * 16384), "r"(16384), "r"(smem_u32(&bars[slot])) : "memory");
            }
        }
        const int n_total = b1 - b0;
        for (int k = (n_total > ns ? n_total - ns : 0); k < n_total; ++k) mbar_wait(&bars[k % ns], (k / ns) & 1);
    }
    __syncthreads();
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)fp;
    const size_t bytes = (size_t)17 * 25600 * 256 * 4;
    float* d;
    cudaMalloc(&d, bytes + (size_t)17 * 25600 * 128 * 4);
    cudaMemset(d, 0, bytes);
    CUtensorMap tm3, tm4;
    const cuuint32_t es[4] = {1, 1, 1, 1};
    {
        const cuuint64_t dims[3] = {256, 25600, 17};
        const cuuint64_t str[2] = {1024, 25600ull * 1024};
        const cuuint32_t box[3] = {32, 128, 1};
        enc(&tm3, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    {
        const cuuint64_t dims[4] = {128, 160, 160, 17};
        const cuuint64_t str[3] = {512, 160 * 512, 160ull * 160 * 512};
        const cuuint32_t box[4] = {32, 42, 5, 1};
        enc(&tm4, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    CUtensorMap tmS40, tmS64;
    for (int v = 0; v < 2; ++v) {
        const cuuint64_t dims[3] = {640, 640, 72};
        const cuuint64_t str[2] = {2560, 640ull * 2560};
        const cuuint32_t box[3] = {v == 0 ? 40u : 64u, 21, 1};
        CUresult r = enc(v == 0 ? &tmS40 : &tmS64, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode stem box %d: %d\n", v, (int)r);
    }
    cudaFuncSetAttribute(tma_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaFuncSetAttribute(tma_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaFuncSetAttribute(tma_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaFuncSetAttribute(tma_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaFuncSetAttribute(tma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaFuncSetAttribute(tma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    printf("pattern,stages_in_flight,box_bytes,total_MB,us,GB_per_s,GB_per_s_per_SM\n");
    for (int pat = 0; pat < 6; ++pat)
        for (int ns : {2, 3, 4, 6, 8, 12}) {
            if ((pat == 3 || pat == 4) && ns != 4) continue;
            const int slot = pat == 1 ? 27 * 1024 : ((pat == 3 || pat == 4) ? 5376 : 16 * 1024);
            if ((size_t)ns * slot > 220 * 1024) continue;
            const int total = (pat == 0 || pat == 5) ? 17 * 200 * 6 : (pat == 1 ? 17 * 216 * 4 : ((pat == 3 || pat == 4) ? 800 * 72 : (int)(bytes / 16384 / 4)));
            const double box_bytes = pat == 5 ? 16384 + 65536.0 / 6 : pat == 1 ? 210 * 128 : (pat == 3 ? 40 * 21 * 4 : (pat == 4 ? 64 * 21 * 4 : 16384));
            float ms = 0;
            for (int rep = 0; rep < 3; ++rep) {
                cudaEventRecord(e0);
                if (pat == 0) tma_kernel<0><<<148, 128, ns * slot>>>(tm3, d, ns, total, slot);
                else if (pat == 1) tma_kernel<1><<<148, 128, ns * slot>>>(tm4, d, ns, total, slot);
                else if (pat == 5) tma_kernel<5><<<148, 128, ns * slot>>>(tm3, d, ns, total, slot);
                else if (pat == 3) tma_kernel<3><<<148, 128, ns * slot>>>(tmS40, d, ns, total, slot);
                else if (pat == 4) tma_kernel<4><<<148, 128, ns * slot>>>(tmS64, d, ns, total, slot);
                else tma_kernel<2><<<148, 128, ns * slot>>>(tm3, d, ns, total, slot);
                cudaEventRecord(e1);
                cudaError_t e = cudaEventSynchronize(e1);
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                cudaEventElapsedTime(&ms, e0, e1);
            }
            const double mb = total * box_bytes / 1e6;
            printf("%d,%d,%.0f,%.1f,%.1f,%.0f,%.1f\n", pat, ns, box_bytes, mb, ms * 1e3, mb / ms, mb / ms / 148);
        }
    return 0;
}
