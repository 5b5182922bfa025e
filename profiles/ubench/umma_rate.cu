// umma_rate.cu - microbenchmark: cycles per tcgen05.mma (cta_group::1, M=128, operands in shared memory) as a function of
// kind (tf32 / bf16), N and the A-operand layout (no-swizzle vs 128-byte swizzle).  Build: see profiles/ubench/README.
// One elected thread issues R MMAs back to back on the same operands, commits, waits; reported = (t1 - t0) / R.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}

template <int TF32>
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int swz, int R, int nacc, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tptr;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tptr)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tptr;
    if (threadIdx.x == 0) {
        const uint32_t fmt = TF32 ? 2u : 1u;
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t a0 = smem_u32(smem), b0 = a0 + 16384;
        const uint64_t ad = swz ? desc(a0, 16, 1024, 2) : desc(a0, 129 * 16, 128, 0);
        const uint64_t bd = desc(b0, N * 16, 128, 0);
        const long long t0 = clock64();
        const uint32_t d0 = tm, d1 = tm + (uint32_t)((nacc - 1) * N);
        for (int i = 0; i < R; i += 8) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const uint32_t d = (u & 1) ? d1 : d0;
                if (TF32)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(1) : "memory");
                else
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(1) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        mbar_wait(&bar, 0);
        const long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
}

int main() {
    long long* d;
    cudaMalloc(&d, 8);
    cudaFuncSetAttribute(rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
    cudaFuncSetAttribute(rate_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
    const int R = 2048;
    printf("kind,N,a_layout,accumulators,grid,cycles_per_mma\n");
    for (int tf = 1; tf >= 0; --tf)
        for (int N : {32, 64, 128, 256})
            for (int swz = 0; swz < 2; ++swz)
                for (int nacc : {1, 2})
                    for (int grid : {1, 148}) {
                        if (nacc * N > 512) continue;
                        for (int rep = 0; rep < 2; ++rep) {
                            if (tf) rate_kernel<1><<<grid, 128, 48 * 1024>>>(N, swz, R, nacc, d);
                            else rate_kernel<0><<<grid, 128, 48 * 1024>>>(N, swz, R, nacc, d);
                        }
                        long long c = 0;
                        cudaError_t e = cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
                        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                        printf("%s,%d,%s,%d,%d,%.1f\n", tf ? "tf32" : "bf16", N, swz ? "sw128" : "noswz", nacc, grid, (double)c / R);
                    }
    return 0;
}
