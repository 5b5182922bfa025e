// umma_common.cuh - PTX wrappers (mbarrier, bulk TMA, cp.async, tcgen05 alloc/mma/commit/ld), the UMMA shared-memory
// descriptor and the operand-layout constants shared by the tensor-core convolution kernels (conv_umma*.cu).
#pragma once
#include "smg_internal.cuh"

namespace smg {

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    }
}
// waiting with back-off: a failed probe is followed by a short sleep, so that warps parked on a barrier for microseconds
// (epilogue warps during the K loop, transform warps waiting for TMA data) do not spend shared-memory pipe cycles polling.
// ncu on conv_umma_tma_kernel: 4.7 M of 16.7 M LSU shared-memory wavefronts were barrier probes.
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity, uint32_t ns) {
    const uint32_t addr = smem_u32(bar);
    for (;;) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) break;
        __nanosleep(ns);
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// Ampere-style async copies (LDGSTS): raw activations land in shared memory without holding registers
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// Programmatic dependent launch: a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start while
// its predecessor in the stream is still draining; pdl_launch_dependents() lets OUR successor do the same, pdl_wait() blocks
// until the predecessor has completed and its writes are visible.  Both are no-ops for a normal launch.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
template <int ELT>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    if (ELT == 4) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
            : "memory");
    }
}
// 32 lanes x 32 columns of 32-bit accumulators -> 32 registers per thread (thread = TMEM lane)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// shared-memory matrix descriptor, no swizzle, K-major: start>>4 | LBO>>4 <<16 | SBO>>4 <<32 | version 1 <<46
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// ------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------
struct UmmaDev {
    const float* in;
    int in_cstride, cin, hin;
    int prologue_mode;
    const double* in_stats;
    int in_stats_stride;
    const float* gamma;
    const float* beta;
    const float* scale;
    const float* shift;
    int relu;
    const uint8_t* w;  // packed stage images
    float* out;
    int out_cstride, out_coff, cout;
    double* out_stats;
    int out_stats_stride;
    int hout;
    // 3x3 patch tiling
    int wp, ht, tiles_x;
    // 1: cp.async in-place producer (wins when the grid is small and the K loop is a latency chain);
    // 0: register producer (wins when the launch is bandwidth-bound: one shared-memory pass instead of three)
    int async_producer;
    // multi-tile kernel (conv_umma_mt.cu): tiles of one sample, tiles walked by one CTA
    int tiles_per_sample, tiles_per_cta;
    // persistent kernels: how many TMA boxes ahead of its shared-memory ring the loader prefetches into L2 (0 = off)
    int l2_prefetch = 0;
    // data-gradient convolutions: the output IS the gradient w.r.t. relu(bn(x)) of the BatchNorm that produced this
    // convolution's forward input, so the epilogue can do that BatchNorm's backward REDUCTION on the tile it holds:
    // S1 += dz, S2 += dz * xhat with dz = out masked by relu'(bn(x)) -> bnr_sums [S][cout] double2 (nullptr = off)
    const float* bnr_x = nullptr;      // raw BatchNorm input, NHWC on the output's pixel grid, channels [0, cout)
    int bnr_x_cstride = 0;
    const double* bnr_stats = nullptr; // its (sum, sumsq): [S][bnr_stats_stride] double2
    int bnr_stats_stride = 0;
    const float* bnr_gamma = nullptr;
    const float* bnr_beta = nullptr;
    double* bnr_sums = nullptr;
};

constexpr int UM = 128;         // rows per tile (UMMA M)
constexpr int KC = 32;          // channels per K group
constexpr int NA = 3;           // A stages (1x1)
constexpr int NB1 = 3;          // B stages (1x1)
constexpr int NB9 = 6;          // B stages (3x3)
constexpr int MAX_WP = 42;
constexpr int PATCH_ROWS = UM + 2 * MAX_WP + 2;  // 214

template <int ELT> struct EltCfg;
template <> struct EltCfg<4> {
    static constexpr int CH = 8;          // 16-byte chunks per 32-channel group
    static constexpr int EPC = 4;         // elements per chunk
    static constexpr int A_ROWS = 129;    // padded rows: (rows mod 8) == 1 -> conflict-free 16 B stores
    static constexpr int P_ROWS = 215;
    static constexpr uint32_t FMT = 2;    // TF32
};
template <> struct EltCfg<2> {
    static constexpr int CH = 4;
    static constexpr int EPC = 8;
    static constexpr int A_ROWS = 130;    // (rows mod 8) == 2
    static constexpr int P_ROWS = 218;
    static constexpr uint32_t FMT = 1;    // BF16
};

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
}

void umma_patch_geometry(int hw, int* wp, int* ht);

}  // namespace smg
