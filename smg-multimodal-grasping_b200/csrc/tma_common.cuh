// tma_common.cuh - tensor-map TMA helpers shared by the kernels whose activation operand is fetched by
// cp.async.bulk.tensor (conv1_t.cu, conv3_wt.cu, stem_umma.cu): tile loads, the 128-byte-swizzle UMMA
// descriptor, the warp transpose-reduction used by the register epilogues and the driver entry point for
// cuTensorMapEncodeTiled (resolved through the runtime, so the library does not link libcuda).
#pragma once
#include <cuda.h>

#include "umma_common.cuh"

namespace smg {

__device__ __forceinline__ void tma_tile_3d(void* smem_dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::
            "r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ void tma_tile_4d(void* smem_dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3,
                                            uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::
            "r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
        : "memory");
}

// L2 prefetch of a tensor-map box: HBM -> L2 only, no shared memory and no barrier involved.  Lets a loader run further ahead
// than its shared-memory ring (the later cp.async.bulk.tensor of the same box then completes with L2 latency).
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* tm, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(reinterpret_cast<uint64_t>(tm)),
                 "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* tm, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(reinterpret_cast<uint64_t>(tm)),
                 "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}

// non-blocking probe of an mbarrier phase
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}

// SWIZZLE_128B K-major operand: rows of 128 B, 8-row atoms of 1024 B (SBO), LBO unused (1), version 1, layout type 2.
// Measured on B200: the swizzle phase follows the ABSOLUTE shared-memory address, so a start address advanced by whole
// rows (+128 B each) or by K steps (+32 B) needs no base_offset as long as the buffer itself is 1024-byte aligned.
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// lane l ends with the sum over the warp's 32 lanes of x[l]; 31 shuffles (halving exchange); x is destroyed
__device__ __forceinline__ float warp_transpose_sum(float (&x)[32], int lane) {
#pragma unroll
    for (int half = 16; half >= 1; half >>= 1) {
        const bool up = (lane & half) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float send = up ? x[i] : x[i + half];
            const float keep = up ? x[i + half] : x[i];
            x[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
        }
    }
    return x[0];
}

// 32 lanes x 16 columns of 32-bit values from registers into tensor memory (thread = TMEM lane of the warp's quadrant);
// callers issue tcgen05.wait::st before they signal the consumer
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float4& a, const float4& b, const float4& c, const float4& d) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(__float_as_uint(a.x)), "r"(__float_as_uint(a.y)), "r"(__float_as_uint(a.z)), "r"(__float_as_uint(a.w)),
        "r"(__float_as_uint(b.x)), "r"(__float_as_uint(b.y)), "r"(__float_as_uint(b.z)), "r"(__float_as_uint(b.w)),
        "r"(__float_as_uint(c.x)), "r"(__float_as_uint(c.y)), "r"(__float_as_uint(c.z)), "r"(__float_as_uint(c.w)),
        "r"(__float_as_uint(d.x)), "r"(__float_as_uint(d.y)), "r"(__float_as_uint(d.z)), "r"(__float_as_uint(d.w))
        : "memory");
}

// tcgen05.mma kind::tf32 with the A operand in tensor memory and the B operand in shared memory
__device__ __forceinline__ void umma_ts_tf32(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn();   // tma_common.cu; nullptr if the driver does not export it

// swizzle_bytes value selecting the 128-byte swizzle with 32-byte atomicity (CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): the
// shared-memory image MN-major tf32 UMMA operands need (layout type SWIZZLE_128B_BASE32B)
constexpr int kSwizzle128Atom32 = 12832;

// fp32 tensor map with 128- or 64-byte swizzle and zero fill; dims/strides innermost first (strides in bytes, rank-1 of them)
int make_tensor_map_f32(CUtensorMap* tm, const float* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                        const cuuint32_t* box, int swizzle_bytes = 128);

// SWIZZLE_64B K-major operand: rows of 64 B, 8-row atoms of 512 B (SBO), layout type 4; same absolute-address rule
__device__ __forceinline__ uint64_t make_desc_sw64(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;
    return d;
}

}  // namespace smg
