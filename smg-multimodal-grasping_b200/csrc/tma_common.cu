// tma_common.cu - host helpers shared by the TMA-fed kernels: the driver entry point of cuTensorMapEncodeTiled (resolved
// through the runtime, so the library does not link libcuda), fp32 tensor maps, and the per-handle (= per-device) opt-in
// to more than 48 KB of dynamic shared memory.
#include "tma_common.cuh"

namespace smg {

EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int make_tensor_map_f32(CUtensorMap* tm, const float* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                        const cuuint32_t* box, int swizzle_bytes) {
    EncodeTiledFn enc = encode_tiled_fn();
    SMG_CHECK(enc != nullptr, SMG_ERR_CUDA, "tensor map: cuTensorMapEncodeTiled not available from the driver");
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<float*>(base), dims, strides, box,
                           estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                               : (swizzle_bytes == kSwizzle128Atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B),
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SMG_CHECK(r == CUDA_SUCCESS, SMG_ERR_CUDA, "tensor map: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return SMG_OK;
}

// Function attributes are per device: the opt-in is remembered in the handle (one handle = one GPU), not in a process-wide
// static, so a second handle on another GPU of the same process sets it again for its own device.
int ensure_dyn_smem(smg_handle* h, const void* kernel, int bytes) {
    for (const void* k : h->smem_opt_in)
        if (k == kernel) return SMG_OK;
    SMG_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    h->smem_opt_in.push_back(kernel);
    return SMG_OK;
}

}  // namespace smg
