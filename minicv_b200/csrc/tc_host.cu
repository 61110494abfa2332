// tc_host.cu -- host helpers of the tcgen05 kernels: the TMA tensor-map builder.
// cuTensorMapEncodeTiled is a driver-API entry point; it is resolved through the runtime
// (cudaGetDriverEntryPoint) so the library does not link libcuda directly.
#include "tc_ptx.cuh"

#include <algorithm>
#include <mutex>

namespace mcv {
namespace tc {
namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

}  // namespace

int make_map(CUtensorMap* map, const void* base, int64_t rows, int rowBytes) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return -1; }
    const cuuint64_t dims[2] = {(cuuint64_t)rowBytes, (cuuint64_t)std::max<int64_t>(rows, 1)};
    const cuuint64_t strides[1] = {(cuuint64_t)rowBytes};
    const cuuint32_t box[2] = {128, (cuuint32_t)TILE_M};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return -1; }
    return 0;
}

}  // namespace tc
}  // namespace mcv
