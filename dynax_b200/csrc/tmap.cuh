// TMA tensor maps over the [T, N*w] row arrays of the rollout kernels (u[T,N,5], target[T,N], ctrl[T,N]).
//
// A CTA owns TN consecutive systems, i.e. TN*w contiguous floats of every time row.  Round 1 moved each row with its
// own 1-D cp.async.bulk; measured on B200 (tools/tma_probe.cu -> profiles/r02_tma_probe.txt) one producer thread
// issues such a copy every ~45 ns whatever its size below ~2 KB, so narrow tiles (few systems per CTA: the launches
// that cannot fill the GPU) were bound by the number of copies, not by bytes.  A 3-D tensor map
//     dims = (inner, N*w/inner, T)      box = (inner, TN*w/inner, KF)
// moves a whole stage -- KF rows of a CTA's slice -- with ONE cp.async.bulk.tensor (~17 ns per row for 128..640-byte
// rows, bandwidth-bound above).  The shared-memory image is the row-major [KF][TN*w] block the consumers read, rows
// past the end of the horizon or of the system axis arrive as zeros, and the transaction count is always the full
// box.  `inner` = the largest of a few candidates (<= 256 elements, the box limit) dividing both N*w and TN*w.
#pragma once
#include <cuda.h>

#include "api_common.cuh"

namespace dynax {

typedef CUresult (*TensorMapEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                      const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline TensorMapEncodeFn tensor_map_encoder() {
    static TensorMapEncodeFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<TensorMapEncodeFn>(p);
    }();
    return fn;
}

// Largest inner extent (floats) usable for rows of N*w floats tiled TN*w at a time; 0 = no tensor path (the caller
// keeps the per-row copies).  At least 32 floats: shorter lines made the copy slower than per-row copies.
inline int row_map_inner(int64_t N, int w, int TN) {
    static const int cand[] = {256, 160, 128, 80, 64, 40, 32};
    for (int c : cand)
        if ((N * w) % c == 0 && (TN * w) % c == 0 && (TN * w) / c <= 256) return c;
    return 0;
}

// base: [T, N*w] fp32, 16-byte aligned, N % 4 == 0.  Returns DYNAX_OK and fills *map, or an error.
inline int make_row_map(CUtensorMap *map, const float *base, int64_t N, int w, int64_t T, int TN, int KF, int inner) {
    TensorMapEncodeFn enc = tensor_map_encoder();
    if (!enc) return fail(DYNAX_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t dims[3] = {(cuuint64_t)inner, (cuuint64_t)(N * w / inner), (cuuint64_t)T};
    cuuint64_t strides[2] = {(cuuint64_t)inner * 4, (cuuint64_t)N * w * 4};
    cuuint32_t box[3] = {(cuuint32_t)inner, (cuuint32_t)(TN * w / inner), (cuuint32_t)KF};
    cuuint32_t es[3] = {1, 1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(base), dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(DYNAX_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for N=%lld w=%d TN=%d KF=%d inner=%d",
                                       (int)r, (long long)N, w, TN, KF, inner);
    return DYNAX_OK;
}

}  // namespace dynax
