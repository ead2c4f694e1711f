// Micro-benchmark: how fast can ONE producer thread per CTA feed a shared-memory ring with rows of a [T, N*w] fp32
// array -- (a) one 1-D cp.async.bulk per row (what the rollout kernels did in round 1), (b) one cp.async.bulk.tensor
// (3-D tensor map: box = KF rows x row) per stage.  No compute: the "consumer" (the same warp) only waits for the
// stage and releases it, so the time per row is the feed rate of the ring.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/tma_probe tools/tma_probe.cu
//   tools/tma_probe        (prints a table; ~2 s)
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) {                                                                \
            printf("%s failed: %s (%s:%d)\n", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(1);                                                                            \
        }                                                                                       \
    } while (0)

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t c) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c) : "memory");
}
__device__ __forceinline__ void mbar_expect(uint64_t *b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(s32(b)), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void bulk1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)),
                 "l"(src), "r"(bytes), "r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void tensor3d(void *dst, const CUtensorMap *map, int c0, int c1, int c2, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(s32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(s32(bar)) : "memory");
}

// row = row_floats fp32 of this CTA's slice; stage = KF rows; S stages.  MODE 0: 1-D copies, 1: one tensor copy per stage
template <int MODE>
__global__ void probe(const float *g, const __grid_constant__ CUtensorMap map, int64_t row_stride_floats, int row_floats,
                      int inner, int KF, int S, int chunks, float *sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    float *ring = reinterpret_cast<float *>(smem);
    const int stage_floats = KF * row_floats;
    uint64_t *full = reinterpret_cast<uint64_t *>(ring + (size_t)S * stage_floats);
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const int64_t col0 = (int64_t)blockIdx.x * row_floats;
    auto load = [&](int c) {
        const int s = c % S;
        mbar_expect(&full[s], (uint32_t)stage_floats * 4u);
        if (MODE == 0) {
            for (int r = 0; r < KF; ++r)
                bulk1d(ring + (size_t)s * stage_floats + (size_t)r * row_floats, g + ((int64_t)c * KF + r) * row_stride_floats + col0,
                       (uint32_t)row_floats * 4u, &full[s]);
        } else {
            tensor3d(ring + (size_t)s * stage_floats, &map, 0, (int)(col0 / inner), c * KF, &full[s]);
        }
    };
    float acc = 0.f;
    for (int c = 0; c < S && c < chunks; ++c) load(c);
    for (int c = 0; c < chunks; ++c) {
        const int s = c % S;
        mbar_wait(&full[s], (uint32_t)((c / S) & 1));
        acc += ring[(size_t)s * stage_floats];   // touch
        if (c + S < chunks) load(c + S);
    }
    if (acc == 123.456f) sink[0] = acc;
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    EncodeFn encode = (EncodeFn)fn;
    const int T = 4096;
    float *sink;
    CK(cudaMalloc(&sink, 4));
    printf("%-8s %-6s %-4s %-3s %-5s %10s %10s %10s\n", "mode", "row_B", "KF", "S", "cta/sm", "ns/row", "ns/copy", "GB/s(chip)");
    const int row_floats_list[] = {32, 128, 160, 320, 640, 768, 1280};
    for (int cps = 1; cps <= 2; ++cps)
        for (int rf : row_floats_list)
            for (int S = 2; S <= 4; S += 2)
                for (int mode = 0; mode < 2; ++mode) {
                    const int KF = 16, ctas = sms * cps;
                    const int64_t ncols = (int64_t)ctas * rf;   // floats per global row
                    float *g;
                    CK(cudaMalloc(&g, sizeof(float) * (size_t)T * ncols));
                    CK(cudaMemset(g, 0, sizeof(float) * (size_t)T * ncols));
                    int inner = rf % 160 == 0 ? 160 : (rf % 128 == 0 ? 128 : rf % 32 == 0 ? 32 : rf);
                    if (inner > 256) inner = 32;
                    CUtensorMap map;
                    cuuint64_t dims[3] = {(cuuint64_t)inner, (cuuint64_t)(ncols / inner), (cuuint64_t)T};
                    cuuint64_t strides[2] = {(cuuint64_t)inner * 4, (cuuint64_t)ncols * 4};
                    cuuint32_t box[3] = {(cuuint32_t)inner, (cuuint32_t)(rf / inner), (cuuint32_t)KF};
                    cuuint32_t es[3] = {1, 1, 1};
                    CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, g, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                    if (r != CUDA_SUCCESS) { printf("encode failed %d (rf=%d inner=%d)\n", (int)r, rf, inner); return 1; }
                    const size_t smem = sizeof(float) * (size_t)S * KF * rf + 8 * S + 128;
                    auto k = mode == 0 ? probe<0> : probe<1>;
                    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    const int chunks = T / KF;
                    cudaEvent_t e0, e1;
                    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
                    k<<<ctas, 32, smem>>>(g, map, ncols, rf, inner, KF, S, chunks, sink);
                    CK(cudaDeviceSynchronize());
                    CK(cudaEventRecord(e0));
                    for (int it = 0; it < 5; ++it) k<<<ctas, 32, smem>>>(g, map, ncols, rf, inner, KF, S, chunks, sink);
                    CK(cudaEventRecord(e1));
                    CK(cudaDeviceSynchronize());
                    float ms;
                    CK(cudaEventElapsedTime(&ms, e0, e1));
                    ms /= 5;
                    const double ns_row = ms * 1e6 / T;
                    printf("%-8s %-6d %-4d %-3d %-5d %10.1f %10.1f %10.0f\n", mode == 0 ? "bulk1d" : "tensor3d", rf * 4, KF, S, cps, ns_row,
                           mode == 0 ? ns_row : ns_row * KF, (double)T * ncols * 4 / (ms * 1e-3) / 1e9);
                    CK(cudaFree(g));
                }
    return 0;
}
