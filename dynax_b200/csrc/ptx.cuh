// Thin inline-PTX wrappers for sm_100a: mbarrier, bulk async copy (TMA 1-D, SASS UBLKCP),
// proxy fences, cache policies.  No CUTLASS dependency.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dynax {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

// Make mbarrier.init visible to the async proxy (TMA complete_tx) before first use.
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

// try_wait suspends in hardware for a bounded time, so this is not a hot spin.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// L2 eviction policy for data that is streamed exactly once.
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

// Producer-warp fallback for rows that TMA cannot move (not 16-byte aligned / sized): `rows` rows of `count`
// floats, global row stride `src_stride`, shared row stride `dst_stride`.  Narrow rows (a handful of systems:
// the reference's own single-building examples) are flattened over the lanes so that the loads of ALL rows are
// in flight together instead of one dependent round trip per row.
__device__ __forceinline__ void copy_rows_plain(float *dst, int dst_stride, const float *src, int64_t src_stride,
                                                int rows, int count, int lane) {
    if (count < 32) {
        const int total = rows * count;
#pragma unroll 4
        for (int e = lane; e < total; e += 32) {
            const int r = e / count, i = e - r * count;
            dst[r * dst_stride + i] = __ldg(src + r * src_stride + i);
        }
    } else {
        for (int r = 0; r < rows; ++r) {
#pragma unroll 4
            for (int i = lane; i < count; i += 32) dst[r * dst_stride + i] = __ldg(src + r * src_stride + i);
        }
    }
}

// Asynchronous flavour of copy_rows_plain: 4-byte cp.async (LDGSTS) per element, so the producer warp only ISSUES
// the copies (no load round trips on its critical path); completion is reported to an mbarrier by
// cp_async_arrive() from every lane (barrier count 32).  Used when rows are not TMA-able (N % 4 != 0).
__device__ __forceinline__ void cp_async4(float *smem_dst, const float *gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_arrive(uint64_t *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void copy_rows_async(float *dst, int dst_stride, const float *src, int64_t src_stride,
                                                int rows, int count, int lane) {
    if (count < 32) {
        const int total = rows * count;
        for (int e = lane; e < total; e += 32) {
            const int r = e / count, i = e - r * count;
            cp_async4(dst + r * dst_stride + i, src + r * src_stride + i);
        }
    } else {
        for (int r = 0; r < rows; ++r) {
#pragma unroll 4
            for (int i = lane; i < count; i += 32) cp_async4(dst + r * dst_stride + i, src + r * src_stride + i);
        }
    }
}

// Rows that start on a 4-byte boundary only (N % 4 != 0): row r lands in its shared slot at the offset its global address
// has modulo 16 bytes (`row_shift`), so that the interior of the row is 16-byte aligned on both sides (see
// copy_rows_shift_bulk below).  Consumers read row r at dst + r * dst_pitch + row_shift(src + r * src_stride).
__device__ __forceinline__ int row_shift(const float *row) { return (int)((reinterpret_cast<uintptr_t>(row) >> 2) & 3); }

// L2 policy for data every CTA re-reads (a shared input series).
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

// global -> shared bulk copy, completion signalled on an mbarrier (complete_tx::bytes).
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar,
                                         uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

// global -> shared tensor copy through a 3-D tensor map (tmap.cuh): one box per call, completion (always the full
// box in bytes) on an mbarrier.  dst 128-byte aligned; `map` must live in param/const space (__grid_constant__).
__device__ __forceinline__ void tensor_g2s_3d(void *smem_dst, const void *map, int c0, int c1, int c2, uint64_t *bar,
                                              uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

// shared -> global tensor copy through a 3-D tensor map (bulk async-group completion); elements of the box that fall
// outside the tensor are not written.
__device__ __forceinline__ void tensor_s2g_3d(const void *map, int c0, int c1, int c2, const void *smem_src) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(map), "r"(c0),
                 "r"(c1), "r"(c2), "r"(smem_u32(smem_src))
                 : "memory");
}

// shared -> global bulk copy (bulk async-group completion).
__device__ __forceinline__ void bulk_s2g(void *gmem_dst, const void *smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
                 "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }

// Wait until at most N of this thread's bulk groups are still READING their shared-memory source.
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

template <int N>
__device__ __forceinline__ void bulk_wait() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// Order generic-proxy st.shared before async-proxy (bulk copy) reads of the same bytes.
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// Shifted layout (row_shift above): the interior of every row as ONE 1-D bulk copy (TMA) -- 16-byte aligned on both sides
// thanks to the shift; lane r issues row r's copy --, the < 16-byte head and tail as 4-byte cp.async.  The bulk bytes are
// announced on `bar` with expect_tx; the caller then lets EVERY lane arrive (cp_async_arrive, barrier count 32), which
// also covers the lanes' own copies.  dst 16-byte aligned, dst_pitch a multiple of 4 floats with room for the shift (row
// length + 4), rows <= 32.  Measured at N = 151 553 (profiles/r02_unaligned_ab.txt): 4-byte cp.async by the producer lanes
// 0.36-0.62 of the HBM roof, 16-byte cp.async by the lanes no better (the producer warp shares its issue slots with the
// computing warps: its instruction count per stage is what matters), one bulk copy per row from lane 0 0.57-0.71, from
// lanes in parallel 0.69-0.82.
__device__ __forceinline__ void copy_rows_shift_bulk(float *dst, int dst_pitch, const float *src, int64_t src_stride, int rows,
                                                     int count, int lane, uint64_t *bar, uint64_t policy) {
    // lane r issues row r's bulk copy (rows <= 32): the copies leave together instead of one every ~45 ns from one thread
    uint32_t body = 0;
    int mis = 0, head = 0;
    if (lane < rows) {
        mis = row_shift(src + lane * src_stride);
        head = (4 - mis) & 3;
        head = head < count ? head : count;
        body = (uint32_t)((count - head) >> 2) * 16u;
    }
    uint32_t bytes = body;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) bytes += __shfl_xor_sync(0xffffffffu, bytes, o);
    if (lane == 0 && bytes) mbar_expect_tx(bar, bytes);
    __syncwarp();
    if (lane < rows && body) bulk_g2s(dst + lane * dst_pitch + mis + head, src + lane * src_stride + head, body, bar, policy);
    // head and tail: lane -> (row lane / 4 of a group of eight rows, element lane % 4 < 3), two predicated copies per lane
    for (int r0 = 0; r0 < rows; r0 += 8) {
        const int r = r0 + (lane >> 2), e = lane & 3;
        if (r < rows && e < 3) {
            const float *s = src + r * src_stride;
            const int m2 = row_shift(s);
            int h2 = (4 - m2) & 3;
            h2 = h2 < count ? h2 : count;
            const int b2 = (count - h2) & ~3, t2 = count - h2 - b2;
            float *d = dst + r * dst_pitch + m2;
            if (e < h2) cp_async4(d + e, s + e);
            if (e < t2) cp_async4(d + h2 + b2 + e, s + h2 + b2 + e);
        }
    }
}

}  // namespace ptx
}  // namespace dynax
