// MLP-dynamics RK4 rollout on the 5th-gen tensor cores (tcgen05 + TMEM) -- config 5 / row A7.
//
// The MLP layers are real dense contractions ACROSS THE BATCH: a CTA owns 128 trajectories = the M
// dimension of a UMMA tile.  Per MLP evaluation (4 per RK4 step):
//
//   row thread r (TMEM lane r)      z_r = [x_r ; u_r]  --split hi/lo-->  tcgen05.st  -> A (TMEM, 8 cols)
//   one elected thread              D  = A x W1^T      3 x tcgen05.mma.kind::tf32 (M128 N64 K8),  A from TMEM,
//                                                      B = W1 in smem (K-major, no swizzle), D in TMEM (64 cols)
//   row thread r                    tcgen05.ld D -> h1 = tanh(D + b1) -> split hi/lo -> tcgen05.st -> A (64 cols)
//   one elected thread              D  = h1 x W2^T     3 x 8 tcgen05.mma (K = 8 per instruction)
//   row thread r                    tcgen05.ld D -> h2 = tanh(D + b2);  rhs = W3 h2 + b3 on the FMA pipe (N=3)
//
// fp32-grade accuracy from TF32 tensor cores: every operand is split a = a_hi + a_lo (a_hi = top 19
// bits), and a.b ~= a_hi.b_hi + a_lo.b_hi + a_hi.b_lo (3 MMAs, error ~2^-21 relative); accumulation
// is fp32 in TMEM.  tanh = 1 - 2/(2^(2x log2 e) + 1) with ex2.approx/rcp.approx (abs error ~5e-7),
// evaluated in pairs that share one reciprocal (3 MUFU ops per 2 tanh):
// MUFU.TANH's 2^-11 would break the 1e-5 trajectory tolerance.
//
// u rows / x,y rows stream through the same TMA (cp.async.bulk) + mbarrier rings as rollout_fwd_kernel.
// TMEM: 256 columns per CTA (A_hi 64 | A_lo 64 | D 64) -> 2 CTAs per SM; smem ~52 KB per CTA.
#pragma once
#include "ptx.cuh"

namespace dynax {

struct NodeTcParams {
    const float *W1, *b1, *W2, *b2, *W3, *b3;
    const float *x0, *u;
    float *xsol, *ysol, *x_final;
    float *stage_x;  // [T,N,9] or NULL: x parts of the RK4 stage points z2, z3, z4 of every step (saved for the VJP)
    int64_t N, T;
    float dt;
    int rk4, bulk_in, bulk_out;
};

namespace tc {

constexpr int TN = 128, HID = 64, NX = 3, NU = 5, NIN = 8, S = 4;
constexpr uint32_t TMEM_COLS = 256, COL_AHI = 0, COL_ALO = 64, COL_D = 128;
// kind::tf32, D=F32, A=B=TF32, both K-major, N=64, M=128   (cute/arch/mma_sm100_desc.hpp InstrDescriptor)
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
// canonical K-major, no-swizzle operand tile: core matrix = 8 rows x 16 B; SBO = 128 B between 8-row
// groups, LBO = 1024 B between 16-byte K chunks (8 row groups of 128 B each)
constexpr uint32_t SBO = 128, LBO = 1024;

__device__ __forceinline__ int wofs(int n, int k) { return (k >> 2) * 256 + (n >> 3) * 32 + (n & 7) * 4 + (k & 3); }

__device__ __forceinline__ uint64_t smem_desc(const void *p) {
    const uint64_t addr = (uint64_t)(ptx::smem_u32(p) >> 4) & 0x3fffu;
    return addr | ((uint64_t)(LBO >> 4) << 16) | ((uint64_t)(SBO >> 4) << 32) | (1ull << 46);
}

__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(IDESC), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(ptx::smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void rows_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld16(uint32_t addr, float *v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(addr)
        : "memory");
    wait_ld();
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_st16(uint32_t addr, const uint32_t *r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        ::"r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}

__device__ __forceinline__ void tmem_st8(uint32_t addr, const uint32_t *r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(addr), "r"(r[0]),
                 "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}

// a = hi + lo with hi = top 19 bits (what a TF32 operand keeps); lo is exact in fp32
__device__ __forceinline__ void split_tf32(float a, uint32_t &hi, uint32_t &lo) {
    hi = __float_as_uint(a) & 0xffffe000u;
    lo = __float_as_uint(a - __uint_as_float(hi));
}

// Two tanh for THREE MUFU ops (the kernel is bound by the MUFU pipe): tanh(x) = 1 - 2/(e^{2x}+1) and the two
// reciprocals share one rcp:  r = 1/(a b),  1/a = r b,  1/b = r a.  The exponent is clamped at 2^60 so
// a*b stays finite (tanh is 1 to fp32 precision far below that).
__device__ __forceinline__ void fast_tanh2(float x0, float x1, float &t0, float &t1) {
    float e0, e1, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fminf(x0 * 2.885390081777927f, 60.f)));  // e^(2x)
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fminf(x1 * 2.885390081777927f, 60.f)));
    const float a = e0 + 1.0f, b = e1 + 1.0f;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a * b));
    t0 = fmaf(-2.0f, r * b, 1.0f);
    t1 = fmaf(-2.0f, r * a, 1.0f);
}

// shared-memory map (floats unless noted)
constexpr int F_W2HI = 0, F_W2LO = F_W2HI + HID * HID, F_W1HI = F_W2LO + HID * HID, F_W1LO = F_W1HI + HID * NIN,
              F_B1 = F_W1LO + HID * NIN, F_B2 = F_B1 + HID, F_W3 = F_B2 + HID, F_B3 = F_W3 + NX * HID,
              F_IN = F_B3 + 4, F_XO = F_IN + S * TN * NU, F_YO = F_XO + 2 * TN * NX, F_ZO = F_YO + 2 * TN,
              F_END = F_ZO + 2 * TN * 3 * NX;
constexpr int NBAR = 2 * S + 4 + 1;
constexpr size_t SMEM_USED = sizeof(float) * F_END + sizeof(uint64_t) * NBAR + 16;
// TMEM holds 512 columns per SM and each CTA allocates 256: pad the request so at most 2 CTAs are
// resident (a third would only block inside tcgen05.alloc while pinning registers and smem)
constexpr size_t SMEM_BYTES = SMEM_USED > 100 * 1024 ? SMEM_USED : 100 * 1024;

}  // namespace tc

#ifndef DYNAX_NODE_TC_HELPERS_ONLY  // node_vjp_tc.cuh reuses the tc:: helpers above without this kernel
__global__ void __launch_bounds__(tc::TN + 32, 2) node_tc_kernel(const NodeTcParams P) {
    using namespace tc;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *sm = reinterpret_cast<float *>(smem_raw);
    float *s_in = sm + F_IN, *s_xo = sm + F_XO, *s_yo = sm + F_YO, *s_zo = sm + F_ZO;
    uint64_t *full = reinterpret_cast<uint64_t *>(sm + F_END);
    uint64_t *empty = full + S, *ofull = empty + S, *ofree = ofull + 2, *mma_bar = ofree + 2;
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(mma_bar + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t N = P.N, T = P.T;
    const int64_t n0 = (int64_t)blockIdx.x * TN;
    const int valid = (int)((N - n0) < TN ? (N - n0) : TN);
    const bool emit_x = P.xsol != nullptr, emit_y = P.ysol != nullptr, emit_z = P.stage_x != nullptr && P.rk4;
    const bool emit = emit_x || emit_y || emit_z;

    // ---- one-time setup: barriers, weights split hi/lo into the canonical UMMA layout, TMEM ----
    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            ptx::mbar_init(&full[s], 1);
            ptx::mbar_init(&empty[s], TN / 32);
        }
        ptx::mbar_init(&ofull[0], TN / 32); ptx::mbar_init(&ofull[1], TN / 32);
        ptx::mbar_init(&ofree[0], 1); ptx::mbar_init(&ofree[1], 1);
        ptx::mbar_init(mma_bar, 1);
        ptx::fence_mbar_init();
    }
    for (int i = tid; i < HID * HID; i += blockDim.x) {
        const int nrow = i / HID, k = i % HID;
        uint32_t hi, lo;
        split_tf32(__ldg(P.W2 + i), hi, lo);
        sm[F_W2HI + wofs(nrow, k)] = __uint_as_float(hi);
        sm[F_W2LO + wofs(nrow, k)] = __uint_as_float(lo);
    }
    for (int i = tid; i < HID * NIN; i += blockDim.x) {
        const int nrow = i / NIN, k = i % NIN;
        uint32_t hi, lo;
        split_tf32(__ldg(P.W1 + i), hi, lo);
        sm[F_W1HI + wofs(nrow, k)] = __uint_as_float(hi);
        sm[F_W1LO + wofs(nrow, k)] = __uint_as_float(lo);
    }
    for (int i = tid; i < HID; i += blockDim.x) {
        sm[F_B1 + i] = __ldg(P.b1 + i);
        sm[F_B2 + i] = __ldg(P.b2 + i);
    }
    for (int i = tid; i < NX * HID; i += blockDim.x) sm[F_W3 + i] = __ldg(P.W3 + i);
    if (tid < NX) sm[F_B3 + tid] = __ldg(P.b3 + tid);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(s_tmem)),
                     "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    ptx::fence_proxy_async_smem();  // weights were written through the generic proxy; UMMA reads them via the async proxy
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *s_tmem;

    if (warp == TN / 32) {
        // ------------------------------- producer warp (same protocol as rollout_fwd_kernel) -----------
        const uint64_t pol = ptx::policy_evict_first();
        auto load_step = [&](int64_t c) {
            const int s = (int)(c % S);
            float *dst = s_in + s * TN * NU;
            const float *src = P.u + (c * N + n0) * NU;
            if (P.bulk_in) {
                if (lane == 0) {
                    const uint32_t bytes = (uint32_t)valid * NU * 4u;
                    ptx::mbar_arrive_expect_tx(&full[s], bytes);
                    ptx::bulk_g2s(dst, src, bytes, &full[s], pol);
                }
            } else {
                for (int i = lane; i < valid * NU; i += 32) dst[i] = __ldg(src + i);
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&full[s]);
            }
        };
        auto store_step = [&](int64_t c) {
            const int o = (int)(c & 1);
            const float *sx = s_xo + o * TN * NX, *sy = s_yo + o * TN, *sz = s_zo + o * TN * 3 * NX;
            if (P.bulk_out) {
                if (lane == 0) {
                    if (emit_x) ptx::bulk_s2g(P.xsol + (c * N + n0) * NX, sx, (uint32_t)valid * NX * 4u);
                    if (emit_y) ptx::bulk_s2g(P.ysol + (c * N + n0), sy, (uint32_t)valid * 4u);
                    if (emit_z) ptx::bulk_s2g(P.stage_x + (c * N + n0) * 3 * NX, sz, (uint32_t)valid * 3 * NX * 4u);
                    ptx::bulk_commit();
                    ptx::bulk_wait_read<0>();
                    ptx::mbar_arrive(&ofree[o]);
                }
            } else {
                if (emit_x) for (int i = lane; i < valid * NX; i += 32) P.xsol[(c * N + n0) * NX + i] = sx[i];
                if (emit_y) for (int i = lane; i < valid; i += 32) P.ysol[(c * N + n0) + i] = sy[i];
                if (emit_z) for (int i = lane; i < valid * 3 * NX; i += 32) P.stage_x[(c * N + n0) * 3 * NX + i] = sz[i];
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&ofree[o]);
            }
        };
        for (int64_t c = 0; c < T && c < S; ++c) load_step(c);
        for (int64_t c = 0; c < T; ++c) {
            ptx::mbar_wait(&empty[c % S], (uint32_t)((c / S) & 1));
            if (c + S < T) load_step(c + S);
            if (emit) {
                ptx::mbar_wait(&ofull[c & 1], (uint32_t)((c >> 1) & 1));
                store_step(c);
            }
        }
        if (lane == 0) ptx::bulk_wait<0>();
        return;
    }

    // --------------------------------- row threads: one trajectory each ---------------------------------
    const bool active = tid < valid;
    const int64_t i = active ? (n0 + tid) : (N - 1);
    float x[NX];
#pragma unroll
    for (int d = 0; d < NX; ++d) x[d] = __ldg(P.x0 + i * NX + d);
    const float dt = P.dt;
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);  // this warp's 32 TMEM lanes
    const uint64_t dW1hi = smem_desc(sm + F_W1HI), dW1lo = smem_desc(sm + F_W1LO);
    const uint64_t dW2hi = smem_desc(sm + F_W2HI), dW2lo = smem_desc(sm + F_W2LO);
    uint32_t ph = 0;  // parity of mma_bar

    // one MLP evaluation for the CTA's 128 rows; every row thread calls it with its own (z, r)
    auto mlp = [&](const float *z, float *r) {
        uint32_t hi[16], lo[16];
        // ---- layer 1: D = z W1^T
#pragma unroll
        for (int c = 0; c < NIN; ++c) split_tf32(z[c], hi[c], lo[c]);
        tmem_st8(lane_base + COL_AHI, hi);
        tmem_st8(lane_base + COL_ALO, lo);
        wait_st();
        fence_before();
        rows_sync();
        if (tid == 0) {
            fence_after();
            mma_ts(tmem + COL_D, tmem + COL_AHI, dW1hi, 0u);
            mma_ts(tmem + COL_D, tmem + COL_ALO, dW1hi, 1u);
            mma_ts(tmem + COL_D, tmem + COL_AHI, dW1lo, 1u);
            mma_commit(mma_bar);
        }
        ptx::mbar_wait(mma_bar, ph);
        ph ^= 1u;
        fence_after();
        // ---- h1 = tanh(D + b1) -> A (hi/lo)
#pragma unroll
        for (int c = 0; c < HID; c += 16) {
            float v[16];
            tmem_ld16(lane_base + COL_D + c, v);
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
                float t0, t1;
                fast_tanh2(v[j] + sm[F_B1 + c + j], v[j + 1] + sm[F_B1 + c + j + 1], t0, t1);
                split_tf32(t0, hi[j], lo[j]);
                split_tf32(t1, hi[j + 1], lo[j + 1]);
            }
            tmem_st16(lane_base + COL_AHI + c, hi);
            tmem_st16(lane_base + COL_ALO + c, lo);
        }
        wait_st();
        fence_before();
        rows_sync();
        // ---- layer 2: D = h1 W2^T   (K = 64 = 8 instructions of K = 8, times the 3 split products)
        if (tid == 0) {
            fence_after();
#pragma unroll
            for (int s = 0; s < HID / 8; ++s)  // descriptor start address advances 2 K-chunks = 2*LBO bytes (>>4)
                mma_ts(tmem + COL_D, tmem + COL_AHI + 8 * s, dW2hi + (uint64_t)((2 * LBO * s) >> 4), s ? 1u : 0u);
#pragma unroll
            for (int s = 0; s < HID / 8; ++s)
                mma_ts(tmem + COL_D, tmem + COL_ALO + 8 * s, dW2hi + (uint64_t)((2 * LBO * s) >> 4), 1u);
#pragma unroll
            for (int s = 0; s < HID / 8; ++s)
                mma_ts(tmem + COL_D, tmem + COL_AHI + 8 * s, dW2lo + (uint64_t)((2 * LBO * s) >> 4), 1u);
            mma_commit(mma_bar);
        }
        ptx::mbar_wait(mma_bar, ph);
        ph ^= 1u;
        fence_after();
        // ---- h2 = tanh(D + b2);  rhs = W3 h2 + b3  (N = 3: FMA pipe)
#pragma unroll
        for (int d = 0; d < NX; ++d) r[d] = sm[F_B3 + d];
#pragma unroll
        for (int c = 0; c < HID; c += 16) {
            float v[16];
            tmem_ld16(lane_base + COL_D + c, v);
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
                float h0, h1;
                fast_tanh2(v[j] + sm[F_B2 + c + j], v[j + 1] + sm[F_B2 + c + j + 1], h0, h1);
#pragma unroll
                for (int d = 0; d < NX; ++d) {
                    r[d] = fmaf(sm[F_W3 + d * HID + c + j], h0, r[d]);
                    r[d] = fmaf(sm[F_W3 + d * HID + c + j + 1], h1, r[d]);
                }
            }
        }
    };

    const int stages = P.rk4 ? 4 : 1;
    const float hs = P.rk4 ? dt / 6.0f : dt;
    for (int64_t c = 0; c < T; ++c) {
        const int s = (int)(c % S), o = (int)(c & 1);
        ptx::mbar_wait(&full[s], (uint32_t)((c / S) & 1));
        if (emit) ptx::mbar_wait(&ofree[o], (uint32_t)(((c >> 1) & 1) ^ 1));
        float z[NIN], k[NX], acc[NX];
#pragma unroll
        for (int d = 0; d < NU; ++d) z[NX + d] = s_in[s * TN * NU + tid * NU + d];
        if (emit_y) s_yo[o * TN + tid] = x[0];  // y_k = x_k[0], pre-update
#pragma unroll
        for (int d = 0; d < NX; ++d) {
            z[d] = x[d];
            acc[d] = 0.f;
        }
#pragma unroll 1
        for (int st = 0; st < stages; ++st) {
            mlp(z, k);
            const float wk = (st == 1 || st == 2) ? 2.0f : 1.0f;
            const float cs = (st == 2) ? dt : 0.5f * dt;
#pragma unroll
            for (int d = 0; d < NX; ++d) {
                acc[d] = fmaf(wk, k[d], acc[d]);
                z[d] = fmaf(cs, k[d], x[d]);
            }
            if (emit_z && st < 3) {  // the next stage point (z2, z3, z4): what the VJP would otherwise recompute
#pragma unroll
                for (int d = 0; d < NX; ++d) s_zo[o * TN * 3 * NX + tid * 3 * NX + st * NX + d] = z[d];
            }
        }
#pragma unroll
        for (int d = 0; d < NX; ++d) x[d] = fmaf(hs, acc[d], x[d]);
        if (emit_x) {
#pragma unroll
            for (int d = 0; d < NX; ++d) s_xo[o * TN * NX + tid * NX + d] = x[d];
        }
        if (emit) ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
            ptx::mbar_arrive(&empty[s]);
            if (emit) ptx::mbar_arrive(&ofull[o]);
        }
    }
    if (P.x_final != nullptr && active) {
#pragma unroll
        for (int d = 0; d < NX; ++d) P.x_final[i * NX + d] = x[d];
    }
    // ---- teardown: every row thread is done with TMEM before warp 0 frees it
    fence_before();
    rows_sync();
    if (warp == 0) {
        fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
    }
}
#endif  // DYNAX_NODE_TC_HELPERS_ONLY

}  // namespace dynax
