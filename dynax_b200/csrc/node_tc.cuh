// MLP-dynamics RK4 rollout on the 5th-gen tensor cores (tcgen05 + TMEM) -- config 5 / row A7.
//
// The MLP layers are real dense contractions ACROSS THE BATCH: a CTA owns 128 trajectories = the M
// dimension of a UMMA tile.  Per MLP evaluation (4 per RK4 step):
//
//   row thread r (TMEM lane r)      z_r = [x_r ; u_r]  --split hi/lo-->  tcgen05.st  -> A (TMEM, 8 cols)
//   MMA warp (one elected lane)     D1 = A x W1^T      3 x tcgen05.mma.kind::tf32 (M128 N=HID K8),  A from TMEM,
//                                                      B = W1 in smem (K-major, no swizzle), D in TMEM
//   row thread r                    tcgen05.ld D1 -> h1 = tanh(D1 + b1) -> split hi/lo -> tcgen05.st -> A, 16 columns
//                                   at a time; each slice is published on its own mbarrier
//   MMA warp                        D2 += h1[slice] x W2[:, slice]^T    3 x 2 tcgen05.mma per slice, issued while the
//                                   row threads are still computing the following slices
//   row thread r                    tcgen05.ld D2 -> h2 = tanh(D2 + b2);  rhs = W3 h2 + b3 on the FMA pipe (N=3)
//
// fp32-grade accuracy from TF32 tensor cores: every operand is split a = a_hi + a_lo (a_hi = top 19
// bits), and a.b ~= a_hi.b_hi + a_lo.b_hi + a_hi.b_lo (3 MMAs, error ~2^-21 relative); accumulation
// is fp32 in TMEM.  tanh from ex2.approx/rcp.approx (abs error ~5e-7); MUFU.TANH's 2^-11 would break the 1e-5
// trajectory tolerance.
//
// u rows / x,y rows stream through the same TMA (cp.async.bulk) + mbarrier rings as rollout_fwd_kernel.
// The tc:: namespace below holds the helpers shared with the VJP kernel (node_vjp_tc.cuh, hidden width 64).
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
    int rk4, bulk_in, bulk_out, shared_u;  // shared_u: u is ONE series [T,NU] for every trajectory
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

// a = hi + lo with hi = a ROUNDED TO NEAREST at 19 bits (what a TF32 operand keeps; cvt.rna.tf32.f32 is the same two
// integer instructions); lo = a - hi is exact in fp32, signed, |lo| <= 2^-12 |a|.  With a truncating split lo always has
// the sign of a, and so has what the tensor core drops of it (it reads 11 significant bits of lo): a bias of ~2^-21 per
// product that accumulates coherently over the 8192 chained evaluations of config 5; rounded, the dropped parts have
// random signs.
__device__ __forceinline__ void split_tf32(float a, uint32_t &hi, uint32_t &lo) {
    hi = (__float_as_uint(a) + 0x1000u) & 0xffffe000u;
    lo = __float_as_uint(a - __uint_as_float(hi));
}

constexpr float K2LOG2E = 2.885390081777927f;

__device__ __forceinline__ float ex2v(float x) {  // volatile: the MUFU stream keeps the order of the source
    float e;
    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x));
    return e;
}
__device__ __forceinline__ float rcpv(float x) {
    float r;
    asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// Four tanh for FIVE MUFU ops, in stages so that the caller can fix the order of the MUFU stream (a warp issues in
// order and the MUFU pipe takes one instruction per 8 cycles: reciprocals queued behind all 16 exponentials of a slice
// would expose the whole tail of the slice).  Inputs: pre-activations (v0,v1),(v2,v3) WITHOUT bias and
// bk = b * 2log2(e) - 1.  With h_i = 2^(s_i - 1) = e^{2x_i}/2 (clamped at 2^29, so the product of four denominators
// stays finite):  a_i = h_i + 1/2,  d_i = h_i - 1/2 (exact),  tanh(x_i) = d_i / a_i;  r = 1/(a0 a1 a2 a3),
// 1/a0 = r (a2 a3) a1, ...  Outputs (tanh0, tanh2), (tanh1, tanh3).
struct TanhQuad {
    float s[4];          // clamped exponents
    float2 a02, a13, d02, d13, pp;
    float r;
    __device__ __forceinline__ void front(float2 v01, float2 v23, float4 bk) {
        const float2 kk = make_float2(K2LOG2E, K2LOG2E);
        const float2 s01 = __ffma2_rn(v01, kk, make_float2(bk.x, bk.y));
        const float2 s23 = __ffma2_rn(v23, kk, make_float2(bk.z, bk.w));
        s[0] = fminf(s01.x, 29.f); s[1] = fminf(s01.y, 29.f); s[2] = fminf(s23.x, 29.f); s[3] = fminf(s23.y, 29.f);
    }
    __device__ __forceinline__ void exps() {  // 4 MUFU
        const float e0 = ex2v(s[0]), e1 = ex2v(s[1]), e2 = ex2v(s[2]), e3 = ex2v(s[3]);
        const float2 half2 = make_float2(0.5f, 0.5f), mhalf2 = make_float2(-0.5f, -0.5f);
        const float2 h02 = make_float2(e0, e2), h13 = make_float2(e1, e3);
        a02 = __fadd2_rn(h02, half2); a13 = __fadd2_rn(h13, half2);
        d02 = __fadd2_rn(h02, mhalf2); d13 = __fadd2_rn(h13, mhalf2);
        pp = __fmul2_rn(a02, a13);  // (a0 a1, a2 a3)
    }
    __device__ __forceinline__ void recip() { r = rcpv(pp.x * pp.y); }  // 1 MUFU
    // d/a rather than 1 - 1/a: d is exact, so small |x| keep their RELATIVE accuracy up to ex2.approx's own error (1 - 1/a
    // cancels: 2.5e-7 absolute whatever |tanh|; measured on config 5 at dt = 900, 2048 steps: 90th percentile of the
    // per-trajectory error 6.5e-5 with 1 - 1/a, 3.8e-5 with d/a), for two more packed multiplications per pair
    // (-2.7 % throughput).
    __device__ __forceinline__ void finish(float2 &t02, float2 &t13) const {
        const float2 q = make_float2(r * pp.y, r * pp.x);  // (1/(a0 a1), 1/(a2 a3))
        t02 = __fmul2_rn(__fmul2_rn(d02, a13), q);          // d0 a1 / (a0 a1), d2 a3 / (a2 a3)
        t13 = __fmul2_rn(__fmul2_rn(d13, a02), q);
    }
};

// Sixteen tanh in natural order: t[i] = tanh(x_i + b_i) from the pre-activations v[i] and bk[i] = b_i * 2log2(e) - 1
// (16-byte aligned shared array).
__device__ __forceinline__ void tanh16(const float (&v)[16], const float *bk, float (&t)[16]) {
    TanhQuad Q[4];
#pragma unroll
    for (int q = 0; q < 4; ++q)
        Q[q].front(make_float2(v[4 * q], v[4 * q + 1]), make_float2(v[4 * q + 2], v[4 * q + 3]),
                   *reinterpret_cast<const float4 *>(bk + 4 * q));
#pragma unroll
    for (int q = 0; q < 4; ++q) Q[q].exps();
#pragma unroll
    for (int q = 0; q < 4; ++q) Q[q].recip();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float2 t02, t13;
        Q[q].finish(t02, t13);
        t[4 * q] = t02.x; t[4 * q + 2] = t02.y; t[4 * q + 1] = t13.x; t[4 * q + 3] = t13.y;
    }
}

}  // namespace tc

#ifndef DYNAX_NODE_TC_HELPERS_ONLY  // node_vjp_tc.cuh reuses the tc:: helpers above without this kernel

// ------------------------------------------------------------------------------------------------------------
// Forward kernel, templated on the hidden width HID in {32, 64, 128}.
//
// Warp roles: warps 0-3 = 128 row threads (one trajectory = one TMEM lane each), warp 4 = TMA producer (u in,
// x / y / stage points out), warp 5 = MMA issuer (one elected lane).
//
// TMEM (4 HID columns):  A_hi | A_lo | D1 | D2.  Layer 1 accumulates into D1, layer 2 into D2, so the layer-2
// MMAs of a 16-column K slice are issued AS SOON AS the row threads have stored that slice of h1 = tanh(D1 + b1)
// (one mbarrier per slice): the tensor pipe runs under the tanh phase of the same tile and only the last slices'
// MMAs are exposed.  D1 is still being read while D2 is written -- hence two accumulators.
//
// tanh phase (MUFU-bound: 16 MUFU lanes per SM): four tanh share ONE reciprocal,
//     tanh(x) = d/a,  a = 2^(2x log2e - 1) + 1/2,  d = 2^(2x log2e - 1) - 1/2;   r = 1/(a0 a1 a2 a3),  1/a0 = r (a2 a3) a1, ...
// = 5 MUFU per 4 tanh (d is exact, so small |x| keep their RELATIVE accuracy up to ex2.approx's own error; 1 - 1/a
// would cancel), and all FP32 arithmetic runs as packed pairs (FFMA2/FADD2/FMUL2), which leaves the tanh of a quad in
// the order (0,2),(1,3): the K order of the A operand is permuted accordingly inside every group of four columns, and
// the staged copies of W2 (K index) and W3 follow that permutation (free: staged once per CTA).
// ------------------------------------------------------------------------------------------------------------
namespace tcf {
using tc::TN; using tc::NX; using tc::NU; using tc::NIN; using tc::S; using tc::SBO; using tc::K2LOG2E; using tc::TanhQuad;
__host__ __device__ constexpr int perm4(int k) { return (k & ~3) | (((k & 1) << 1) | ((k >> 1) & 1)); }  // 0,2,1,3 (involution)

template <int HID>
struct Cfg {
    static_assert(HID == 32 || HID == 64 || HID == 128, "hidden width");
    static constexpr int NTHR = TN + 64;              // row threads + producer warp + MMA warp
    static constexpr int NSL = HID / 16;              // 16-column slices of a hidden layer (processed in pairs)
    static constexpr uint32_t TMEM_COLS = 4 * HID, COL_AHI = 0, COL_ALO = HID, COL_D1 = 2 * HID, COL_D2 = 3 * HID;
    // kind::tf32, D=F32, A=B=TF32, both K-major, N=HID, M=128
    static constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(HID >> 3) << 17) | ((128u >> 4) << 24);
    static constexpr uint32_t LBO = 16 * HID;  // bytes between 16-byte K chunks: HID/8 row groups of 128 B
    static constexpr int F_W2HI = 0, F_W2LO = F_W2HI + HID * HID, F_W1HI = F_W2LO + HID * HID, F_W1LO = F_W1HI + HID * NIN,
                         F_B1K = F_W1LO + HID * NIN, F_B2K = F_B1K + HID, F_W3P = F_B2K + HID, F_B3 = F_W3P + NX * HID,
                         F_IN = F_B3 + 4, F_XO = F_IN + S * TN * NU, F_YO = F_XO + 2 * TN * NX, F_ZO = F_YO + 2 * TN,
                         F_END = F_ZO + 2 * TN * 3 * NX;
    static constexpr int NBAR = 2 * S + 4 + 3 + NSL;
    static constexpr size_t SMEM_USED = sizeof(float) * F_END + sizeof(uint64_t) * NBAR + 16;
    // 512 TMEM columns per SM: pad the request so that no more CTAs become resident than can allocate (one more
    // would only block inside tcgen05.alloc while pinning registers and shared memory)
    static constexpr size_t SMEM_MIN = HID == 32 ? 46 * 1024 : (HID == 64 ? 100 * 1024 : 0);
    static constexpr size_t SMEM_BYTES = SMEM_USED > SMEM_MIN ? SMEM_USED : SMEM_MIN;
    static __device__ __forceinline__ int wofs(int n, int k) { return (k >> 2) * (4 * HID) + (n >> 3) * 32 + (n & 7) * 4 + (k & 3); }
    static __device__ __forceinline__ uint64_t desc(const void *p) {
        const uint64_t addr = (uint64_t)(ptx::smem_u32(p) >> 4) & 0x3fffu;
        return addr | ((uint64_t)(LBO >> 4) << 16) | ((uint64_t)(SBO >> 4) << 32) | (1ull << 46);
    }
};

__device__ __forceinline__ void mma_i(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// tcgen05.ld of 16 columns WITHOUT the wait: the caller overlaps the load with arithmetic on the previous slice and
// calls ld_fence() on the registers before their first use.
__device__ __forceinline__ void tmem_ld16_async(uint32_t addr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(addr)
        : "memory");
}
__device__ __forceinline__ void ld_fence(uint32_t (&r)[16]) {  // the operands tie every later use of r[] to the wait
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
}

}  // namespace tcf

template <int HID>
__global__ void __launch_bounds__((tcf::Cfg<HID>::NTHR), HID == 128 ? 1 : 2) node_tc_kernel(const NodeTcParams P) {
    using namespace tcf;
    using C = Cfg<HID>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *sm = reinterpret_cast<float *>(smem_raw);
    float *s_in = sm + C::F_IN, *s_xo = sm + C::F_XO, *s_yo = sm + C::F_YO, *s_zo = sm + C::F_ZO;
    uint64_t *full = reinterpret_cast<uint64_t *>(sm + C::F_END);
    uint64_t *empty = full + S, *ofull = empty + S, *ofree = ofull + 2, *z_bar = ofree + 2, *mma1_bar = z_bar + 1,
             *mma2_bar = mma1_bar + 1, *slice_bar = mma2_bar + 1;
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(slice_bar + C::NSL);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t N = P.N, T = P.T;
    const int64_t n0 = (int64_t)blockIdx.x * TN;
    const int valid = (int)((N - n0) < TN ? (N - n0) : TN);
    const bool emit_x = P.xsol != nullptr, emit_y = P.ysol != nullptr, emit_z = P.stage_x != nullptr && P.rk4;
    const bool emit = emit_x || emit_y || emit_z;
    const int stages = P.rk4 ? 4 : 1;
    constexpr int ROWWARPS = TN / 32;

    // ---- one-time setup: barriers, weights split hi/lo into the canonical UMMA layout, TMEM ----
    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            ptx::mbar_init(&full[s], 1);
            ptx::mbar_init(&empty[s], TN / 32);
        }
        ptx::mbar_init(&ofull[0], TN / 32); ptx::mbar_init(&ofull[1], TN / 32);
        ptx::mbar_init(&ofree[0], 1); ptx::mbar_init(&ofree[1], 1);
        ptx::mbar_init(z_bar, TN / 32);
        ptx::mbar_init(mma1_bar, 1);
        ptx::mbar_init(mma2_bar, 1);
        for (int s = 0; s < C::NSL; ++s) ptx::mbar_init(&slice_bar[s], TN / 32);
        ptx::fence_mbar_init();
    }
    for (int i = tid; i < HID * HID; i += blockDim.x) {
        const int nrow = i / HID, k = i % HID;  // W2[nrow][k]; A column perm4(k) holds h1[k]
        uint32_t hi, lo;
        tc::split_tf32(__ldg(P.W2 + i), hi, lo);
        sm[C::F_W2HI + C::wofs(nrow, perm4(k))] = __uint_as_float(hi);
        sm[C::F_W2LO + C::wofs(nrow, perm4(k))] = __uint_as_float(lo);
    }
    for (int i = tid; i < HID * NIN; i += blockDim.x) {
        const int nrow = i / NIN, k = i % NIN;
        uint32_t hi, lo;
        tc::split_tf32(__ldg(P.W1 + i), hi, lo);
        sm[C::F_W1HI + C::wofs(nrow, k)] = __uint_as_float(hi);
        sm[C::F_W1LO + C::wofs(nrow, k)] = __uint_as_float(lo);
    }
    for (int i = tid; i < HID; i += blockDim.x) {
        sm[C::F_B1K + i] = fmaf(__ldg(P.b1 + i), K2LOG2E, -1.0f);
        sm[C::F_B2K + i] = fmaf(__ldg(P.b2 + i), K2LOG2E, -1.0f);
    }
    for (int i = tid; i < NX * HID; i += blockDim.x) {  // [quad][d][position]: position p of a quad holds column perm4(p)
        const int d = i / HID, k = i % HID;
        sm[C::F_W3P + (k >> 2) * (4 * NX) + d * 4 + (perm4(k) & 3)] = __ldg(P.W3 + i);
    }
    if (tid < NX) sm[C::F_B3 + tid] = __ldg(P.b3 + tid);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(s_tmem)),
                     "r"(C::TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    ptx::fence_proxy_async_smem();  // weights were written through the generic proxy; UMMA reads them via the async proxy
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    const uint32_t tmem = *s_tmem;

    if (warp == ROWWARPS) {
        // ------------------------------- producer warp (same protocol as rollout_fwd_kernel) -----------
        const uint64_t pol = ptx::policy_evict_first();
        auto load_step = [&](int64_t c) {
            const int s = (int)(c % S);
            float *dst = s_in + s * TN * NU;
            if (P.shared_u) {  // one row of NU floats for every trajectory
                if (lane < NU) dst[lane] = __ldg(P.u + c * NU + lane);
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&full[s]);
                return;
            }
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

    if (warp == ROWWARPS + 1) {
        // ------------------------------- MMA warp: one elected lane issues every tcgen05.mma -----------
        if (lane == 0) {
            const uint64_t dW1hi = C::desc(sm + C::F_W1HI), dW1lo = C::desc(sm + C::F_W1LO);
            const uint64_t dW2hi = C::desc(sm + C::F_W2HI), dW2lo = C::desc(sm + C::F_W2LO);
            const uint32_t aHi = tmem + C::COL_AHI, aLo = tmem + C::COL_ALO, d1 = tmem + C::COL_D1, d2 = tmem + C::COL_D2;
            uint32_t ph = 0;
            for (int64_t e = 0; e < T * stages; ++e) {
                // layer 1:  D1 = z W1^T   (K = 8, three split products)
                ptx::mbar_wait(z_bar, ph);
                tc::fence_after();
                mma_i(d1, aHi, dW1hi, C::IDESC, 0u);
                mma_i(d1, aLo, dW1hi, C::IDESC, 1u);
                mma_i(d1, aHi, dW1lo, C::IDESC, 1u);
                tc::mma_commit(mma1_bar);
                // layer 2:  D2 = h1 W2^T, slice by slice as the row threads publish h1
#pragma unroll
                for (int sl = 0; sl < C::NSL; ++sl) {
                    ptx::mbar_wait(&slice_bar[sl], ph);
                    tc::fence_after();
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk) {
                        const int k8 = 2 * sl + kk;  // K step of 8 columns; B start address advances 2 K chunks
                        const uint64_t adv = (uint64_t)((2 * C::LBO * k8) >> 4);
                        mma_i(d2, aHi + 8 * k8, dW2hi + adv, C::IDESC, k8 ? 1u : 0u);
                        mma_i(d2, aLo + 8 * k8, dW2hi + adv, C::IDESC, 1u);
                        mma_i(d2, aHi + 8 * k8, dW2lo + adv, C::IDESC, 1u);
                    }
                }
                tc::mma_commit(mma2_bar);
                ph ^= 1u;
            }
        }
        return;
    }

    // --------------------------------- row threads: one trajectory each ---------------------------------
    const int row = tid;
    const bool active = row < valid;
    const int64_t i = active ? (n0 + row) : (N - 1);
    const float dt = P.dt;
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);  // this warp's 32 TMEM lanes
    uint32_t ph = 0;  // parity of z_bar / slice_bar / mma1_bar / mma2_bar: each completes once per evaluation

    // One 16-column slice of a tanh phase.  `cur` holds the pre-activations (tcgen05.ld already waited for), `mid()` runs
    // after the first eight exponentials are queued (the caller publishes the PREVIOUS slice and prefetches the next one
    // there: both only wait for operations issued long before), `out(q, t02, t13)` consumes the tanh of quad q.
    auto slice = [&](const uint32_t (&cur)[16], const float *bias, auto &&mid, auto &&out) {
        TanhQuad Q[4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
            Q[q].front(make_float2(__uint_as_float(cur[4 * q]), __uint_as_float(cur[4 * q + 1])),
                       make_float2(__uint_as_float(cur[4 * q + 2]), __uint_as_float(cur[4 * q + 3])),
                       *reinterpret_cast<const float4 *>(bias + 4 * q));
        Q[0].exps();
        Q[1].exps();
        mid();
        Q[0].recip();
        Q[2].exps();
        Q[1].recip();
        Q[3].exps();
        Q[2].recip();
        Q[3].recip();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float2 t02, t13;
            Q[q].finish(t02, t13);
            out(q, t02, t13);
        }
    };
    auto publish = [&](int sl) {  // slice sl of h1 is in TMEM: let the MMA warp issue its K steps
        tc::wait_st();
        tc::fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&slice_bar[sl]);
    };

    // h1 = tanh(D1 + b1) -> A (hi/lo), slice by slice.  Slice j is published in the middle of slice j+1's arithmetic,
    // the next slice's tcgen05.ld is issued there too.
    auto phase1 = [&]() {
        uint32_t hi[16], lo[16], va[16], vb[16];
        ptx::mbar_wait(mma1_bar, ph);
        tc::fence_after();
        tmem_ld16_async(lane_base + C::COL_D1, va);
        ld_fence(va);
#pragma unroll
        for (int sl = 0; sl < C::NSL; ++sl) {
            uint32_t(&cur)[16] = (sl & 1) ? vb : va;
            uint32_t(&nxt)[16] = (sl & 1) ? va : vb;
            slice(cur, sm + C::F_B1K + 16 * sl,
                  [&]() {
                      if (sl > 0) publish(sl - 1);
                      if (sl + 1 < C::NSL) tmem_ld16_async(lane_base + C::COL_D1 + 16 * (sl + 1), nxt);
                  },
                  [&](int q, float2 t02, float2 t13) {
                      hi[4 * q + 0] = (__float_as_uint(t02.x) + 0x1000u) & 0xffffe000u;  // round to nearest (see split_tf32)
                      hi[4 * q + 1] = (__float_as_uint(t02.y) + 0x1000u) & 0xffffe000u;
                      hi[4 * q + 2] = (__float_as_uint(t13.x) + 0x1000u) & 0xffffe000u;
                      hi[4 * q + 3] = (__float_as_uint(t13.y) + 0x1000u) & 0xffffe000u;
                      const float2 l02 = __fadd2_rn(t02, make_float2(-__uint_as_float(hi[4 * q + 0]), -__uint_as_float(hi[4 * q + 1])));
                      const float2 l13 = __fadd2_rn(t13, make_float2(-__uint_as_float(hi[4 * q + 2]), -__uint_as_float(hi[4 * q + 3])));
                      lo[4 * q + 0] = __float_as_uint(l02.x);
                      lo[4 * q + 1] = __float_as_uint(l02.y);
                      lo[4 * q + 2] = __float_as_uint(l13.x);
                      lo[4 * q + 3] = __float_as_uint(l13.y);
                  });
            tc::tmem_st16(lane_base + C::COL_AHI + 16 * sl, hi);
            tc::tmem_st16(lane_base + C::COL_ALO + 16 * sl, lo);
            if (sl + 1 < C::NSL) ld_fence(nxt);
        }
        publish(C::NSL - 1);
    };
    // h2 = tanh(D2 + b2) and W3 h2 (N = 3: FMA pipe; two partial sums per output as one packed pair)
    auto phase2 = [&](float *part) {
        uint32_t va[16], vb[16];
        ptx::mbar_wait(mma2_bar, ph);
        tc::fence_after();
        float2 r2[NX];
#pragma unroll
        for (int d = 0; d < NX; ++d) r2[d] = make_float2(0.f, 0.f);
        tmem_ld16_async(lane_base + C::COL_D2, va);
        ld_fence(va);
#pragma unroll
        for (int sl = 0; sl < C::NSL; ++sl) {
            uint32_t(&cur)[16] = (sl & 1) ? vb : va;
            uint32_t(&nxt)[16] = (sl & 1) ? va : vb;
            slice(cur, sm + C::F_B2K + 16 * sl,
                  [&]() {
                      if (sl + 1 < C::NSL) tmem_ld16_async(lane_base + C::COL_D2 + 16 * (sl + 1), nxt);
                  },
                  [&](int q, float2 t02, float2 t13) {
                      const float4 *w3 = reinterpret_cast<const float4 *>(sm + C::F_W3P + (4 * sl + q) * (4 * NX));
#pragma unroll
                      for (int d = 0; d < NX; ++d) {
                          const float4 w = w3[d];  // (W3[d][c], W3[d][c+2], W3[d][c+1], W3[d][c+3])
                          r2[d] = __ffma2_rn(t02, make_float2(w.x, w.y), r2[d]);
                          r2[d] = __ffma2_rn(t13, make_float2(w.z, w.w), r2[d]);
                      }
                  });
            if (sl + 1 < C::NSL) ld_fence(nxt);
        }
#pragma unroll
        for (int d = 0; d < NX; ++d) part[d] = r2[d].x + r2[d].y;
    };

    {
        float x[NX];
#pragma unroll
        for (int d = 0; d < NX; ++d) x[d] = __ldg(P.x0 + i * NX + d);
        // one MLP evaluation for the CTA's 128 rows
        auto mlp = [&](const float *z, float *r) {
            uint32_t hi[NIN], lo[NIN];
            // ---- layer 1 operand: z split hi/lo -> A columns 0..7
#pragma unroll
            for (int c = 0; c < NIN; ++c) tc::split_tf32(z[c], hi[c], lo[c]);
            tc::tmem_st8(lane_base + C::COL_AHI, hi);
            tc::tmem_st8(lane_base + C::COL_ALO, lo);
            tc::wait_st();
            tc::fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(z_bar);
            float part[NX];
            phase1();
            phase2(part);
#pragma unroll
            for (int d = 0; d < NX; ++d) r[d] = sm[C::F_B3 + d] + part[d];
            ph ^= 1u;
        };

        const float hs = P.rk4 ? dt / 6.0f : dt;
        const int ustride = P.shared_u ? 0 : NU;
        for (int64_t c = 0; c < T; ++c) {
            const int s = (int)(c % S), o = (int)(c & 1);
            ptx::mbar_wait(&full[s], (uint32_t)((c / S) & 1));
            if (emit) ptx::mbar_wait(&ofree[o], (uint32_t)(((c >> 1) & 1) ^ 1));
            float z[NIN], k[NX], acc[NX];
#pragma unroll
            for (int d = 0; d < NU; ++d) z[NX + d] = s_in[s * TN * NU + row * ustride + d];
            if (emit_y) s_yo[o * TN + row] = x[0];  // y_k = x_k[0], pre-update
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
                    for (int d = 0; d < NX; ++d) s_zo[o * TN * 3 * NX + row * 3 * NX + st * NX + d] = z[d];
                }
            }
#pragma unroll
            for (int d = 0; d < NX; ++d) x[d] = fmaf(hs, acc[d], x[d]);
            if (emit_x) {
#pragma unroll
                for (int d = 0; d < NX; ++d) s_xo[o * TN * NX + row * NX + d] = x[d];
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
    }
    // ---- teardown: every row thread is done with TMEM (and has seen the last MMA complete) before warp 0 frees it
    tc::fence_before();
    tc::rows_sync();
    if (warp == 0) {
        tc::fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(C::TMEM_COLS) : "memory");
    }
}
#endif  // DYNAX_NODE_TC_HELPERS_ONLY

}  // namespace dynax
