// Reverse-mode gradient of the MLP-dynamics RK4 rollout on the 5th-gen tensor cores (row A7, config 5).
//
// Same exact discrete adjoint as node_vjp.cu (see the derivation in its header); what changes is where
// the work runs.  A CTA owns 128 trajectories = TMEM lanes = the M dimension of a UMMA tile, and every
// contraction whose M is the batch goes to tcgen05 with the 3xTF32 split of node_tc.cuh:
//
//   forward at the stage point   D = z W1^T        ->  h1 = tanh(D + b1)                 (K=8,  N=64)
//                                D = h1 W2^T       ->  h2 = tanh(D + b2)                 (K=64, N=64)
//   backward (data path)         delta2 = (W3^T d3) o (1 - h2^2)                         (FMA pipe, 3 x 64)
//                                D = delta2 W2     ->  delta1 = D o (1 - h1^2)           (B = W2^T copy in smem)
//                                D = delta1 W1     ->  zbar                              (B = W1^T copy, N=16)
//
// The weight gradients dW = sum_rows delta^T h contract over the BATCH (K = rows); their operands would have
// to be staged through shared memory as hi/lo pairs in UMMA layout (544 shared stores per row and stage), so
// they stay on the FMA pipe as fp32 tile GEMMs over [feature][row] tiles (every thread owns a 4x8 patch of
// dW2) -- and run INSIDE the windows in which the row threads would otherwise wait for an MMA to complete:
//   window 3 (delta2 W2 in flight):  dW3, db3, db2, dW2 part 1          window 4 (delta1 W1): dW2 part 2
//   windows 1 and 2 of the NEXT stage or forward evaluation: dW2 part 3; dW1, db1 (Z tile double-buffered)
// fp32 register accumulators, flushed to fp64 global accumulators every 16 steps (as node_vjp.cu).
// Stage points are recomputed from the saved forward trajectory with three forward evaluations per step.
// One CTA per SM (225 KB of shared memory: weights + transposed copies, hi and lo, 79 KB; tiles 146 KB), so the CTA
// runs TWO threads per trajectory (256 threads; thread t and t+128 share row t & 127 and TMEM lane quadrant
// (t/32) % 4): each owns 32 of the 64 hidden columns in the tanh / split / tile phases and 64 of the 128 rows in
// the weight-gradient contractions, which gives every SM sub-partition two warps to interleave.
// PARITY UNPINNED BY THE REFERENCE; pinned to torch-autograd fp64 and to the CUDA-core kernel by the tests.
#pragma once
#define DYNAX_NODE_TC_HELPERS_ONLY
#include "node_tc.cuh"

namespace dynax {

struct NodeVjpTcParams {
    const float *W1, *b1, *W2, *b2, *W3, *b3;
    const float *x0, *u, *xsol, *g_xsol, *g_ysol;
    const float *stage_x;  // [T,N,9] stage points saved by the forward kernel, or NULL (recomputed: 3 more evaluations per step)
    float *g_x0, *g_u;
    double *acc;  // fp64 accumulators, layout of node_vjp.cu (A_W1 .. A_B3), zeroed by the caller
    int64_t N, T;
    float dt;
    int rk4;
    int shared_u;  // u is ONE series [T,5] for every trajectory; g_u is then [T,5], summed over trajectories (zeroed by the caller)
};

namespace vtc {
using namespace tc;

constexpr int LD = 132;  // padded tile row: conflict-free float4 rows
constexpr int NTHR = 2 * TN, HC = HID / 2;  // two threads per row, HC hidden columns each
constexpr uint32_t IDESC16 = (1u << 4) | (2u << 7) | (2u << 10) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
constexpr uint32_t LBO16 = 256;  // N = 16: two 8-row groups of 128 B per 16-byte K chunk
__device__ __forceinline__ int wofs16(int n, int k) { return (k >> 2) * 64 + (n >> 3) * 32 + (n & 7) * 4 + (k & 3); }
__device__ __forceinline__ uint64_t smem_desc16(const void *p) {
    const uint64_t addr = (uint64_t)(ptx::smem_u32(p) >> 4) & 0x3fffu;
    return addr | ((uint64_t)(LBO16 >> 4) << 16) | ((uint64_t)(SBO >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma_ts_i(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// shared-memory map (floats)
constexpr int G_W2HI = 0, G_W2LO = G_W2HI + HID * HID, G_W2THI = G_W2LO + HID * HID, G_W2TLO = G_W2THI + HID * HID,
              G_W1HI = G_W2TLO + HID * HID, G_W1LO = G_W1HI + HID * NIN, G_W1THI = G_W1LO + HID * NIN,
              G_W1TLO = G_W1THI + 16 * HID, G_B1 = G_W1TLO + 16 * HID, G_B2 = G_B1 + HID, G_W3 = G_B2 + HID,
              G_B3 = G_W3 + NX * HID, T_H1 = G_B3 + 4, T_H2 = T_H1 + HID * LD, T_D2 = T_H2 + HID * LD,
              T_D1 = T_D2 + HID * LD, T_Z = T_D1 + HID * LD, T_D3 = T_Z + 2 * NIN * LD, T_R = T_D3 + 4 * LD,
              G_END = T_R + 2 * 4 * TN;  // T_R: the two column halves' partial sums of the 64 -> 3 output layer
constexpr size_t SMEM_BYTES = sizeof(float) * G_END + sizeof(uint64_t) + 16;
// fp64 accumulator layout (identical to nv:: in node_vjp.cu)
constexpr int A_W1 = 0, A_B1 = A_W1 + HID * NIN, A_W2 = A_B1 + HID, A_B2 = A_W2 + HID * HID, A_W3 = A_B2 + HID,
              A_B3 = A_W3 + NX * HID, A_END = A_B3 + NX;
constexpr int FLUSH = 16;
}  // namespace vtc

__global__ void __launch_bounds__(vtc::NTHR, 1) node_vjp_tc_kernel(const NodeVjpTcParams P) {
    using namespace vtc;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *sm = reinterpret_cast<float *>(smem_raw);
    float *H1 = sm + T_H1, *H2 = sm + T_H2, *D2 = sm + T_D2, *D1 = sm + T_D1, *Zt = sm + T_Z, *D3 = sm + T_D3, *R = sm + T_R;
    uint64_t *mma_bar = reinterpret_cast<uint64_t *>(sm + G_END);
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(mma_bar + 1);

    const int tid = threadIdx.x, warp = tid >> 5;
    const int row = tid & (TN - 1), half = tid >> 7;  // this thread's trajectory / column half (and row half)
    const int c_lo = half * HC;                       // hidden columns c_lo .. c_lo + HC - 1
    const int r_lo = half * (TN / 2);                 // tile rows r_lo .. r_lo + TN/2 - 1 in the contractions
    const int64_t N = P.N, T = P.T;
    const int64_t i = (int64_t)blockIdx.x * TN + row;
    const bool active = i < N;
    const int64_t iu = active ? i : 0;

    // ---- one-time setup: barrier, weights (and their transposes) split hi/lo into the canonical UMMA layout, TMEM
    if (tid == 0) {
        ptx::mbar_init(mma_bar, 1);
        ptx::fence_mbar_init();
    }
    for (int e = tid; e < HID * HID; e += NTHR) {
        const int j = e / HID, c = e % HID;  // W2[j][c]: output j, input c
        uint32_t hi, lo;
        split_tf32(__ldg(P.W2 + e), hi, lo);
        sm[G_W2HI + wofs(j, c)] = __uint_as_float(hi);   // forward:  B[n = j][k = c]
        sm[G_W2LO + wofs(j, c)] = __uint_as_float(lo);
        sm[G_W2THI + wofs(c, j)] = __uint_as_float(hi);  // backward: B[n = c][k = j]
        sm[G_W2TLO + wofs(c, j)] = __uint_as_float(lo);
    }
    for (int e = tid; e < 16 * HID; e += NTHR) {  // zero the padded rows 8..15 of the W1^T tile
        sm[G_W1THI + e] = 0.f;
        sm[G_W1TLO + e] = 0.f;
    }
    __syncthreads();
    for (int e = tid; e < HID * NIN; e += NTHR) {
        const int a = e / NIN, c = e % NIN;  // W1[a][c]: hidden a, input c
        uint32_t hi, lo;
        split_tf32(__ldg(P.W1 + e), hi, lo);
        sm[G_W1HI + wofs(a, c)] = __uint_as_float(hi);
        sm[G_W1LO + wofs(a, c)] = __uint_as_float(lo);
        sm[G_W1THI + wofs16(c, a)] = __uint_as_float(hi);
        sm[G_W1TLO + wofs16(c, a)] = __uint_as_float(lo);
    }
    for (int e = tid; e < HID; e += NTHR) {
        sm[G_B1 + e] = fmaf(__ldg(P.b1 + e), K2LOG2E, -1.0f);  // pre-scaled for tc::tanh16
        sm[G_B2 + e] = fmaf(__ldg(P.b2 + e), K2LOG2E, -1.0f);
    }
    for (int e = tid; e < NX * HID; e += NTHR) sm[G_W3 + e] = __ldg(P.W3 + e);
    if (tid < NX) sm[G_B3 + tid] = __ldg(P.b3 + tid);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(s_tmem)),
                     "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    ptx::fence_proxy_async_smem();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *s_tmem;
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);  // a warp reaches TMEM lanes 32 (w % 4) ..
    const uint64_t dW1hi = smem_desc(sm + G_W1HI), dW1lo = smem_desc(sm + G_W1LO);
    const uint64_t dW2hi = smem_desc(sm + G_W2HI), dW2lo = smem_desc(sm + G_W2LO);
    const uint64_t dW2Thi = smem_desc(sm + G_W2THI), dW2Tlo = smem_desc(sm + G_W2TLO);
    const uint64_t dW1Thi = smem_desc16(sm + G_W1THI), dW1Tlo = smem_desc16(sm + G_W1TLO);
    uint32_t ph = 0;
    const float dt = P.dt;
    const bool rk4 = P.rk4 != 0;

    // ---- this thread's share of the weight gradients (fp32, flushed to fp64 every FLUSH steps)
    // (dW2: two partial sums per entry -- even / odd tile rows -- so the inner product runs as packed FFMA2)
    float2 aW2[4][8];
    float aW1[4], aW3[2], aB = 0.f, aB3 = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        aW1[a] = 0.f;
#pragma unroll
        for (int b = 0; b < 8; ++b) aW2[a][b] = make_float2(0.f, 0.f);
    }
    aW3[0] = aW3[1] = 0.f;
    // dW2 patch: rows j0..j0+3, cols i0 + 8b (interleaved: the 8 lanes of a quarter-warp then read 8 CONSECUTIVE tile
    // rows, which LD = 132 spreads over all 32 banks; 8 rows apart they would all hit the same banks)
    const int j0 = (row >> 3) * 4, i0 = row & 7;
    const int w1i = row >> 1, w1c = (row & 1) * 4;      // dW1[w1i][w1c..w1c+3]
    const int w3d = row >> 5, w3j = row & 31;           // dW3[w3d][w3j], dW3[w3d][w3j + 32]   (row < 96)

    auto flush = [&]() {
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                atomicAdd(P.acc + A_W2 + (j0 + a) * HID + i0 + 8 * b, (double)aW2[a][b].x + (double)aW2[a][b].y);
                aW2[a][b] = make_float2(0.f, 0.f);
            }
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            atomicAdd(P.acc + A_W1 + w1i * NIN + w1c + a, (double)aW1[a]);
            aW1[a] = 0.f;
        }
        if (row < 96) {
            atomicAdd(P.acc + A_W3 + w3d * HID + w3j, (double)aW3[0]);
            atomicAdd(P.acc + A_W3 + w3d * HID + w3j + 32, (double)aW3[1]);
        }
        aW3[0] = aW3[1] = 0.f;
        atomicAdd(P.acc + (row < 64 ? A_B2 + row : A_B1 + row - 64), (double)aB);
        aB = 0.f;
        if (row < NX) atomicAdd(P.acc + A_B3 + row, (double)aB3);
        aB3 = 0.f;
    };

    // ---- tile contractions (run inside the MMA windows); this thread covers tile rows r_lo .. r_lo + 63
    auto contract_w2 = [&](int r_begin, int r_end) {
        for (int r = r_begin; r < r_end; r += 4) {
            float4 dj[4], hi[8];
#pragma unroll
            for (int a = 0; a < 4; ++a) dj[a] = *reinterpret_cast<const float4 *>(D2 + (j0 + a) * LD + r);
#pragma unroll
            for (int b = 0; b < 8; ++b) hi[b] = *reinterpret_cast<const float4 *>(H1 + (i0 + 8 * b) * LD + r);
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    aW2[a][b] = __ffma2_rn(make_float2(dj[a].x, dj[a].y), make_float2(hi[b].x, hi[b].y), aW2[a][b]);
                    aW2[a][b] = __ffma2_rn(make_float2(dj[a].z, dj[a].w), make_float2(hi[b].z, hi[b].w), aW2[a][b]);
                }
        }
    };
    auto contract_w3_b3_b2 = [&]() {
        for (int r = r_lo; r < r_lo + TN / 2; r += 4) {
            if (row < 96) {
                const float4 d3r = *reinterpret_cast<const float4 *>(D3 + w3d * LD + r);
#pragma unroll
                for (int a = 0; a < 2; ++a) {
                    const float4 hr = *reinterpret_cast<const float4 *>(H2 + (w3j + 32 * a) * LD + r);
                    aW3[a] = fmaf(d3r.x, hr.x, aW3[a]); aW3[a] = fmaf(d3r.y, hr.y, aW3[a]);
                    aW3[a] = fmaf(d3r.z, hr.z, aW3[a]); aW3[a] = fmaf(d3r.w, hr.w, aW3[a]);
                }
            }
            if (row < 64) {
                const float4 br = *reinterpret_cast<const float4 *>(D2 + row * LD + r);
                aB += (br.x + br.y) + (br.z + br.w);
            }
            if (row < NX) {
                const float4 b3 = *reinterpret_cast<const float4 *>(D3 + row * LD + r);
                aB3 += (b3.x + b3.y) + (b3.z + b3.w);
            }
        }
    };
    auto contract_w1_b1 = [&](int zbuf) {
        const float *Z = Zt + zbuf * NIN * LD;
        for (int r = r_lo; r < r_lo + TN / 2; r += 4) {
            const float4 d1r = *reinterpret_cast<const float4 *>(D1 + w1i * LD + r);
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const float4 zr = *reinterpret_cast<const float4 *>(Z + (w1c + a) * LD + r);
                aW1[a] = fmaf(d1r.x, zr.x, aW1[a]); aW1[a] = fmaf(d1r.y, zr.y, aW1[a]);
                aW1[a] = fmaf(d1r.z, zr.z, aW1[a]); aW1[a] = fmaf(d1r.w, zr.w, aW1[a]);
            }
            if (row >= 64) {
                const float4 br = *reinterpret_cast<const float4 *>(D1 + (row - 64) * LD + r);
                aB += (br.x + br.y) + (br.z + br.w);
            }
        }
    };

    // publish the A operand the row threads just stored and let one thread issue the MMAs
    auto a_ready = [&]() {
        wait_st();
        fence_before();
        __syncthreads();
    };
    auto mma_wait = [&]() {
        ptx::mbar_wait(mma_bar, ph);
        ph ^= 1u;
        fence_after();
    };
    auto issue_l1 = [&]() {  // D = z W1^T  (K = 8)
        if (tid == 0) {
            fence_after();
            mma_ts(tmem + COL_D, tmem + COL_AHI, dW1hi, 0u);
            mma_ts(tmem + COL_D, tmem + COL_ALO, dW1hi, 1u);
            mma_ts(tmem + COL_D, tmem + COL_AHI, dW1lo, 1u);
            mma_commit(mma_bar);
        }
    };
    auto issue_k64 = [&](uint64_t bhi, uint64_t blo, uint32_t idesc, uint32_t lbo) {  // D = A B^T, K = 64
        if (tid == 0) {
            fence_after();
#pragma unroll
            for (int s = 0; s < HID / 8; ++s)
                mma_ts_i(tmem + COL_D, tmem + COL_AHI + 8 * s, bhi + (uint64_t)((2 * lbo * s) >> 4), idesc, s ? 1u : 0u);
#pragma unroll
            for (int s = 0; s < HID / 8; ++s)
                mma_ts_i(tmem + COL_D, tmem + COL_ALO + 8 * s, bhi + (uint64_t)((2 * lbo * s) >> 4), idesc, 1u);
#pragma unroll
            for (int s = 0; s < HID / 8; ++s)
                mma_ts_i(tmem + COL_D, tmem + COL_AHI + 8 * s, blo + (uint64_t)((2 * lbo * s) >> 4), idesc, 1u);
            mma_commit(mma_bar);
        }
    };
    auto store_z = [&](const float *z) {  // the 8 input columns: written by the row's first thread
        if (half == 0) {
            uint32_t hi[NIN], lo[NIN];
#pragma unroll
            for (int c = 0; c < NIN; ++c) split_tf32(z[c], hi[c], lo[c]);
            tmem_st8(lane_base + COL_AHI, hi);
            tmem_st8(lane_base + COL_ALO, lo);
        }
    };

    // Weight-gradient work left over from the previous backward stage, drained in the next windows:
    //   dW2 iterations W2_SPLIT..15 (tiles H1, D2 stay intact until the next backward stage's h1 phase) -> next window 1
    //   dW1 / db1               (tiles D1, Z[zbuf^1] stay intact until the next stage's delta1 phase) -> next window 2
    // A forward evaluation touches no tile, so its two windows serve as well.
    constexpr int ITERS = TN / 2 / 4, W2_A = 6, W2_B = 11;  // dW2 iterations: [0,W2_A) window 3, [W2_A,W2_B) window 4, rest later
    int zbuf = 0;             // Z tile buffer the NEXT backward stage writes
    bool pending_w2 = false, pending_w1 = false;
    auto drain_w2 = [&]() {
        if (pending_w2) contract_w2(r_lo + 4 * W2_B, r_lo + 4 * ITERS);
        pending_w2 = false;
    };
    auto drain_w1 = [&]() {
        if (pending_w1) contract_w1_b1(zbuf ^ 1);
        pending_w1 = false;
    };

    // forward evaluation only (stage points): r = MLP(z).  Both threads of a row return the same r.
    auto mlp_forward = [&](const float *z, float *r) {
        uint32_t hi[16], lo[16];
        store_z(z);
        a_ready();
        issue_l1();
        drain_w2();  // window 1
        mma_wait();
#pragma unroll
        for (int cc = 0; cc < HC; cc += 16) {
            const int c = c_lo + cc;
            float v[16], t[16];
            tmem_ld16(lane_base + COL_D + c, v);
            tanh16(v, sm + G_B1 + c, t);
#pragma unroll
            for (int j = 0; j < 16; ++j) split_tf32(t[j], hi[j], lo[j]);
            tmem_st16(lane_base + COL_AHI + c, hi);
            tmem_st16(lane_base + COL_ALO + c, lo);
        }
        a_ready();
        issue_k64(dW2hi, dW2lo, IDESC, LBO);
        drain_w1();  // window 2
        mma_wait();
        float pr[NX];
#pragma unroll
        for (int d = 0; d < NX; ++d) pr[d] = 0.f;
#pragma unroll
        for (int cc = 0; cc < HC; cc += 16) {
            const int c = c_lo + cc;
            float v[16], t[16];
            tmem_ld16(lane_base + COL_D + c, v);
            tanh16(v, sm + G_B2 + c, t);
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
#pragma unroll
                for (int d = 0; d < NX; ++d) {
                    const float4 w = *reinterpret_cast<const float4 *>(sm + G_W3 + d * HID + c + j);
                    pr[d] = fmaf(w.x, t[j], pr[d]); pr[d] = fmaf(w.y, t[j + 1], pr[d]);
                    pr[d] = fmaf(w.z, t[j + 2], pr[d]); pr[d] = fmaf(w.w, t[j + 3], pr[d]);
                }
            }
        }
        // exchange the two halves' partial sums of the output layer (same order in both threads)
#pragma unroll
        for (int d = 0; d < NX; ++d) R[(half * 4 + d) * TN + row] = pr[d];
        __syncthreads();
#pragma unroll
        for (int d = 0; d < NX; ++d) r[d] = sm[G_B3 + d] + (R[d * TN + row] + R[(4 + d) * TN + row]);
    };

    // backward pass of the MLP at stage point z with output cotangent d3 -> zbar (8 values, both threads of a row)
    auto mlp_backward = [&](const float *z, const float *d3, float *zb) {
        uint32_t hi[16], lo[16];
        float h1[HC];
        float *Z = Zt + zbuf * NIN * LD;
        if (half == 0) {
#pragma unroll
            for (int c = 0; c < NIN; ++c) Z[c * LD + row] = z[c];
#pragma unroll
            for (int d = 0; d < NX; ++d) D3[d * LD + row] = d3[d];
        }
        store_z(z);
        a_ready();
        issue_l1();
        drain_w2();  // window 1
        mma_wait();
        // h1 = tanh(D + b1): kept in registers for delta1, tile for dW2, A operand for layer 2
#pragma unroll
        for (int cc = 0; cc < HC; cc += 16) {
            const int c = c_lo + cc;
            float v[16], t[16];
            tmem_ld16(lane_base + COL_D + c, v);
            tanh16(v, sm + G_B1 + c, t);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                h1[cc + j] = t[j];
                H1[(c + j) * LD + row] = t[j];
                split_tf32(t[j], hi[j], lo[j]);
            }
            tmem_st16(lane_base + COL_AHI + c, hi);
            tmem_st16(lane_base + COL_ALO + c, lo);
        }
        a_ready();
        issue_k64(dW2hi, dW2lo, IDESC, LBO);
        drain_w1();  // window 2
        mma_wait();
        // h2 = tanh(D + b2);  delta2 = (W3^T d3) o (1 - h2^2)  -> tiles, A operand
#pragma unroll
        for (int cc = 0; cc < HC; cc += 16) {
            const int c = c_lo + cc;
            float v[16], t[16];
            tmem_ld16(lane_base + COL_D + c, v);
            tanh16(v, sm + G_B2 + c, t);
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
                float g[4] = {0.f, 0.f, 0.f, 0.f};   // (W3^T d3)_(c+j..c+j+3)
#pragma unroll
                for (int d = 0; d < NX; ++d) {
                    const float4 w = *reinterpret_cast<const float4 *>(sm + G_W3 + d * HID + c + j);
                    g[0] = fmaf(w.x, d3[d], g[0]); g[1] = fmaf(w.y, d3[d], g[1]);
                    g[2] = fmaf(w.z, d3[d], g[2]); g[3] = fmaf(w.w, d3[d], g[3]);
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float tq = t[j + q], e = g[q] * (1.f - tq * tq);
                    H2[(c + j + q) * LD + row] = tq;
                    D2[(c + j + q) * LD + row] = e;
                    split_tf32(e, hi[j + q], lo[j + q]);
                }
            }
            tmem_st16(lane_base + COL_AHI + c, hi);
            tmem_st16(lane_base + COL_ALO + c, lo);
        }
        a_ready();
        issue_k64(dW2Thi, dW2Tlo, IDESC, LBO);  // D = delta2 W2
        contract_w3_b3_b2();                    // window 3
        contract_w2(r_lo, r_lo + 4 * W2_A);
        mma_wait();
        // delta1 = D o (1 - h1^2) -> tile, A operand
#pragma unroll
        for (int cc = 0; cc < HC; cc += 16) {
            const int c = c_lo + cc;
            float v[16];
            tmem_ld16(lane_base + COL_D + c, v);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float e = v[j] * (1.f - h1[cc + j] * h1[cc + j]);
                D1[(c + j) * LD + row] = e;
                split_tf32(e, hi[j], lo[j]);
            }
            tmem_st16(lane_base + COL_AHI + c, hi);
            tmem_st16(lane_base + COL_ALO + c, lo);
        }
        a_ready();
        issue_k64(dW1Thi, dW1Tlo, IDESC16, LBO16);  // D = delta1 W1  (16 columns, 8 used)
        contract_w2(r_lo + 4 * W2_A, r_lo + 4 * W2_B);  // window 4
        mma_wait();
        {
            float v[16];
            tmem_ld16(lane_base + COL_D, v);
#pragma unroll
            for (int c = 0; c < NIN; ++c) zb[c] = v[c];
        }
        pending_w2 = pending_w1 = true;
        zbuf ^= 1;
    };

    float lam[NX] = {0.f, 0.f, 0.f};
    const int stages = rk4 ? 4 : 1;
    for (int64_t k = T - 1; k >= 0; --k) {
        float x[NX], u[NU], zx[4][NX], kk[NX];
#pragma unroll
        for (int d = 0; d < NX; ++d) {
            x[d] = active ? (k == 0 ? __ldg(P.x0 + iu * NX + d) : __ldg(P.xsol + ((k - 1) * N + iu) * NX + d)) : 0.f;
            if (active && P.g_xsol) lam[d] += __ldg(P.g_xsol + (k * N + iu) * NX + d);  // cotangent of xsol[k] = x_{k+1}
        }
#pragma unroll
        for (int d = 0; d < NU; ++d) u[d] = active ? __ldg(P.u + (P.shared_u ? k : k * N + iu) * NU + d) : 0.f;
        // stage points (x parts; the u part is the same for all four)
#pragma unroll
        for (int d = 0; d < NX; ++d) zx[0][d] = x[d];
        if (rk4 && P.stage_x != nullptr) {
#pragma unroll
            for (int s = 0; s < 3; ++s)
#pragma unroll
                for (int d = 0; d < NX; ++d) zx[s + 1][d] = active ? __ldg(P.stage_x + (k * N + iu) * 3 * NX + s * NX + d) : 0.f;
        } else if (rk4) {
#pragma unroll 1
            for (int s = 0; s < 3; ++s) {
                float z[NIN];
#pragma unroll
                for (int d = 0; d < NX; ++d) z[d] = zx[s][d];
#pragma unroll
                for (int d = 0; d < NU; ++d) z[NX + d] = u[d];
                mlp_forward(z, kk);
                const float cs = (s == 2) ? dt : 0.5f * dt;
#pragma unroll
                for (int d = 0; d < NX; ++d) zx[s + 1][d] = fmaf(cs, kk[d], x[d]);
            }
        }
        // reverse through the stages
        float xbar[NX] = {0.f, 0.f, 0.f}, ubar[NU] = {0.f, 0.f, 0.f, 0.f, 0.f}, carry[NX] = {0.f, 0.f, 0.f};
#pragma unroll 1
        for (int s = stages - 1; s >= 0; --s) {
            const float ws = rk4 ? ((s == 0 || s == 3) ? dt / 6.0f : dt / 3.0f) : dt;
            const float cn = (s == 2) ? dt : 0.5f * dt;  // coefficient of k_s inside z_{s+1}
            float d3[NX], zb[NIN], z[NIN];
#pragma unroll
            for (int d = 0; d < NX; ++d) {
                d3[d] = (s == stages - 1) ? ws * lam[d] : fmaf(cn, carry[d], ws * lam[d]);
                if (!active) d3[d] = 0.f;
                z[d] = zx[s][d];
            }
#pragma unroll
            for (int d = 0; d < NU; ++d) z[NX + d] = u[d];
            mlp_backward(z, d3, zb);
#pragma unroll
            for (int d = 0; d < NX; ++d) {
                carry[d] = zb[d];
                xbar[d] += zb[d];
            }
#pragma unroll
            for (int d = 0; d < NU; ++d) ubar[d] += zb[NX + d];
        }
#pragma unroll
        for (int d = 0; d < NX; ++d) lam[d] += xbar[d];
        if (active && P.g_ysol) lam[0] += __ldg(P.g_ysol + k * N + iu);  // y_k = x_k[0]
        if (P.g_u && half == 0) {
            if (P.shared_u) {  // one row per step: sum over this warp's trajectories, one atomic per warp and channel
#pragma unroll
                for (int d = 0; d < NU; ++d) {
                    float v = active ? ubar[d] : 0.f;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    if ((tid & 31) == 0) atomicAdd(P.g_u + k * NU + d, v);
                }
            } else if (active) {
#pragma unroll
                for (int d = 0; d < NU; ++d) P.g_u[(k * N + iu) * NU + d] = ubar[d];
            }
        }
        if (((T - 1 - k) % FLUSH) == FLUSH - 1) flush();  // (pending tile work lands in a later flush)
    }
    __syncthreads();  // the last stage's tiles are complete
    drain_w2();
    drain_w1();
    flush();
    if (active && P.g_x0 && half == 0) {
#pragma unroll
        for (int d = 0; d < NX; ++d) P.g_x0[iu * NX + d] = lam[d];
    }
    fence_before();
    __syncthreads();
    if (warp == 0) {
        fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
    }
}

}  // namespace dynax
