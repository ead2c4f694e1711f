// Reverse-mode gradient of the MLP-dynamics RK4 rollout for the WIDE hidden layer (hidden = 128), CUDA cores.
//
// Same exact discrete adjoint of RK4 as node_vjp.cu (derivation in its header).  At hidden = 128 the four
// [feature][row] tiles of a 128-trajectory CTA alone would be 270 KB, and a thread cannot hold 128 hidden activations
// plus a 128-entry patch of dW2, so the work is laid out differently:
//
//   * a CTA owns 32 trajectories; 8 warps, warp q = hidden units [16 q, 16 q + 16), lane = trajectory.  The eight threads
//     of a trajectory carry the same x / lambda / stage points (cheap) and split every layer eight ways; whatever a layer
//     needs from the whole hidden vector it reads from the shared tiles H1 / D2 ([feature][row], 36-float rows);
//   * the weight gradients are contracted over the tile's 32 rows by all 256 threads (dW2: an 8 x 8 patch per thread)
//     and accumulated in fp32 SHARED-memory accumulators (every entry has one owner: no atomics), flushed to the fp64
//     global accumulators every 16 steps (as node_vjp.cu);
//   * weights are read from shared memory as warp-uniform float4 broadcasts (W2 rows for both directions: the backward
//     pass of layer 2 walks W2[j][16 q .. 16 q + 15] for every j, so no transposed copy is needed).
//
// Shared memory: weights 72 KB + tiles 74 KB + accumulators 72 KB + exchange 10 KB = 224 KB, one CTA per SM.
// tanhf is the accurate one (as the other CUDA-core kernels).  PARITY UNPINNED BY THE REFERENCE; the test oracle is torch
// autograd (fp64) over the restated rollout.
#pragma once
#include "api_common.cuh"
#include "mlp_model.cuh"

namespace dynax {

namespace nvw {
constexpr int HID = 128, TR = 32, NW = 8, NTHR = 32 * NW, HQ = HID / NW;  // rows per CTA, warps, threads, hidden units per warp
constexpr int NX = 3, NU = 5, NIN = 8, LDT = TR + 4;          // tile rows: 32 trajectories + 4 floats (float4-aligned)
constexpr int SW2 = HID + 2;                                  // row pitch of the dW2 accumulator (bank spread of the patch owners)
using M = MlpModel<3, 5, 1, HID>;
constexpr int O_W = 0;
constexpr int O_H1 = (M::SMEM_FLOATS + 3) & ~3, O_H2 = O_H1 + HID * LDT, O_D2 = O_H2 + HID * LDT, O_D1 = O_D2 + HID * LDT,
              O_Z = O_D1 + HID * LDT, O_D3 = O_Z + NIN * LDT, O_ZB = O_D3 + 4 * LDT, O_PR = O_ZB,  // (PR: forward evaluations only)
              O_AW2 = O_ZB + NW * NIN * TR, O_AW1 = O_AW2 + HID * SW2, O_AW3 = O_AW1 + HID * NIN, O_AB1 = O_AW3 + NX * HID,
              O_AB2 = O_AB1 + HID, O_AB3 = O_AB2 + HID, O_END = O_AB3 + 4;
constexpr size_t SMEM_BYTES = sizeof(float) * O_END;
static_assert(SMEM_BYTES <= 227 * 1024, "one CTA per SM");
// fp64 accumulator layout (the layout of nv::L<HID> in node_vjp.cu)
constexpr int A_W1 = 0, A_B1 = A_W1 + HID * NIN, A_W2 = A_B1 + HID, A_B2 = A_W2 + HID * HID, A_W3 = A_B2 + HID,
              A_B3 = A_W3 + NX * HID, A_END = A_B3 + NX;
constexpr int FLUSH = 16;
}  // namespace nvw

struct NodeVjpWideParams {
    nvw::M::Ptrs mp;
    const float *x0, *u, *xsol, *g_xsol, *g_ysol;
    const float *stage_x;  // [T,N,9] stage points saved by the forward kernel, or NULL (recomputed)
    float *g_x0, *g_u;
    double *acc;
    int64_t N, T;
    float dt;
    int shared_u;
};

__global__ void __launch_bounds__(nvw::NTHR, 1) node_vjp_wide_kernel(const NodeVjpWideParams P) {
    using namespace nvw;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *sm = reinterpret_cast<float *>(smem_raw);
    const float *w = sm + O_W;
    float *H1 = sm + O_H1, *H2 = sm + O_H2, *D2 = sm + O_D2, *D1 = sm + O_D1, *Z = sm + O_Z, *D3 = sm + O_D3;
    float *ZB = sm + O_ZB, *PR = sm + O_PR;
    float *aW2 = sm + O_AW2, *aW1 = sm + O_AW1, *aW3 = sm + O_AW3, *aB1 = sm + O_AB1, *aB2 = sm + O_AB2, *aB3 = sm + O_AB3;
    const int tid = threadIdx.x, q = tid >> 5, r = tid & 31;   // warp = its share of the hidden units, lane = trajectory
    const int h0 = q * HQ;
    const int64_t N = P.N, T = P.T;
    const int64_t i = (int64_t)blockIdx.x * TR + r;
    const bool active = i < N;
    const int64_t iu = active ? i : 0;
    M::block_init(sm + O_W, P.mp, tid, NTHR);
    for (int e = tid; e < O_END - O_AW2; e += NTHR) sm[O_AW2 + e] = 0.f;
    __syncthreads();
    const float dt = P.dt;
    const bool rk4 = P.mp.rk4 != 0;

    // ---- layer 1 for this warp's hidden units: h1 = tanh(W1 z + b1) -> registers and the H1 tile
    auto layer1 = [&](const float *z, float *h1) {
#pragma unroll
        for (int jj = 0; jj < HQ; ++jj) {
            const int j = h0 + jj;
            const float4 a = *reinterpret_cast<const float4 *>(w + M::O_W1 + j * NIN);
            const float4 b = *reinterpret_cast<const float4 *>(w + M::O_W1 + j * NIN + 4);
            float s = w[M::O_B1 + j];
            s = fmaf(a.x, z[0], s); s = fmaf(a.y, z[1], s); s = fmaf(a.z, z[2], s); s = fmaf(a.w, z[3], s);
            s = fmaf(b.x, z[4], s); s = fmaf(b.y, z[5], s); s = fmaf(b.z, z[6], s); s = fmaf(b.w, z[7], s);
            h1[jj] = tanhf(s);
            H1[j * LDT + r] = h1[jj];
        }
    };
    // ---- layer 2 pre-activations for this warp's hidden units from the whole H1 tile column of this trajectory
    auto layer2 = [&](float *pre) {
#pragma unroll
        for (int jj = 0; jj < HQ; ++jj) pre[jj] = w[M::O_B2 + h0 + jj];
#pragma unroll 2
        for (int i0 = 0; i0 < HID; i0 += 4) {
            const float v0 = H1[i0 * LDT + r], v1 = H1[(i0 + 1) * LDT + r], v2 = H1[(i0 + 2) * LDT + r], v3 = H1[(i0 + 3) * LDT + r];
#pragma unroll
            for (int jj = 0; jj < HQ; ++jj) {
                const float4 c = *reinterpret_cast<const float4 *>(w + M::O_W2 + (h0 + jj) * HID + i0);
                pre[jj] = fmaf(c.x, v0, pre[jj]); pre[jj] = fmaf(c.y, v1, pre[jj]);
                pre[jj] = fmaf(c.z, v2, pre[jj]); pre[jj] = fmaf(c.w, v3, pre[jj]);
            }
        }
    };

    // forward evaluation only (stage points): rhs = MLP(z).  All threads of a trajectory return the same value.
    auto mlp_forward = [&](const float *z, float *rhs) {
        float h1[HQ], pre[HQ];
        layer1(z, h1);
        __syncthreads();
        layer2(pre);
        float pr[NX] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int jj = 0; jj < HQ; ++jj) {
            const float h2 = tanhf(pre[jj]);
#pragma unroll
            for (int d = 0; d < NX; ++d) pr[d] = fmaf(w[M::O_W3 + d * HID + h0 + jj], h2, pr[d]);
        }
#pragma unroll
        for (int d = 0; d < NX; ++d) PR[(q * 4 + d) * TR + r] = pr[d];
        __syncthreads();
#pragma unroll
        for (int d = 0; d < NX; ++d) {
            float s = 0.f;
#pragma unroll
            for (int qq = 0; qq < NW; ++qq) s += PR[(qq * 4 + d) * TR + r];   // same order in every thread of the trajectory
            rhs[d] = w[M::O_B3 + d] + s;
        }
        __syncthreads();  // PR and H1 are rewritten by the next evaluation
    };

    // backward pass of the MLP at stage point z with output cotangent d3 -> zbar (all threads of a trajectory);
    // adds this stage's contribution to the shared weight-gradient accumulators (collective over the CTA)
    auto mlp_backward = [&](const float *z, const float *d3, float *zbar) {
        float h1[HQ], pre[HQ];
        if (q == 0) {
#pragma unroll
            for (int c = 0; c < NIN; ++c) Z[c * LDT + r] = z[c];
#pragma unroll
            for (int d = 0; d < NX; ++d) D3[d * LDT + r] = d3[d];
        }
        layer1(z, h1);
        __syncthreads();
        layer2(pre);
#pragma unroll
        for (int jj = 0; jj < HQ; ++jj) {
            const int j = h0 + jj;
            const float h2 = tanhf(pre[jj]);
            float g = 0.f;  // (W3^T d3)_j
#pragma unroll
            for (int d = 0; d < NX; ++d) g = fmaf(w[M::O_W3 + d * HID + j], d3[d], g);
            H2[j * LDT + r] = h2;
            D2[j * LDT + r] = g * (1.f - h2 * h2);
        }
        __syncthreads();
        // delta1 for this warp's hidden units: (W2^T delta2)_i (1 - h1_i^2), walking W2[j][h0 .. h0 + HQ - 1] for every j
        float *acc1 = pre;  // (pre is dead)
#pragma unroll
        for (int ii = 0; ii < HQ; ++ii) acc1[ii] = 0.f;
#pragma unroll 2
        for (int j = 0; j < HID; ++j) {
            const float dj = D2[j * LDT + r];
#pragma unroll
            for (int ii = 0; ii < HQ; ii += 4) {
                const float4 c = *reinterpret_cast<const float4 *>(w + M::O_W2 + j * HID + h0 + ii);
                acc1[ii] = fmaf(c.x, dj, acc1[ii]); acc1[ii + 1] = fmaf(c.y, dj, acc1[ii + 1]);
                acc1[ii + 2] = fmaf(c.z, dj, acc1[ii + 2]); acc1[ii + 3] = fmaf(c.w, dj, acc1[ii + 3]);
            }
        }
        float zp[NIN];
#pragma unroll
        for (int c = 0; c < NIN; ++c) zp[c] = 0.f;
#pragma unroll
        for (int ii = 0; ii < HQ; ++ii) {
            const int ih = h0 + ii;
            const float d1 = acc1[ii] * (1.f - h1[ii] * h1[ii]);
            D1[ih * LDT + r] = d1;
            const float4 a = *reinterpret_cast<const float4 *>(w + M::O_W1 + ih * NIN);
            const float4 b = *reinterpret_cast<const float4 *>(w + M::O_W1 + ih * NIN + 4);
            zp[0] = fmaf(a.x, d1, zp[0]); zp[1] = fmaf(a.y, d1, zp[1]); zp[2] = fmaf(a.z, d1, zp[2]); zp[3] = fmaf(a.w, d1, zp[3]);
            zp[4] = fmaf(b.x, d1, zp[4]); zp[5] = fmaf(b.y, d1, zp[5]); zp[6] = fmaf(b.z, d1, zp[6]); zp[7] = fmaf(b.w, d1, zp[7]);
        }
#pragma unroll
        for (int c = 0; c < NIN; ++c) ZB[(q * NIN + c) * TR + r] = zp[c];
        __syncthreads();
#pragma unroll
        for (int c = 0; c < NIN; ++c) {
            float s = 0.f;
#pragma unroll
            for (int qq = 0; qq < NW; ++qq) s += ZB[(qq * NIN + c) * TR + r];   // same order in every thread of the trajectory
            zbar[c] = s;
        }
        // ---- batch contractions over the tile's 32 rows into the shared accumulators (one owner per entry)
        {   // dW2[j][i] += sum_r D2[j][r] H1[i][r]: rows j0 .. j0+7, columns i0 + 16 b (16 x 16 threads, 8 x 8 patches)
            const int j0 = (tid >> 4) * 8, i0 = tid & 15;
            float p[8][8];
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int b = 0; b < 8; ++b) p[a][b] = 0.f;
#pragma unroll 1
            for (int rr = 0; rr < TR; rr += 4) {
                float4 dj[8], hi[8];
#pragma unroll
                for (int a = 0; a < 8; ++a) dj[a] = *reinterpret_cast<const float4 *>(D2 + (j0 + a) * LDT + rr);
#pragma unroll
                for (int b = 0; b < 8; ++b) hi[b] = *reinterpret_cast<const float4 *>(H1 + (i0 + 16 * b) * LDT + rr);
#pragma unroll
                for (int a = 0; a < 8; ++a)
#pragma unroll
                    for (int b = 0; b < 8; ++b) {
                        p[a][b] = fmaf(dj[a].x, hi[b].x, p[a][b]); p[a][b] = fmaf(dj[a].y, hi[b].y, p[a][b]);
                        p[a][b] = fmaf(dj[a].z, hi[b].z, p[a][b]); p[a][b] = fmaf(dj[a].w, hi[b].w, p[a][b]);
                    }
            }
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int b = 0; b < 8; ++b) aW2[(j0 + a) * SW2 + i0 + 16 * b] += p[a][b];
        }
        if (tid < HID) {   // thread t < 128 owns hidden unit t: dW1[t][0..7], dW3[0..2][t], db1[t], db2[t]
            float s1[NIN], s3[NX], sb1 = 0.f, sb2 = 0.f;
#pragma unroll
            for (int c = 0; c < NIN; ++c) s1[c] = 0.f;
#pragma unroll
            for (int d = 0; d < NX; ++d) s3[d] = 0.f;
#pragma unroll 2
            for (int rr = 0; rr < TR; rr += 4) {
                const float4 d1r = *reinterpret_cast<const float4 *>(D1 + tid * LDT + rr);
                const float4 d2r = *reinterpret_cast<const float4 *>(D2 + tid * LDT + rr);
                const float4 h2r = *reinterpret_cast<const float4 *>(H2 + tid * LDT + rr);
                sb1 += (d1r.x + d1r.y) + (d1r.z + d1r.w);
                sb2 += (d2r.x + d2r.y) + (d2r.z + d2r.w);
#pragma unroll
                for (int c = 0; c < NIN; ++c) {
                    const float4 zr = *reinterpret_cast<const float4 *>(Z + c * LDT + rr);
                    s1[c] = fmaf(d1r.x, zr.x, s1[c]); s1[c] = fmaf(d1r.y, zr.y, s1[c]);
                    s1[c] = fmaf(d1r.z, zr.z, s1[c]); s1[c] = fmaf(d1r.w, zr.w, s1[c]);
                }
#pragma unroll
                for (int d = 0; d < NX; ++d) {
                    const float4 dr = *reinterpret_cast<const float4 *>(D3 + d * LDT + rr);
                    s3[d] = fmaf(dr.x, h2r.x, s3[d]); s3[d] = fmaf(dr.y, h2r.y, s3[d]);
                    s3[d] = fmaf(dr.z, h2r.z, s3[d]); s3[d] = fmaf(dr.w, h2r.w, s3[d]);
                }
            }
#pragma unroll
            for (int c = 0; c < NIN; ++c) aW1[tid * NIN + c] += s1[c];
#pragma unroll
            for (int d = 0; d < NX; ++d) aW3[d * HID + tid] += s3[d];
            aB1[tid] += sb1;
            aB2[tid] += sb2;
            if (tid < NX) {
                float s = 0.f;
                for (int rr = 0; rr < TR; ++rr) s += D3[tid * LDT + rr];
                aB3[tid] += s;
            }
        }
        __syncthreads();  // the tiles (and ZB) are rewritten by the next stage
    };

    auto flush = [&]() {  // shared fp32 accumulators -> global fp64 accumulators; called by all threads between stages
        for (int e = tid; e < HID * HID; e += NTHR) {
            const int j = e / HID, c = e % HID;
            atomicAdd(P.acc + A_W2 + e, (double)aW2[j * SW2 + c]);
            aW2[j * SW2 + c] = 0.f;
        }
        for (int e = tid; e < HID * NIN; e += NTHR) { atomicAdd(P.acc + A_W1 + e, (double)aW1[e]); aW1[e] = 0.f; }
        for (int e = tid; e < NX * HID; e += NTHR) { atomicAdd(P.acc + A_W3 + e, (double)aW3[e]); aW3[e] = 0.f; }
        if (tid < HID) {
            atomicAdd(P.acc + A_B1 + tid, (double)aB1[tid]); aB1[tid] = 0.f;
            atomicAdd(P.acc + A_B2 + tid, (double)aB2[tid]); aB2[tid] = 0.f;
        }
        if (tid < NX) { atomicAdd(P.acc + A_B3 + tid, (double)aB3[tid]); aB3[tid] = 0.f; }
        __syncthreads();
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
        // reverse through the stages (see node_vjp.cu)
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
        if (P.g_u && q == 0) {
            if (P.shared_u) {  // one row per step: sum over this warp's trajectories, one atomic per channel
#pragma unroll
                for (int d = 0; d < NU; ++d) {
                    float v = active ? ubar[d] : 0.f;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    if (r == 0) atomicAdd(P.g_u + k * NU + d, v);
                }
            } else if (active) {
#pragma unroll
                for (int d = 0; d < NU; ++d) P.g_u[(k * N + iu) * NU + d] = ubar[d];
            }
        }
        if (((T - 1 - k) % FLUSH) == FLUSH - 1) flush();
    }
    flush();
    if (active && P.g_x0 && q == 0) {
#pragma unroll
        for (int d = 0; d < NX; ++d) P.g_x0[iu * NX + d] = lam[d];
    }
}

}  // namespace dynax
