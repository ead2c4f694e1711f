// Reverse-mode gradient of the MLP-dynamics RK4 rollout (row A7, what jax.grad would derive through
// nn.scan over the general fx/fy path, /root/reference/dynax/core/base_block_state_space.py:62-63).
// CUDA-core implementation: exact discrete adjoint of classical RK4, one CTA = 128 trajectories.
//
// Per step, last to first (lambda = dL/dx_{k+1}):
//   recompute the stage points  z1 = x_k, z2 = x_k + dt/2 k1, z3 = x_k + dt/2 k2, z4 = x_k + dt k3
//   kbar4 = dt/6 lambda                       zbar4 = J(z4)^T kbar4
//   kbar3 = dt/3 lambda + dt   zbar4[x]       zbar3 = J(z3)^T kbar3
//   kbar2 = dt/3 lambda + dt/2 zbar3[x]       zbar2 = J(z2)^T kbar2
//   kbar1 = dt/6 lambda + dt/2 zbar2[x]       zbar1 = J(z1)^T kbar1
//   lambda <- lambda + sum_s zbar_s[x] (+ output cotangent),   g_u[k] = sum_s zbar_s[u]
// Each J^T product is a backward pass through the MLP at a recomputed stage point; the weight
// gradients dW = sum over rows of delta^T h are batch contractions done as shared-memory tile GEMMs
// (every thread owns a 4x8 patch of dW2), accumulated in registers and flushed to fp64 global
// accumulators every 16 steps.  x_k comes from the forward trajectory (xsol), so nothing is re-integrated.
// PARITY UNPINNED BY THE REFERENCE; the test oracle is torch autograd (fp64) over the restated rollout.
#include <string.h>

#include <stdlib.h>

#include "api_common.cuh"
#include "mlp_model.cuh"
#include "node_vjp_tc.cuh"
#include "node_vjp_wide.cuh"

namespace dynax {

namespace nv {
constexpr int TN = 128, NX = 3, NU = 5, NIN = 8, LD = 132;  // LD: padded tile row (bank-conflict-free float4 rows)
constexpr int FLUSH = 16;
// Per hidden width (32 or 64 here; the tcgen05 kernel of node_vjp_tc.cuh is the hidden = 64 product path):
template <int HID>
struct L {
    static_assert(HID == 32 || HID == 64, "tile GEMM thread mappings below");
    using M = MlpModel<3, 5, 1, HID>;
    static constexpr int O_TILES = (M::SMEM_FLOATS + 3) & ~3;
    static constexpr int O_H1 = O_TILES, O_H2 = O_H1 + HID * LD, O_D2 = O_H2 + HID * LD, O_D1 = O_D2 + HID * LD,
                         O_Z = O_D1 + HID * LD, O_D3 = O_Z + NIN * LD, O_END = O_D3 + 4 * LD;
    static constexpr size_t SMEM_BYTES = sizeof(float) * O_END;
    // fp64 accumulator layout
    static constexpr int A_W1 = 0, A_B1 = A_W1 + HID * NIN, A_W2 = A_B1 + HID, A_B2 = A_W2 + HID * HID, A_W3 = A_B2 + HID,
                         A_B3 = A_W3 + NX * HID, A_END = A_B3 + NX;
    // thread mappings of the batch contractions (128 threads):
    static constexpr int PJ = HID / 16, PI = HID / 8;  // dW2: 16 x 8 threads, PJ rows x PI interleaved columns each
    static constexpr int W1PT = HID / 16;              // dW1: HID*8 entries, W1PT consecutive ones of a row each
    static constexpr int W3PT = HID / 32;              // dW3: 3 x 32 threads, W3PT columns (32 apart) each
};
constexpr int A_END_MAX = nvw::A_END;  // the widest instantiated hidden layer (128: node_vjp_wide.cuh)
static_assert(nvw::A_END >= L<64>::A_END, "workspace covers every hidden width");
}  // namespace nv

template <int HID>
struct NodeVjpParams {
    typename nv::L<HID>::M::Ptrs mp;
    const float *x0, *u, *xsol, *g_xsol, *g_ysol;
    const float *stage_x;  // [T,N,9] stage points saved by the forward kernel, or NULL (recomputed)
    float *g_x0, *g_u;
    double *acc;
    int64_t N, T;
    float dt;
    int shared_u;  // u is ONE series [T,5]; g_u is then [T,5], summed over trajectories (zeroed by the caller)
};

template <int HID>
__global__ void __launch_bounds__(nv::TN, 1) node_vjp_kernel(const NodeVjpParams<HID> P) {
    using namespace nv;
    using Ly = L<HID>;
    using M = typename Ly::M;
    constexpr int O_H1 = Ly::O_H1, O_H2 = Ly::O_H2, O_D2 = Ly::O_D2, O_D1 = Ly::O_D1, O_Z = Ly::O_Z, O_D3 = Ly::O_D3;
    constexpr int A_W1 = Ly::A_W1, A_B1 = Ly::A_B1, A_W2 = Ly::A_W2, A_B2 = Ly::A_B2, A_W3 = Ly::A_W3, A_B3 = Ly::A_B3;
    constexpr int PJ = Ly::PJ, PI = Ly::PI, W1PT = Ly::W1PT, W3PT = Ly::W3PT;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *sm = reinterpret_cast<float *>(smem_raw);
    float *H1 = sm + O_H1, *H2 = sm + O_H2, *D2 = sm + O_D2, *D1 = sm + O_D1, *Z = sm + O_Z, *D3 = sm + O_D3;
    const int tid = threadIdx.x;
    const int64_t N = P.N, T = P.T;
    const int64_t i = (int64_t)blockIdx.x * TN + tid;
    const bool active = i < N;
    M::block_init(sm, P.mp, tid, TN);
    __syncthreads();
    M model;
    model.load(P.mp, 0, sm);
    const float *w = sm;
    const float dt = P.dt;
    const bool rk4 = P.mp.rk4 != 0;

    // register accumulators (this thread's share of every weight gradient)
    float aW2[PJ][PI], aW1[W1PT], aW3[W3PT], aB = 0.f, aB3 = 0.f;
#pragma unroll
    for (int a = 0; a < PJ; ++a)
#pragma unroll
        for (int b = 0; b < PI; ++b) aW2[a][b] = 0.f;
#pragma unroll
    for (int a = 0; a < W1PT; ++a) aW1[a] = 0.f;
#pragma unroll
    for (int a = 0; a < W3PT; ++a) aW3[a] = 0.f;
    // dW2 patch: rows j0..j0+PJ-1, cols i0 + 8b (interleaved: the 8 lanes of a quarter-warp then read 8 CONSECUTIVE tile
    // rows, which LD = 132 spreads over all 32 banks; 8 rows apart they would all hit the same banks)
    const int j0 = (tid >> 3) * PJ, i0 = tid & 7;
    const int w1i = (tid * W1PT) / NIN, w1c = (tid * W1PT) % NIN;  // dW1[w1i][w1c..w1c+W1PT-1]
    const int w3d = tid >> 5, w3j = tid & 31;                      // dW3[w3d][w3j + 32a], a < W3PT   (tid < 96)
    const bool has_b = tid < 2 * HID;                              // db2[tid] (tid < HID) or db1[tid - HID]

    auto flush = [&]() {
#pragma unroll
        for (int a = 0; a < PJ; ++a)
#pragma unroll
            for (int b = 0; b < PI; ++b) {
                atomicAdd(P.acc + A_W2 + (j0 + a) * HID + i0 + 8 * b, (double)aW2[a][b]);
                aW2[a][b] = 0.f;
            }
#pragma unroll
        for (int a = 0; a < W1PT; ++a) {
            atomicAdd(P.acc + A_W1 + w1i * NIN + w1c + a, (double)aW1[a]);
            aW1[a] = 0.f;
        }
        if (tid < 96) {
#pragma unroll
            for (int a = 0; a < W3PT; ++a) atomicAdd(P.acc + A_W3 + w3d * HID + w3j + 32 * a, (double)aW3[a]);
        }
#pragma unroll
        for (int a = 0; a < W3PT; ++a) aW3[a] = 0.f;
        if (has_b) atomicAdd(P.acc + (tid < HID ? A_B2 + tid : A_B1 + tid - HID), (double)aB);
        aB = 0.f;
        if (tid < NX) atomicAdd(P.acc + A_B3 + tid, (double)aB3);
        aB3 = 0.f;
    };

    // Backward pass of the MLP at stage point z with output cotangent d3 -> zbar; adds this stage's
    // contribution to the weight-gradient accumulators (collective over the CTA).
    auto mlp_backward = [&](const float *z, const float *d3, float *zbar) {
        float h1[HID], acc1[HID];
#pragma unroll
        for (int a = 0; a < HID; ++a) {
            float s = w[M::O_B1 + a];
#pragma unroll
            for (int c = 0; c < NIN; c += 4) {
                const float4 q = *reinterpret_cast<const float4 *>(w + M::O_W1 + a * NIN + c);
                s = fmaf(q.x, z[c], s); s = fmaf(q.y, z[c + 1], s); s = fmaf(q.z, z[c + 2], s); s = fmaf(q.w, z[c + 3], s);
            }
            h1[a] = tanhf(s);
            H1[a * LD + tid] = h1[a];
            acc1[a] = 0.f;
        }
#pragma unroll
        for (int c = 0; c < NIN; ++c) Z[c * LD + tid] = z[c];
#pragma unroll
        for (int d = 0; d < NX; ++d) D3[d * LD + tid] = d3[d];
#pragma unroll 2
        for (int j = 0; j < HID; ++j) {
            const float *row = w + M::O_W2 + j * HID;
            float4 q[HID / 4];
            float a0 = w[M::O_B2 + j], a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
            for (int c = 0; c < HID / 4; ++c) {
                q[c] = *reinterpret_cast<const float4 *>(row + 4 * c);
                a0 = fmaf(q[c].x, h1[4 * c], a0); a1 = fmaf(q[c].y, h1[4 * c + 1], a1);
                a2 = fmaf(q[c].z, h1[4 * c + 2], a2); a3 = fmaf(q[c].w, h1[4 * c + 3], a3);
            }
            const float h2 = tanhf((a0 + a1) + (a2 + a3));
            float g = 0.f;  // (W3^T d3)_j
#pragma unroll
            for (int d = 0; d < NX; ++d) g = fmaf(w[M::O_W3 + d * HID + j], d3[d], g);
            const float d2 = g * (1.f - h2 * h2);
            H2[j * LD + tid] = h2;
            D2[j * LD + tid] = d2;
#pragma unroll
            for (int c = 0; c < HID / 4; ++c) {  // acc1 += W2[j,:] * d2   (-> W2^T delta2)
                acc1[4 * c] = fmaf(q[c].x, d2, acc1[4 * c]); acc1[4 * c + 1] = fmaf(q[c].y, d2, acc1[4 * c + 1]);
                acc1[4 * c + 2] = fmaf(q[c].z, d2, acc1[4 * c + 2]); acc1[4 * c + 3] = fmaf(q[c].w, d2, acc1[4 * c + 3]);
            }
        }
#pragma unroll
        for (int c = 0; c < NIN; ++c) zbar[c] = 0.f;
#pragma unroll
        for (int a = 0; a < HID; ++a) {
            const float d1 = acc1[a] * (1.f - h1[a] * h1[a]);
            D1[a * LD + tid] = d1;
#pragma unroll
            for (int c = 0; c < NIN; c += 4) {
                const float4 q = *reinterpret_cast<const float4 *>(w + M::O_W1 + a * NIN + c);
                zbar[c] = fmaf(q.x, d1, zbar[c]); zbar[c + 1] = fmaf(q.y, d1, zbar[c + 1]);
                zbar[c + 2] = fmaf(q.z, d1, zbar[c + 2]); zbar[c + 3] = fmaf(q.w, d1, zbar[c + 3]);
            }
        }
        __syncthreads();
        // ---- batch contractions over the CTA's 128 rows
        for (int r = 0; r < TN; r += 4) {
            float4 dj[PJ], hi[PI];
#pragma unroll
            for (int a = 0; a < PJ; ++a) dj[a] = *reinterpret_cast<const float4 *>(D2 + (j0 + a) * LD + r);
#pragma unroll
            for (int b = 0; b < PI; ++b) hi[b] = *reinterpret_cast<const float4 *>(H1 + (i0 + 8 * b) * LD + r);
#pragma unroll
            for (int a = 0; a < PJ; ++a)
#pragma unroll
                for (int b = 0; b < PI; ++b) {
                    aW2[a][b] = fmaf(dj[a].x, hi[b].x, aW2[a][b]); aW2[a][b] = fmaf(dj[a].y, hi[b].y, aW2[a][b]);
                    aW2[a][b] = fmaf(dj[a].z, hi[b].z, aW2[a][b]); aW2[a][b] = fmaf(dj[a].w, hi[b].w, aW2[a][b]);
                }
            const float4 d1r = *reinterpret_cast<const float4 *>(D1 + w1i * LD + r);
#pragma unroll
            for (int a = 0; a < W1PT; ++a) {
                const float4 zr = *reinterpret_cast<const float4 *>(Z + (w1c + a) * LD + r);
                aW1[a] = fmaf(d1r.x, zr.x, aW1[a]); aW1[a] = fmaf(d1r.y, zr.y, aW1[a]);
                aW1[a] = fmaf(d1r.z, zr.z, aW1[a]); aW1[a] = fmaf(d1r.w, zr.w, aW1[a]);
            }
            if (tid < 96) {
                const float4 d3r = *reinterpret_cast<const float4 *>(D3 + w3d * LD + r);
#pragma unroll
                for (int a = 0; a < W3PT; ++a) {
                    const float4 hr = *reinterpret_cast<const float4 *>(H2 + (w3j + 32 * a) * LD + r);
                    aW3[a] = fmaf(d3r.x, hr.x, aW3[a]); aW3[a] = fmaf(d3r.y, hr.y, aW3[a]);
                    aW3[a] = fmaf(d3r.z, hr.z, aW3[a]); aW3[a] = fmaf(d3r.w, hr.w, aW3[a]);
                }
            }
            if (has_b) {
                const float4 br = *reinterpret_cast<const float4 *>((tid < HID ? D2 + tid * LD : D1 + (tid - HID) * LD) + r);
                aB += (br.x + br.y) + (br.z + br.w);
            }
            if (tid < NX) {
                const float4 b3 = *reinterpret_cast<const float4 *>(D3 + tid * LD + r);
                aB3 += (b3.x + b3.y) + (b3.z + b3.w);
            }
        }
        __syncthreads();
    };

    float lam[NX] = {0.f, 0.f, 0.f};
    const int64_t iu = active ? i : 0;
    for (int64_t k = T - 1; k >= 0; --k) {
        float x[NX], u[NU], z[4][NIN], kk[NX];
#pragma unroll
        for (int d = 0; d < NX; ++d) {
            x[d] = active ? (k == 0 ? P.x0[iu * NX + d] : P.xsol[((k - 1) * N + iu) * NX + d]) : 0.f;
            if (active && P.g_xsol) lam[d] += P.g_xsol[(k * N + iu) * NX + d];   // cotangent of xsol[k] = x_{k+1}
        }
#pragma unroll
        for (int d = 0; d < NU; ++d) u[d] = active ? P.u[(P.shared_u ? k : k * N + iu) * NU + d] : 0.f;
        const int stages = rk4 ? 4 : 1;
        // stage points
#pragma unroll
        for (int s = 0; s < 4; ++s)
#pragma unroll
            for (int d = 0; d < NU; ++d) z[s][NX + d] = u[d];
#pragma unroll
        for (int d = 0; d < NX; ++d) z[0][d] = x[d];
        if (rk4 && P.stage_x != nullptr) {
#pragma unroll
            for (int s = 0; s < 3; ++s)
#pragma unroll
                for (int d = 0; d < NX; ++d) z[s + 1][d] = active ? P.stage_x[(k * N + iu) * 3 * NX + s * NX + d] : 0.f;
        } else if (rk4) {
#pragma unroll 1
            for (int s = 0; s < 3; ++s) {
                model.rhs(z[s], u, kk);
                const float cs = (s == 2) ? dt : 0.5f * dt;
#pragma unroll
                for (int d = 0; d < NX; ++d) z[s + 1][d] = fmaf(cs, kk[d], x[d]);
            }
        }
        // reverse through the stages
        float xbar[NX] = {0.f, 0.f, 0.f}, ubar[NU] = {0.f, 0.f, 0.f, 0.f, 0.f}, carry[NX] = {0.f, 0.f, 0.f};
#pragma unroll 1
        for (int s = stages - 1; s >= 0; --s) {
            // kbar_s = w_s * lambda + c_{s+1} * zbar_{s+1}[x]
            const float ws = rk4 ? ((s == 0 || s == 3) ? dt / 6.0f : dt / 3.0f) : dt;
            const float cn = (s == 2) ? dt : 0.5f * dt;   // coefficient of k_s inside z_{s+1}
            float d3[NX], zb[NIN];
#pragma unroll
            for (int d = 0; d < NX; ++d) d3[d] = active ? fmaf(cn, carry[d], ws * lam[d]) : 0.f;
            if (s == stages - 1) {
#pragma unroll
                for (int d = 0; d < NX; ++d) d3[d] = active ? ws * lam[d] : 0.f;
            }
            mlp_backward(z[s], d3, zb);
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
        if (active && P.g_ysol) lam[0] += P.g_ysol[k * N + iu];   // y_k = x_k[0]
        if (P.g_u) {
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
        if (((T - 1 - k) % FLUSH) == FLUSH - 1) flush();
    }
    flush();
    if (active && P.g_x0) {
#pragma unroll
        for (int d = 0; d < NX; ++d) P.g_x0[iu * NX + d] = lam[d];
    }
}

__global__ void node_vjp_finalize_kernel(const double *acc, int hid, float *gW1, float *gb1, float *gW2, float *gb2,
                                         float *gW3, float *gb3) {
    using namespace nv;
    const int A_B1 = hid * NIN, A_W2 = A_B1 + hid, A_B2 = A_W2 + hid * hid, A_W3 = A_B2 + hid, A_B3 = A_W3 + NX * hid,
              A_END = A_B3 + NX;  // the layout of L<hid>
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= A_END) return;
    const float v = (float)acc[t];
    if (t < A_B1) { if (gW1) gW1[t] = v; }
    else if (t < A_W2) { if (gb1) gb1[t - A_B1] = v; }
    else if (t < A_B2) { if (gW2) gW2[t - A_W2] = v; }
    else if (t < A_W3) { if (gb2) gb2[t - A_B2] = v; }
    else if (t < A_B3) { if (gW3) gW3[t - A_W3] = v; }
    else { if (gb3) gb3[t - A_B3] = v; }
}

template <int HID>
static int launch_node_vjp_cc(const float *W1, const float *b1, const float *W2, const float *b2, const float *W3,
                              const float *b3, const float *x0, const float *u, const float *xsol, const float *stage_x,
                              const float *g_xsol, const float *g_ysol, float *g_x0, float *g_u, double *acc, int64_t N,
                              int64_t T, float dt, int rk4, int shared_u, cudaStream_t st) {
    NodeVjpParams<HID> P;
    P.mp.W1 = W1; P.mp.b1 = b1; P.mp.W2 = W2; P.mp.b2 = b2; P.mp.W3 = W3; P.mp.b3 = b3; P.mp.rk4 = rk4;
    P.x0 = x0; P.u = u; P.xsol = xsol; P.stage_x = stage_x; P.g_xsol = g_xsol; P.g_ysol = g_ysol; P.g_x0 = g_x0; P.g_u = g_u;
    P.acc = acc; P.N = N; P.T = T; P.dt = dt; P.shared_u = shared_u;
    int rc = set_smem(node_vjp_kernel<HID>, nv::L<HID>::SMEM_BYTES);
    if (rc != DYNAX_OK) return rc;
    node_vjp_kernel<HID><<<(unsigned)((N + nv::TN - 1) / nv::TN), nv::TN, nv::L<HID>::SMEM_BYTES, st>>>(P);
    DYNAX_CUDA_TRY(cudaGetLastError());
    return DYNAX_OK;
}

}  // namespace dynax

using namespace dynax;

extern "C" {

size_t dynax_node_rk4_vjp_workspace(void) { return sizeof(double) * ((nv::A_END_MAX + 31) & ~31); }

int dynax_node_rk4_vjp_f32(const float *W1, const float *b1, const float *W2, const float *b2, const float *W3,
                           const float *b3, const float *x0, const float *u, const float *xsol, const float *g_xsol,
                           const float *g_ysol, float *gW1, float *gb1, float *gW2, float *gb2, float *gW3, float *gb3,
                           float *g_x0, float *g_u, int64_t N, int64_t T, int n, int m, int p, int hidden, float dt,
                           uint32_t flags, void *ws, size_t ws_bytes, void *stream) {
    DYNAX_API_RANGE();
    return dynax_node_rk4_vjp_stages_f32(W1, b1, W2, b2, W3, b3, x0, u, xsol, nullptr, g_xsol, g_ysol, gW1, gb1, gW2, gb2,
                                         gW3, gb3, g_x0, g_u, N, T, n, m, p, hidden, dt, flags, ws, ws_bytes, stream);
}

int dynax_node_rk4_vjp_stages_f32(const float *W1, const float *b1, const float *W2, const float *b2, const float *W3,
                                  const float *b3, const float *x0, const float *u, const float *xsol,
                                  const float *stage_x, const float *g_xsol, const float *g_ysol, float *gW1, float *gb1,
                                  float *gW2, float *gb2, float *gW3, float *gb3, float *g_x0, float *g_u, int64_t N,
                                  int64_t T, int n, int m, int p, int hidden, float dt, uint32_t flags, void *ws,
                                  size_t ws_bytes, void *stream) {
    DYNAX_API_RANGE();
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (N < 0 || T < 1) return fail(DYNAX_ERR_INVALID_ARGUMENT, "need N >= 0 and T >= 1");
    if (N == 0) return DYNAX_OK;
    if (!W1 || !b1 || !W2 || !b2 || !W3 || !b3 || !x0 || !u) return fail(DYNAX_ERR_INVALID_ARGUMENT, "weights, x0, u must be non-NULL");
    if (!xsol && T > 1) return fail(DYNAX_ERR_INVALID_ARGUMENT, "xsol (the forward trajectory) must be given");
    if (!g_xsol && !g_ysol) return fail(DYNAX_ERR_INVALID_ARGUMENT, "need g_xsol or g_ysol");
    if (flags & ~(DYNAX_SOLVER_EULER | DYNAX_SHARED_U)) return fail(DYNAX_ERR_INVALID_ARGUMENT, "unsupported flags 0x%x for node_rk4_vjp", flags);
    if (!(n == 3 && m == 5 && p == 1 && (hidden == 32 || hidden == 64 || hidden == 128)))
        return fail(DYNAX_ERR_UNSUPPORTED_DIMS, "MLP gradient kernels are instantiated for (n,m,p)=(3,5,1), hidden in {32,64,128}; got (%d,%d,%d,%d)", n, m, p, hidden);
    const size_t need = dynax_node_rk4_vjp_workspace();
    if (!ws || ws_bytes < need || !aligned16(ws)) return fail(DYNAX_ERR_WORKSPACE, "workspace must hold %zu bytes (16-byte aligned)", need);
    int rc = require_sm100();
    if (rc != DYNAX_OK) return rc;
    double *acc = reinterpret_cast<double *>(ws);
    const int rk4 = (flags & DYNAX_SOLVER_EULER) ? 0 : 1;
    DYNAX_CUDA_TRY(cudaMemsetAsync(ws, 0, need, st));
    const int shared_u = (flags & DYNAX_SHARED_U) ? 1 : 0;
    if (shared_u && g_u) DYNAX_CUDA_TRY(cudaMemsetAsync(g_u, 0, sizeof(float) * 5 * (size_t)T, st));  // summed with atomics
    static_assert(vtc::A_END == nv::L<64>::A_END, "both kernels share the accumulator layout");
    // hidden = 64: tcgen05 kernel (node_vjp_tc.cuh); DYNAX_NODE_VJP_IMPL=0 and hidden = 32: CUDA-core kernel;
    // hidden = 128: the wide CUDA-core kernel (node_vjp_wide.cuh)
    const char *impl = getenv("DYNAX_NODE_VJP_IMPL");
    if (hidden == 128) {
        NodeVjpWideParams Q;
        Q.mp.W1 = W1; Q.mp.b1 = b1; Q.mp.W2 = W2; Q.mp.b2 = b2; Q.mp.W3 = W3; Q.mp.b3 = b3; Q.mp.rk4 = rk4;
        Q.x0 = x0; Q.u = u; Q.xsol = xsol; Q.stage_x = stage_x; Q.g_xsol = g_xsol; Q.g_ysol = g_ysol; Q.g_x0 = g_x0; Q.g_u = g_u;
        Q.acc = acc; Q.N = N; Q.T = T; Q.dt = dt; Q.shared_u = shared_u;
        rc = set_smem(node_vjp_wide_kernel, nvw::SMEM_BYTES);
        if (rc != DYNAX_OK) return rc;
        node_vjp_wide_kernel<<<(unsigned)((N + nvw::TR - 1) / nvw::TR), nvw::NTHR, nvw::SMEM_BYTES, st>>>(Q);
        DYNAX_CUDA_TRY(cudaGetLastError());
        rc = DYNAX_OK;
    } else if (hidden == 32) {
        rc = launch_node_vjp_cc<32>(W1, b1, W2, b2, W3, b3, x0, u, xsol, stage_x, g_xsol, g_ysol, g_x0, g_u, acc, N, T, dt, rk4, shared_u, st);
    } else if (impl && impl[0] == '0') {
        rc = launch_node_vjp_cc<64>(W1, b1, W2, b2, W3, b3, x0, u, xsol, stage_x, g_xsol, g_ysol, g_x0, g_u, acc, N, T, dt, rk4, shared_u, st);
    } else {
        NodeVjpTcParams Q;
        Q.W1 = W1; Q.b1 = b1; Q.W2 = W2; Q.b2 = b2; Q.W3 = W3; Q.b3 = b3;
        Q.x0 = x0; Q.u = u; Q.xsol = xsol; Q.stage_x = stage_x; Q.g_xsol = g_xsol; Q.g_ysol = g_ysol; Q.g_x0 = g_x0; Q.g_u = g_u;
        Q.acc = acc; Q.N = N; Q.T = T; Q.dt = dt; Q.rk4 = rk4; Q.shared_u = shared_u;
        rc = set_smem(node_vjp_tc_kernel, vtc::SMEM_BYTES);
        if (rc != DYNAX_OK) return rc;
        node_vjp_tc_kernel<<<(unsigned)((N + vtc::TN - 1) / vtc::TN), vtc::NTHR, vtc::SMEM_BYTES, st>>>(Q);
        DYNAX_CUDA_TRY(cudaGetLastError());
        rc = DYNAX_OK;
    }
    if (rc != DYNAX_OK) return rc;
    const int a_end = hidden * 8 + hidden + hidden * hidden + hidden + 3 * hidden + 3;
    node_vjp_finalize_kernel<<<(a_end + 255) / 256, 256, 0, st>>>(acc, hidden, gW1, gb1, gW2, gb2, gW3, gb3);
    DYNAX_CUDA_TRY(cudaGetLastError());
    return DYNAX_OK;
}

}  // extern "C"
