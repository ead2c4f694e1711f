// Forward-mode (tangent) evaluation of the fused 4R3C loss + gradient: ONE sweep over u and target.
//
// The 4R3C model has only 7 parameters, so instead of the reverse sweep (rollout_adj.cuh) the 3x7
// sensitivity matrix S = dx/dtheta can ride along with the state ("alt." row):
//     x_{k+1} = x_k + dt (A x_k + B u_k)
//     S_{k+1} = S_k + dt (A S_k + F_k),      F_k[:,j] = (dA/dtheta_j) x_k + (dB/dtheta_j) u_k      (9 non-zeros)
//     loss    = mean_k (x_k[0] - target_k)^2,     dloss/dtheta_j = sum_k (2/T)(x_k[0] - target_k) S_k[0,j]
// This is mathematically the same gradient jax.value_and_grad returns (examples/3-inverse.py:96-112); it
// reads u and target ONCE (24 B per system-step instead of the adjoint's 46 B, no checkpoints, no
// workspace) at the price of ~110 FMAs per step, which moves the kernel from the HBM roof to the FP32
// issue roof.  Same thread-per-system / TMA-ring structure as the other rollout kernels.
//
// Writing rhs_i = (1/C_i) q_i with q0 = (u0-x0)/Rg + (x2-x0)/Ri + u1 + u2, q1 = (u0-x1)/Re + (x2-x1)/Rw + u3,
// q2 = (x0-x2)/Ri + (x1-x2)/Rw + u4, the direct partials are
//     d rhs0/dCai = -rhs0/Cai   d rhs1/dCwe = -rhs1/Cwe   d rhs2/dCwi = -rhs2/Cwi
//     d rhs1/dRe  = -(u0-x1)/(Cwe Re^2)        d rhs0/dRg = -(u0-x0)/(Cai Rg^2)
//     d rhs0/dRi  = -(x2-x0)/(Cai Ri^2)        d rhs2/dRi = -(x0-x2)/(Cwi Ri^2)
//     d rhs1/dRw  = -(x2-x1)/(Cwe Rw^2)        d rhs2/dRw = -(x1-x2)/(Cwi Rw^2)
// Gradient sums: fp32 over kFsFoldSteps steps, then folded into fp64 totals (as in the adjoint kernel).
#pragma once
#include "models.cuh"
#include "ptx.cuh"
#include "rollout_fwd.cuh"  // align4i

namespace dynax {

struct FwdSensParams {
    RCModel::Ptrs mp;
    const float *x0, *u, *target;  // [N,3], [T,N,5], [T,N]
    const float *lb, *ub;          // [7] or NULL
    float *loss, *g_theta;         // [N], [N,7]
    int64_t N, T;
    float dt;
    int bulk;
};

constexpr int kFsFoldSteps = 32;  // fp32 partial sums are folded into the fp64 totals every this many steps

template <int TN, int KF, int S>
struct FwdSensCfg {
    static constexpr int U_STAGE = KF * TN * 5, T_STAGE = KF * TN, STAGE = U_STAGE + T_STAGE;
    static constexpr size_t SMEM_BYTES = sizeof(float) * (size_t)(S * STAGE) + sizeof(uint64_t) * 2 * S;
    static constexpr int THREADS = TN + 32;
};

template <int TN, int KF, int S, int MINB>
__global__ void __launch_bounds__(TN + 32, MINB) rc_fwdsens_kernel(const FwdSensParams P) {
    using Cfg = FwdSensCfg<TN, KF, S>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *s_in = reinterpret_cast<float *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(s_in + S * Cfg::STAGE);
    uint64_t *empty = full + S;
    const int tid = threadIdx.x;
    const int64_t N = P.N, T = P.T;
    const int64_t n0 = (int64_t)blockIdx.x * TN;
    const int valid = (int)((N - n0) < TN ? (N - n0) : TN);
    const int64_t C = (T + KF - 1) / KF;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < S; ++s) {
            ptx::mbar_init(&full[s], 1);
            ptx::mbar_init(&empty[s], TN / 32);
        }
        ptx::fence_mbar_init();
    }
    __syncthreads();

    if (tid >= TN) {  // ------------------------------- producer warp -------------------------------
        const int lane = tid - TN;
        const uint64_t pol = ptx::policy_evict_first();
        auto load_chunk = [&](int64_t c) {
            const int s = (int)(c % S);
            const int64_t k0 = c * KF;
            const int rows = (int)((T - k0) < KF ? (T - k0) : KF);
            float *du = s_in + s * Cfg::STAGE, *dtg = du + Cfg::U_STAGE;
            if (P.bulk) {
                if (lane == 0) {
                    const uint32_t ub = (uint32_t)valid * 20u, tb = (uint32_t)valid * 4u;
                    ptx::mbar_arrive_expect_tx(&full[s], (ub + tb) * (uint32_t)rows);
                    for (int r = 0; r < rows; ++r) {
                        ptx::bulk_g2s(du + r * TN * 5, P.u + ((k0 + r) * N + n0) * 5, ub, &full[s], pol);
                        ptx::bulk_g2s(dtg + r * TN, P.target + (k0 + r) * N + n0, tb, &full[s], pol);
                    }
                }
            } else {
                for (int r = 0; r < rows; ++r) {
                    const float *su = P.u + ((k0 + r) * N + n0) * 5, *st = P.target + (k0 + r) * N + n0;
                    for (int i = lane; i < valid * 5; i += 32) du[r * TN * 5 + i] = __ldg(su + i);
                    for (int i = lane; i < valid; i += 32) dtg[r * TN + i] = __ldg(st + i);
                }
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&full[s]);
            }
        };
        for (int64_t c = 0; c < C && c < S; ++c) load_chunk(c);
        for (int64_t c = 0; c + S < C; ++c) {
            ptx::mbar_wait(&empty[c % S], (uint32_t)((c / S) & 1));
            load_chunk(c + S);
        }
        return;
    }

    // --------------------------------- consumers ---------------------------------
    const int lane = tid & 31;
    const bool active = tid < valid;
    const int64_t i = active ? (n0 + tid) : (N - 1);
    RCModel M;
    M.load(P.mp, i);
    const float *th = P.mp.theta + i * P.mp.stride;
    const float iCai = 1.0f / th[0], iCwe = 1.0f / th[1], iCwi = 1.0f / th[2];
    const float iRe = 1.0f / th[3], iRi = 1.0f / th[4], iRw = 1.0f / th[5], iRg = 1.0f / th[6];
    const float dt = P.dt;
    // dt-scaled coefficients of the sensitivity recurrence
    const float tA00 = dt * M.A00, tA02 = dt * M.A02, tA11 = dt * M.A11, tA12 = dt * M.A12, tA20 = dt * M.A20,
                tA21 = dt * M.A21, tA22 = dt * M.A22;
    const float tCai = -dt * iCai, tCwe = -dt * iCwe, tCwi = -dt * iCwi;
    const float tRe = -dt * iCwe * iRe * iRe, tRg = -dt * iCai * iRg * iRg;
    const float tRi0 = -dt * iCai * iRi * iRi, tRi2 = -dt * iCwi * iRi * iRi, tRw1 = -dt * iCwe * iRw * iRw,
                tRw2 = -dt * iCwi * iRw * iRw;
    float x0 = __ldg(P.x0 + i * 3), x1 = __ldg(P.x0 + i * 3 + 1), x2 = __ldg(P.x0 + i * 3 + 2);
    float s0[7], s1[7], s2[7];  // rows of S = dx/dtheta
    float gp[7];
    double gt[7], sqt = 0.0;
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < 7; ++j) {
        s0[j] = s1[j] = s2[j] = 0.f;
        gp[j] = 0.f;
        gt[j] = 0.0;
    }
    const float c2T = 2.0f / (float)T;

    for (int64_t c = 0; c < C; ++c) {
        const int s = (int)(c % S);
        const int64_t k0 = c * KF;
        const int rows = (int)((T - k0) < KF ? (T - k0) : KF);
        ptx::mbar_wait(&full[s], (uint32_t)((c / S) & 1));
        const float *in = s_in + s * Cfg::STAGE + tid * 5;
        const float *tg = s_in + s * Cfg::STAGE + Cfg::U_STAGE + tid;
#pragma unroll
        for (int jj = 0; jj < KF; ++jj) {
            if (jj < rows) {
                const float u0 = in[jj * TN * 5], u1 = in[jj * TN * 5 + 1], u2 = in[jj * TN * 5 + 2],
                            u3 = in[jj * TN * 5 + 3], u4 = in[jj * TN * 5 + 4];
                // loss and gradient use the PRE-update state (ysol[k] = x_k[0], simulator.py:38)
                const float res = x0 - tg[jj * TN];
                sq = fmaf(res, res, sq);
                const float gy = c2T * res;
#pragma unroll
                for (int j = 0; j < 7; ++j) gp[j] = fmaf(gy, s0[j], gp[j]);
                // rhs in the reference's order (fxx(x) + fxu(u))
                float r[3];
                {
                    const float xx[3] = {x0, x1, x2}, uu[5] = {u0, u1, u2, u3, u4};
                    M.rhs(xx, uu, r);
                }
                // S <- S + dt (A S + F) as FMA chains seeded with S: dt is folded into A's 7 non-zeros and into
                // the 9 direct-partial coefficients (S still advances in residual form; I + dt A is never formed)
                const float d_u0x0 = u0 - x0, d_u0x1 = u0 - x1, d_x2x0 = x2 - x0, d_x2x1 = x2 - x1;
#pragma unroll
                for (int j = 0; j < 7; ++j) {
                    const float a0 = s0[j], a1 = s1[j], a2 = s2[j];
                    s0[j] = fmaf(tA02, a2, fmaf(tA00, a0, a0));
                    s1[j] = fmaf(tA12, a2, fmaf(tA11, a1, a1));
                    s2[j] = fmaf(tA22, a2, fmaf(tA21, a1, fmaf(tA20, a0, a2)));
                }
                s0[0] = fmaf(tCai, r[0], s0[0]);         // d/dCai
                s1[1] = fmaf(tCwe, r[1], s1[1]);         // d/dCwe
                s2[2] = fmaf(tCwi, r[2], s2[2]);         // d/dCwi
                s1[3] = fmaf(tRe, d_u0x1, s1[3]);        // d/dRe
                s0[4] = fmaf(tRi0, d_x2x0, s0[4]);       // d/dRi
                s2[4] = fmaf(-tRi2, d_x2x0, s2[4]);      //   (x0 - x2 = -(x2 - x0))
                s1[5] = fmaf(tRw1, d_x2x1, s1[5]);       // d/dRw
                s2[5] = fmaf(-tRw2, d_x2x1, s2[5]);      //   (x1 - x2 = -(x2 - x1))
                s0[6] = fmaf(tRg, d_u0x0, s0[6]);        // d/dRg
                x0 = fmaf(dt, r[0], x0);
                x1 = fmaf(dt, r[1], x1);
                x2 = fmaf(dt, r[2], x2);
            }
        }
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&empty[s]);
        constexpr int FOLDC = kFsFoldSteps / KF > 0 ? kFsFoldSteps / KF : 1;  // chunks between folds
        if ((c + 1) % FOLDC == 0 || c + 1 == C) {
#pragma unroll
            for (int j = 0; j < 7; ++j) {
                gt[j] += (double)gp[j];
                gp[j] = 0.f;
            }
            sqt += (double)sq;
            sq = 0.f;
        }
    }
    if (!active) return;
    double loss = sqt / (double)T;
    if (P.lb != nullptr && P.ub != nullptr) {
#pragma unroll
        for (int j = 0; j < 7; ++j) {  // 3-inverse.py:103-107, normaliser = ub
            const double t = (double)th[j], lo = (double)P.lb[j], hi = (double)P.ub[j];
            if (t > hi) { loss += (t - hi) / hi; gt[j] += 1.0 / hi; }
            if (t < lo) { loss += (lo - t) / hi; gt[j] -= 1.0 / hi; }
        }
    }
    P.loss[i] = (float)loss;
#pragma unroll
    for (int j = 0; j < 7; ++j) P.g_theta[i * 7 + j] = (float)gt[j];
}

}  // namespace dynax
