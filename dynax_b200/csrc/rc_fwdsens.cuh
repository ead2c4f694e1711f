// Forward-mode (tangent) evaluation of the fused 4R3C loss + gradient: ONE sweep over u and target.
//
// The 4R3C model has only 7 parameters, so instead of the reverse sweep (rollout_adj.cuh) the 3x7
// sensitivity matrix S = dx/dtheta can ride along with the state ("alt." row):
//     x_{k+1} = x_k + dt (A x_k + B u_k)
//     S_{k+1} = S_k + dt (A S_k + F_k),      F_k[:,j] = (dA/dtheta_j) x_k + (dB/dtheta_j) u_k      (9 non-zeros)
//     loss    = mean_k (x_k[0] - target_k)^2,     dloss/dtheta_j = sum_k (2/T)(x_k[0] - target_k) S_k[0,j]
// This is mathematically the same gradient jax.value_and_grad returns (examples/3-inverse.py:96-112); it
// reads u and target ONCE (24 B per system-step instead of the adjoint's 46 B, no checkpoints, no
// workspace) at the price of ~90 FMAs per step, which moves the kernel from the HBM roof towards the FP32
// issue / power roof -- hence the packed-fp32 (FFMA2) form below.
// Same thread-per-system / TMA-ring structure as the other rollout kernels.
//
// Writing rhs_i = (1/C_i) q_i with q0 = (u0-x0)/Rg + (x2-x0)/Ri + u1 + u2, q1 = (u0-x1)/Re + (x2-x1)/Rw + u3,
// q2 = (x0-x2)/Ri + (x1-x2)/Rw + u4, the direct partials are
//     d rhs0/dCai = -rhs0/Cai   d rhs1/dCwe = -rhs1/Cwe   d rhs2/dCwi = -rhs2/Cwi
//     d rhs1/dRe  = -(u0-x1)/(Cwe Re^2)        d rhs0/dRg = -(u0-x0)/(Cai Rg^2)
//     d rhs0/dRi  = -(x2-x0)/(Cai Ri^2)        d rhs2/dRi = -(x0-x2)/(Cwi Ri^2)
//     d rhs1/dRw  = -(x2-x1)/(Cwe Rw^2)        d rhs2/dRw = -(x1-x2)/(Cwi Rw^2)
// Gradient sums: fp32 over kFsFoldSteps steps, then folded into fp64 totals (as in the adjoint kernel).
//
// Packed fp32 (FFMA2/FMUL2/FADD2, sm_100: two IEEE fp32 operations per issue slot):
//   * the columns of S travel in pairs (Cai,Cwe) (Cwi,Re) (Ri,Rw); the seventh column (Rg) travels as scalars
//     (GX0: (Rg, x0_0) (x0_1, x0_2) are two more pairs: d x / d x0, identity at k = 0, no forcing);
//   * rows 0 and 1 of the rhs, the differences u0 - x, x2 - x and the update of (x0,x1) are pairs too.
// Every lane performs the operations of RCModel::rhs in the same order (intrinsics), so x and the loss are
// bit-identical to every other kernel that steps the model.
//
// SPLIT (threads per system): the sensitivity columns are independent of each other given x, so a launch with too
// few systems to fill the GPU (BASELINE.json config 3 sharded over 8 GPUs: 8192 systems per rank) gives every system
// SPLIT threads -- "roles", one warp-uniform role per warp -- that each step x redundantly (bit-identical) and own
// a subset of the columns: the serial chain per thread drops from ~70 to ~35 instructions per step and four times
// as many warps share the 8760-step critical path.  No cross-thread reduction: a role writes its own columns.
#pragma once
#include <cuda.h>

#include "models.cuh"
#include "ptx.cuh"
#include "rollout_adj.cuh"  // warp_sum
#include "rollout_fwd.cuh"  // align4i

namespace dynax {

struct FwdSensParams {
    RCModel::Ptrs mp;
    const float *x0, *u, *target;  // [N,3], [T,N,5] (SHARED_U: [T,5]), [T,N] (SHARED_TGT: [T])
    const float *ctrl;             // CTRL mode (with SHARED_U): per-system channel [T,N] replacing column ctrl_col of u
    int ctrl_col;
    const float *lb, *ub;          // [7] or NULL
    float *loss, *g_theta;         // [N], [N,7]
    float *g_x0;                   // [N,3] d loss / d x0 (GX0 kernels only), or NULL
    double *shared_acc;            // shared theta (mp.stride == 0): fp64 sums over systems [7 grads, loss], zeroed by the caller
    int64_t N, T;
    float dt;
    int bulk;
    // time-chunked launch (CHUNK kernels, gridDim.y = chunks; lti_chunked.cu): chunk c covers steps [c*chunk_len, ...),
    // starts from x0 + c*x0_chunk_stride with S = 0 and reports its LOCAL results for the stitch over chunks
    int64_t chunk_len, x0_chunk_stride;
    float *s_end;   // [C,N,21]  S at the end of the chunk (row-major 3x7)
    double *g_loc;  // [C,N,8]   sum_k gy_k e0^T S_k (7) and sum_k res_k^2
    float *w_out;   // [C,N,3]   sum_k gy_k ((I + dt A)^T)^(k-k0) e0: what the chunk's loss sees of S at its start
    // tensor == 1: per-system rows arrive through 3-D tensor maps (tmap.cuh), ONE cp.async.bulk.tensor per array and
    // stage instead of one 1-D bulk copy per row; inner_* = the maps' inner extents (floats)
    int tensor, inner_u, inner_t;
    alignas(64) CUtensorMap map_u, map_t, map_c;   // u[T,N*5], target[T,N], ctrl[T,N]
};


constexpr int kFsFoldSteps = 32;  // fp32 partial sums are folded into the fp64 totals every this many steps

// SHIFT: the instantiation for rows TMA cannot move (N % 4 != 0; per-system u and target, one thread per system): rows are
// staged at their global offset modulo 16 bytes, so that the interior of a row moves as ONE bulk copy and only its < 16-byte
// head and tail by lanes (ptx::copy_rows_shift_bulk), the row pitches grow by 4 floats, and the computing threads add the
// (warp-uniform) shift of each row to their shared address.
template <int TN, int KF, int S, int SPLIT = 1, bool SHARED_U = false, bool SHARED_TGT = false, bool CTRL = false,
          bool GX0 = false, bool SHIFT = false>
struct FwdSensCfg {
    static_assert(!SHIFT || (SPLIT == 1 && !SHARED_U && !SHARED_TGT && !CTRL), "shifted rows: the per-system layout only");
    static_assert(!SHIFT || KF <= 32, "copy_rows_shift_bulk: one lane per row of a stage");
    static constexpr int NACC = GX0 ? 11 : 8;  // fp64 totals per system: 7 gradient entries (+3 for x0) + squared residual
    // a shared series occupies KF rows of 5 (or 1) floats per stage instead of KF x TN of them
    static constexpr int U_ROW = SHARED_U ? 5 : TN * 5 + (SHIFT ? 4 : 0), T_ROW = SHARED_TGT ? 1 : TN + (SHIFT ? 4 : 0);
    // every block of a stage starts on a 128-byte boundary (destination alignment of cp.async.bulk.tensor)
    static constexpr int U_STAGE = (KF * U_ROW + 31) / 32 * 32, T_STAGE = (KF * T_ROW + 31) / 32 * 32;
    static constexpr int C_STAGE = CTRL ? (KF * TN + 31) / 32 * 32 : 0, STAGE = U_STAGE + T_STAGE + C_STAGE;
    static constexpr size_t RING_BYTES = sizeof(float) * (size_t)(S * STAGE) + sizeof(uint64_t) * 2 * S;
    static constexpr size_t ACC_OFF = (RING_BYTES + 15) / 16 * 16;                 // fp64 totals [NACC][TN]
    static constexpr size_t SMEM_BYTES = ACC_OFF + sizeof(double) * NACC * (size_t)TN;
    static constexpr int CONSUMERS = TN * SPLIT, THREADS = CONSUMERS + 32;
};

// which column pairs (bit q = pair q) and whether the scalar seventh column a role owns
__host__ __device__ constexpr unsigned fs_pair_mask(int split, int role, bool gx0) {
    return split == 1   ? (gx0 ? 0x1fu : 0x07u)
           : split == 2 ? (role == 0 ? 0x03u : (gx0 ? 0x1cu : 0x04u))
                        : (role < 3 ? (1u << role) : (gx0 ? 0x18u : 0u));
}
__host__ __device__ constexpr bool fs_has_c(int split, int role, bool gx0) { return !gx0 && role == split - 1; }

template <int TN, int KF, int S, int SPLIT, int ROLE, bool SHARED_U, bool SHARED_TGT, bool CTRL, bool GX0, bool CHUNK, int UNR,
          bool SHIFT>
__device__ __forceinline__ void fs_consumer(const FwdSensParams &P, unsigned char *smem_raw, float *s_in, uint64_t *full,
                                            uint64_t *empty, const int sys, const int64_t n0, const int valid,
                                            const int64_t cidx, const int T, const int C, const bool self_rows) {
    using Cfg = FwdSensCfg<TN, KF, S, SPLIT, SHARED_U, SHARED_TGT, CTRL, GX0, SHIFT>;
    constexpr unsigned PM = fs_pair_mask(SPLIT, ROLE, GX0);
    constexpr bool HAS_C = fs_has_c(SPLIT, ROLE, GX0);
    constexpr bool LOSS_ROLE = ROLE == SPLIT - 1;   // squared residual, chunk weights, the loss output
    constexpr int NP = GX0 ? 5 : 3;
    constexpr bool NEED_DU = (PM & 0x2u) || (PM & 0x8u) || HAS_C, NEED_DX = (PM & 0x4u) != 0;
    const int lane = threadIdx.x & 31;
    const int64_t N = P.N;
    const bool active = sys < valid;
    const int64_t i = active ? (n0 + sys) : (N - 1);
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
    const float *xs0 = P.x0 + (CHUNK ? cidx * P.x0_chunk_stride : 0) + i * 3;
    float x2 = __ldg(xs0 + 2);
    float2 X01 = make_float2(__ldg(xs0), __ldg(xs0 + 1));
    float sq = 0.f;
    const float c2T = 2.0f / (float)P.T;  // the mean runs over the WHOLE horizon, also in a chunked launch
    const int ccol = CTRL ? P.ctrl_col : -1;
    // fp64 totals live in shared memory ([j][sys]: conflict-free 8-byte accesses), touched once per kFsFoldSteps
    // steps: registers the loop needs for its coefficients.  A role touches the slots of its own columns only.
    double *acc = reinterpret_cast<double *>(smem_raw + Cfg::ACC_OFF) + sys;
#pragma unroll
    for (int q = 0; q < NP; ++q) {
        if (PM >> q & 1u) {
            acc[(2 * q) * TN] = 0.0;
            acc[(2 * q + 1 != 7 ? 2 * q + 1 : 10) * TN] = 0.0;   // x0_0 (column 7) parks in slot 10: 7 holds sq
        }
    }
    if constexpr (HAS_C) acc[6 * TN] = 0.0;
    if constexpr (LOSS_ROLE) acc[7 * TN] = 0.0;
    float2 s0[NP], s1[NP], s2[NP], gp[NP];
    float c0 = 0.f, c1 = 0.f, c2 = 0.f, gc = 0.f;
#pragma unroll
    for (int q = 0; q < NP; ++q) s0[q] = s1[q] = s2[q] = gp[q] = make_float2(0.f, 0.f);
    if constexpr (GX0) {
        s0[3].y = 1.f;
        s1[4].x = 1.f;
        s2[4].y = 1.f;
    }
    const float2 pA00 = make_float2(tA00, tA00), pA02 = make_float2(tA02, tA02), pA11 = make_float2(tA11, tA11),
                 pA12 = make_float2(tA12, tA12), pA20 = make_float2(tA20, tA20), pA21 = make_float2(tA21, tA21),
                 pA22 = make_float2(tA22, tA22);
    const float2 rAd = make_float2(M.A00, M.A11), rAx2 = make_float2(M.A02, M.A12);   // rhs rows 0,1
    const float2 rB0 = make_float2(M.B00, M.B10), rB1 = make_float2(M.B01, M.B13);
    const float2 fCai = make_float2(tCai, 0.f), fCwe = make_float2(0.f, tCwe), fCwi = make_float2(tCwi, 0.f),
                 fRe = make_float2(0.f, tRe), fRi0 = make_float2(tRi0, 0.f), fRw1 = make_float2(0.f, tRw1);
    const float2 neg1 = make_float2(-1.f, -1.f), dt2 = make_float2(dt, dt), pR2 = make_float2(-tRi2, -tRw2);
    // CHUNK: p_j = ((I + dt A)^T)^j e0 and w = sum_k gy_k p_(k-k0), so that the part of this chunk's gradient that
    // comes from S at the chunk start is w^T S_start (added by the stitch over chunks)
    float p0 = 1.f, p1 = 0.f, p2 = 0.f, w0 = 0.f, w1 = 0.f, w2 = 0.f;

    const float *in = nullptr, *tg = nullptr;
    unsigned shu = 0, sht = 0;  // SHIFT: global float index of this chunk's first u / target row (row jj: + jj * N * 5 / N)
    struct Row {
        float u0, u1, u2, u3, u4, t;
    };
    auto load_row = [&](const int jj) {
        Row r;
        const float *inj = in + jj * Cfg::U_ROW + (SHIFT ? (int)((shu + (unsigned)jj * (unsigned)(N * 5)) & 3u) : 0);
        r.u0 = inj[0]; r.u1 = inj[1]; r.u2 = inj[2];
        r.u3 = inj[3]; r.u4 = inj[4];
        if constexpr (CTRL) {  // inputs assembled as in wrapper/envs/rc.py:128-131
            const float cv = tg[Cfg::T_STAGE + (SHARED_TGT ? sys : 0) + jj * TN];
            r.u0 = ccol == 0 ? cv : r.u0; r.u1 = ccol == 1 ? cv : r.u1; r.u2 = ccol == 2 ? cv : r.u2;
            r.u3 = ccol == 3 ? cv : r.u3; r.u4 = ccol == 4 ? cv : r.u4;
        }
        r.t = tg[jj * Cfg::T_ROW + (SHIFT ? (int)((sht + (unsigned)jj * (unsigned)N) & 3u) : 0)];
        return r;
    };
    auto step = [&](const Row &row) {
        const float u0 = row.u0, u1 = row.u1, u2 = row.u2, u3 = row.u3, u4 = row.u4;
        const float res = X01.x - row.t;   // PRE-update state (simulator.py:38)
        if constexpr (LOSS_ROLE) sq = fmaf(res, res, sq);
        const float gy = c2T * res;
        const float2 gy2 = make_float2(gy, gy);
#pragma unroll
        for (int q = 0; q < NP; ++q)
            if (PM >> q & 1u) gp[q] = __ffma2_rn(gy2, s0[q], gp[q]);
        if constexpr (HAS_C) gc = fmaf(gy, c0, gc);
        if constexpr (CHUNK && LOSS_ROLE) {
            w0 = fmaf(gy, p0, w0); w1 = fmaf(gy, p1, w1); w2 = fmaf(gy, p2, w2);
            const float q0 = fmaf(tA20, p2, fmaf(tA00, p0, p0));                     // p + dt A^T p
            const float q1 = fmaf(tA21, p2, fmaf(tA11, p1, p1));
            const float q2 = fmaf(tA22, p2, fmaf(tA12, p1, fmaf(tA02, p0, p2)));
            p0 = q0; p1 = q1; p2 = q2;
        }
        // rhs = (A x) + (B u), rows 0 and 1 as a pair, row 2 scalar
        const float2 U0 = make_float2(u0, u0), X2 = make_float2(x2, x2);
        const float2 AX = __ffma2_rn(rAd, X01, __fmul2_rn(rAx2, X2));
        float2 BU = __ffma2_rn(rB0, U0, __fmul2_rn(rB1, make_float2(u1, u3)));
        BU.x = fmaf(M.B02, u2, BU.x);
        const float2 R01 = __fadd2_rn(AX, BU);
        const float r2 = __fadd_rn(fmaf(M.A22, x2, fmaf(M.A20, X01.x, __fmul_rn(M.A21, X01.y))), __fmul_rn(M.B24, u4));
        float2 DU = make_float2(0.f, 0.f), DX = make_float2(0.f, 0.f);
        if constexpr (NEED_DU) DU = __ffma2_rn(neg1, X01, U0);   // (u0 - x0, u0 - x1), exact as a subtraction
        if constexpr (NEED_DX) DX = __ffma2_rn(neg1, X01, X2);   // (x2 - x0, x2 - x1)
        // S <- S + dt A S
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            if (PM >> q & 1u) {
                const float2 a0 = s0[q], a1 = s1[q], a2 = s2[q];
                s0[q] = __ffma2_rn(pA02, a2, __ffma2_rn(pA00, a0, a0));
                s1[q] = __ffma2_rn(pA12, a2, __ffma2_rn(pA11, a1, a1));
                s2[q] = __ffma2_rn(pA22, a2, __ffma2_rn(pA21, a1, __ffma2_rn(pA20, a0, a2)));
            }
        }
        if constexpr (HAS_C) {  // the same three expressions, one lane
            const float a0 = c0, a1 = c1, a2 = c2;
            c0 = fmaf(tA02, a2, fmaf(tA00, a0, a0));
            c1 = fmaf(tA12, a2, fmaf(tA11, a1, a1));
            c2 = fmaf(tA22, a2, fmaf(tA21, a1, fmaf(tA20, a0, a2)));
        }
        if constexpr ((PM & 0x1u) != 0) {
            s0[0].x = fmaf(tCai, R01.x, s0[0].x);      // d/dCai   (column 0)
            s1[0].y = fmaf(tCwe, R01.y, s1[0].y);      // d/dCwe   (column 1)
        }
        if constexpr ((PM & 0x2u) != 0) {
            s2[1].x = fmaf(tCwi, r2, s2[1].x);         // d/dCwi   (column 2)
            s1[1].y = fmaf(tRe, DU.y, s1[1].y);        // d/dRe    (column 3)
        }
        if constexpr ((PM & 0x4u) != 0) {
            s0[2].x = fmaf(tRi0, DX.x, s0[2].x);       // d/dRi    (column 4), row 0
            s1[2].y = fmaf(tRw1, DX.y, s1[2].y);       // d/dRw    (column 5), row 1
            s2[2] = __ffma2_rn(pR2, DX, s2[2]);        // d/dRi, d/dRw, row 2
        }
        if constexpr ((PM & 0x8u) != 0) s0[3].x = fmaf(tRg, DU.x, s0[3].x);   // d/dRg    (column 6)
        if constexpr (HAS_C) c0 = fmaf(tRg, DU.x, c0);
        X01 = __ffma2_rn(dt2, R01, X01);
        x2 = fmaf(dt, r2, x2);
    };

    // One predicated body per chunk, unrolled UNR steps at a time; the last chunk may be ragged.  (Measured on B200: the
    // per-row predicates double as scheduling fences -- ptxas keeps the loads of step j+1 behind step j, live ranges
    // stay short at 96 registers -- and this form is 3-4 % FASTER than an unpredicated body for full chunks; fetching
    // row j+1 into registers while row j computes is slower too: profiles/r02_headline_variants.txt.)
    // fp32 partial sums are folded into the fp64 totals at every multiple of kFsFoldSteps steps of the horizon and at
    // its end, whatever the chunk length: results do not depend on the tile shape (bit for bit).
    static_assert(KF % kFsFoldSteps == 0 || kFsFoldSteps % KF == 0, "chunk length and fold interval must nest");
    constexpr int FR = KF > kFsFoldSteps ? kFsFoldSteps : KF;       // rows between folds inside a chunk
    constexpr int FOLDC = kFsFoldSteps / KF > 0 ? kFsFoldSteps / KF : 1;
    auto fold = [&]() {
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            if (PM >> q & 1u) {
                acc[(2 * q) * TN] += (double)gp[q].x;
                acc[(2 * q + 1 != 7 ? 2 * q + 1 : 10) * TN] += (double)gp[q].y;
                gp[q] = make_float2(0.f, 0.f);
            }
        }
        if constexpr (HAS_C) {
            acc[6 * TN] += (double)gc;
            gc = 0.f;
        }
        if constexpr (LOSS_ROLE) {
            acc[7 * TN] += (double)sq;
            sq = 0.f;
        }
    };
    // self_rows (rows TMA cannot move, per-system u and target, one thread per system): this thread copies ITS OWN six
    // floats of every row of a chunk into its slice of the ring with 4-byte cp.async -- nobody else reads them, so the
    // stage needs no barrier, only the thread's own cp.async groups (one per chunk, committed even when empty)
    auto self_load = [&](const int c) {
        if constexpr (SPLIT == 1 && !SHARED_U && !SHARED_TGT && !CTRL) {
            if (c < C && active) {
                const int64_t k0 = (int64_t)(CHUNK ? cidx * P.chunk_len : 0) + (int64_t)c * KF;
                const int rows = (T - c * KF) < KF ? (T - c * KF) : KF;
                float *du = s_in + (c % S) * Cfg::STAGE + sys * 5, *dtg = s_in + (c % S) * Cfg::STAGE + Cfg::U_STAGE + sys;
                const float *su = P.u + (k0 * N + i) * 5, *st = P.target + k0 * N + i;
#pragma unroll
                for (int j = 0; j < KF; ++j) {
                    if (j < rows) {
#pragma unroll
                        for (int d = 0; d < 5; ++d) ptx::cp_async4(du + j * Cfg::U_ROW + d, su + j * N * 5 + d);
                        ptx::cp_async4(dtg + j * Cfg::T_ROW, st + j * N);
                    }
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
    };
    if (self_rows) {
        for (int c = 0; c < S; ++c) self_load(c);
    }
    for (int c = 0; c < C; ++c) {
        const int s = c % S;
        const int k0 = c * KF;
        const int rows = (T - k0) < KF ? (T - k0) : KF;
        if (self_rows) asm volatile("cp.async.wait_group %0;" ::"n"(S - 1) : "memory");
        else ptx::mbar_wait(&full[s], (uint32_t)((c / S) & 1));
        in = s_in + s * Cfg::STAGE + (SHARED_U ? 0 : sys * 5);
        tg = s_in + s * Cfg::STAGE + Cfg::U_STAGE + (SHARED_TGT ? 0 : sys);
        if constexpr (SHIFT) {
            const int64_t kg = (int64_t)(CHUNK ? cidx * P.chunk_len : 0) + k0;
            shu = (unsigned)((reinterpret_cast<uintptr_t>(P.u) >> 2) + (uint64_t)((kg * N + n0) * 5));
            sht = (unsigned)((reinterpret_cast<uintptr_t>(P.target) >> 2) + (uint64_t)(kg * N + n0));
        }
#pragma unroll 1
        for (int j0 = 0; j0 < KF; j0 += FR) {
#pragma unroll UNR
            for (int jj = 0; jj < FR; ++jj)
                if (j0 + jj < rows) step(load_row(j0 + jj));
            if (KF > kFsFoldSteps && j0 + FR < KF) fold();      // inside a long chunk; its last block folds below
        }
        if (self_rows) self_load(c + S);
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&empty[s]);
        if ((c + 1) % FOLDC == 0 || c + 1 == C) fold();
    }

    // ---- outputs: every role reports the columns it owns ----
    // column j of the gradient lives in slot j (j < 7); x0 columns: x0_0 -> slot 10, x0_1 -> 8, x0_2 -> 9
    constexpr unsigned COLS = (PM & 1u ? 0x03u : 0u) | (PM & 2u ? 0x0cu : 0u) | (PM & 4u ? 0x30u : 0u) |
                              ((PM & 8u) || HAS_C ? 0x40u : 0u);
    if constexpr (CHUNK) {  // local results of this chunk; the stitch kernel assembles loss and gradient
        if (active) {
            const int64_t o = cidx * N + i;
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                if (PM >> q & 1u) {
                    P.s_end[o * 21 + 2 * q] = s0[q].x;
                    P.s_end[o * 21 + 7 + 2 * q] = s1[q].x;
                    P.s_end[o * 21 + 14 + 2 * q] = s2[q].x;
                    P.s_end[o * 21 + 2 * q + 1] = s0[q].y;
                    P.s_end[o * 21 + 7 + 2 * q + 1] = s1[q].y;
                    P.s_end[o * 21 + 14 + 2 * q + 1] = s2[q].y;
                }
            }
            if constexpr (HAS_C) {
                P.s_end[o * 21 + 6] = c0;
                P.s_end[o * 21 + 13] = c1;
                P.s_end[o * 21 + 20] = c2;
            }
#pragma unroll
            for (int j = 0; j < 7; ++j)
                if (COLS >> j & 1u) P.g_loc[o * 8 + j] = acc[j * TN];
            if constexpr (LOSS_ROLE) {
                P.g_loc[o * 8 + 7] = acc[7 * TN];
                P.w_out[o * 3 + 0] = w0; P.w_out[o * 3 + 1] = w1; P.w_out[o * 3 + 2] = w2;
            }
        }
        return;
    }
    if constexpr (GX0) {
        if (active && P.g_x0 != nullptr) {
            if constexpr ((PM & 0x8u) != 0) P.g_x0[i * 3 + 0] = (float)acc[10 * TN];
            if constexpr ((PM & 0x10u) != 0) {
                P.g_x0[i * 3 + 1] = (float)acc[8 * TN];
                P.g_x0[i * 3 + 2] = (float)acc[9 * TN];
            }
        }
    }
    double loss = LOSS_ROLE ? acc[7 * TN] / (double)T : 0.0;
    if (P.shared_acc != nullptr) {
        // ONE parameter vector for all systems: sum the per-system gradients (warp shuffle in fp64 -> one atomic
        // per warp and output); the penalty is added once by shared_finalize_kernel
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            if (COLS >> j & 1u) {
                const double v = warp_sum(active ? acc[j * TN] : 0.0);
                if (lane == 0) atomicAdd(P.shared_acc + j, v);
            }
        }
        if constexpr (LOSS_ROLE) {
            const double v = warp_sum(active ? loss : 0.0);
            if (lane == 0) atomicAdd(P.shared_acc + 7, v);
        }
        return;
    }
    if (!active) return;
    const bool box = P.lb != nullptr && P.ub != nullptr;
#pragma unroll
    for (int j = 0; j < 7; ++j) {  // 3-inverse.py:103-107, normaliser = ub
        double g = (COLS >> j & 1u) ? acc[j * TN] : 0.0;
        if (box) {
            const double t = (double)th[j], lo = (double)P.lb[j], hi = (double)P.ub[j];
            if (t > hi) { loss += (t - hi) / hi; g += 1.0 / hi; }
            if (t < lo) { loss += (lo - t) / hi; g -= 1.0 / hi; }
        }
        if ((COLS >> j & 1u) && P.g_theta != nullptr) P.g_theta[i * 7 + j] = (float)g;
    }
    if constexpr (LOSS_ROLE)
        if (P.loss != nullptr) P.loss[i] = (float)loss;
}

template <int TN, int KF, int S, int MINB, int SPLIT = 1, bool SHARED_U = false, bool SHARED_TGT = false,
          bool CTRL = false, bool GX0 = false, bool CHUNK = false, int UNR = KF, bool SHIFT = false>
__global__ void __launch_bounds__(TN * SPLIT + 32, MINB) rc_fwdsens_kernel(const __grid_constant__ FwdSensParams P) {
    static_assert(!CTRL || SHARED_U, "the control channel replaces a column of a SHARED input series");
    static_assert(!CHUNK || (!GX0 && !CTRL), "time-chunked launches use the plain kernel");
    static_assert(SPLIT == 1 || SPLIT == 2 || SPLIT == 4, "1, 2 or 4 threads per system");
    static_assert(TN % 32 == 0, "a warp holds one role");
    using Cfg = FwdSensCfg<TN, KF, S, SPLIT, SHARED_U, SHARED_TGT, CTRL, GX0, SHIFT>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *s_in = reinterpret_cast<float *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(s_in + S * Cfg::STAGE);
    uint64_t *empty = full + S;
    const int tid = threadIdx.x;
    const int64_t N = P.N;
    const int64_t n0 = (int64_t)blockIdx.x * TN;
    const int valid = (int)((N - n0) < TN ? (N - n0) : TN);
    // this CTA's slice of the time axis (the whole horizon unless the launch is time-chunked)
    const int64_t cidx = CHUNK ? (int64_t)blockIdx.y : 0;
    const int64_t t_begin = CHUNK ? cidx * P.chunk_len : 0;
    const int T = (int)(CHUNK ? ((P.T - t_begin) < P.chunk_len ? (P.T - t_begin) : P.chunk_len) : P.T);
    const float *const u_base = P.u + t_begin * (SHARED_U ? 5 : N * 5);
    const float *const tgt_base = P.target + t_begin * (SHARED_TGT ? 1 : N);
    const int C = (T + KF - 1) / KF;
    // per-system rows that TMA cannot move (N % 4 != 0) go through 4-byte cp.async from all producer lanes: every
    // lane then arrives on the stage's barrier (count 32) instead of the one elected lane of the TMA path
    constexpr bool ANY_ROWS = !SHARED_U || !SHARED_TGT || CTRL;
    // ... except in the headline layout (per-system u and target, one thread per system), where every computing thread
    // fetches its own rows (fs_consumer: self_load) and the producer warp has nothing to do (N = 151 553, T = 2000:
    // 3.05 -> 2.04 ms; the forward kernel, which also stores x and y per thread, stays with the producer lanes)
    const bool self_rows = !SHIFT && SPLIT == 1 && !SHARED_U && !SHARED_TGT && !CTRL && !P.bulk;
    const bool async_rows = ANY_ROWS && !P.bulk && !self_rows;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < S; ++s) {
            ptx::mbar_init(&full[s], async_rows ? 32 : 1);
            ptx::mbar_init(&empty[s], Cfg::CONSUMERS / 32);
        }
        ptx::fence_mbar_init();
    }
    __syncthreads();

    if (tid >= Cfg::CONSUMERS) {  // ------------------------------- producer warp -------------------------------
        if (self_rows) return;
        const int lane = tid - Cfg::CONSUMERS;
        const uint64_t pol = ptx::policy_evict_first(), pol_keep = ptx::policy_evict_last();
        auto load_chunk = [&](int c) {
            const int s = c % S;
            const int64_t k0 = (int64_t)c * KF;
            const int rows = (int)((T - k0) < KF ? (T - k0) : KF);
            float *du = s_in + s * Cfg::STAGE, *dtg = du + Cfg::U_STAGE;
            // Shared series: one bulk copy per block of rows when it is 16-byte aligned and sized (asynchronous, so the
            // producer does not pay an L2 round trip per chunk), else plain loads by the producer lanes.  Per-system
            // rows: one bulk copy per row, or plain loads.  The arrive that publishes the stage comes last.
            uint32_t sh_tx = 0;
            bool su_bulk = false, st_bulk = false;
            if constexpr (SHARED_U) {
                const float *su = u_base + k0 * 5;
                su_bulk = (reinterpret_cast<uintptr_t>(su) & 15) == 0 && ((rows * 20) & 15) == 0;
                if (su_bulk) sh_tx += (uint32_t)rows * 20u;
                else if (async_rows) ptx::copy_rows_async(du, 0, su, 0, 1, rows * 5, lane);
                else for (int e = lane; e < rows * 5; e += 32) du[e] = __ldg(su + e);
            }
            if constexpr (SHARED_TGT) {
                const float *st = tgt_base + k0;
                st_bulk = (reinterpret_cast<uintptr_t>(st) & 15) == 0 && ((rows * 4) & 15) == 0;
                if (st_bulk) sh_tx += (uint32_t)rows * 4u;
                else if (async_rows) ptx::copy_rows_async(dtg, 0, st, 0, 1, rows, lane);
                else for (int e = lane; e < rows; e += 32) dtg[e] = __ldg(st + e);
            }
            float *dc = dtg + Cfg::T_STAGE;
            if (async_rows) {
                if (lane == 0 && sh_tx != 0) {
                    ptx::mbar_expect_tx(&full[s], sh_tx);
                    if (SHARED_U && su_bulk) ptx::bulk_g2s(du, u_base + k0 * 5, (uint32_t)rows * 20u, &full[s], pol_keep);
                    if (SHARED_TGT && st_bulk) ptx::bulk_g2s(dtg, tgt_base + k0, (uint32_t)rows * 4u, &full[s], pol_keep);
                }
                if constexpr (SHIFT) {
                    ptx::copy_rows_shift_bulk(du, Cfg::U_ROW, u_base + (k0 * N + n0) * 5, N * 5, rows, valid * 5, lane, &full[s], pol);
                    ptx::copy_rows_shift_bulk(dtg, Cfg::T_ROW, tgt_base + k0 * N + n0, N, rows, valid, lane, &full[s], pol);
                } else {
                    if constexpr (!SHARED_U) ptx::copy_rows_async(du, TN * 5, u_base + (k0 * N + n0) * 5, N * 5, rows, valid * 5, lane);
                    if constexpr (!SHARED_TGT) ptx::copy_rows_async(dtg, TN, tgt_base + k0 * N + n0, N, rows, valid, lane);
                }
                if constexpr (CTRL) ptx::copy_rows_async(dc, TN, P.ctrl + k0 * N + n0, N, rows, valid, lane);
                ptx::cp_async_arrive(&full[s]);  // every lane: arrives once its own copies have landed
                return;
            }
            __syncwarp();
            if (P.tensor) {
                // one tensor copy per array: the box is the whole stage (KF rows x this CTA's slice; rows past the end of
                // the horizon and systems past N arrive as zeros, the transaction count is always the full box)
                if (lane == 0) {
                    constexpr uint32_t box_tx = (SHARED_U ? 0u : (uint32_t)(KF * TN * 20)) + (SHARED_TGT ? 0u : (uint32_t)(KF * TN * 4)) +
                                                (CTRL ? (uint32_t)(KF * TN * 4) : 0u);
                    ptx::mbar_arrive_expect_tx(&full[s], sh_tx + box_tx);
                    if (SHARED_U && su_bulk) ptx::bulk_g2s(du, u_base + k0 * 5, (uint32_t)rows * 20u, &full[s], pol_keep);
                    if (SHARED_TGT && st_bulk) ptx::bulk_g2s(dtg, tgt_base + k0, (uint32_t)rows * 4u, &full[s], pol_keep);
                    const int kk = (int)(t_begin + k0);
                    if constexpr (!SHARED_U) ptx::tensor_g2s_3d(du, &P.map_u, 0, (int)(n0 * 5 / P.inner_u), kk, &full[s], pol);
                    if constexpr (!SHARED_TGT) ptx::tensor_g2s_3d(dtg, &P.map_t, 0, (int)(n0 / P.inner_t), kk, &full[s], pol);
                    if constexpr (CTRL) ptx::tensor_g2s_3d(dc, &P.map_c, 0, (int)(n0 / P.inner_t), kk, &full[s], pol);
                }
                return;
            }
            // per-row copies (no tensor map for this layout): lane r issues row r -- one thread gets a bulk copy out every
            // ~45 ns only (profiles/r02_tma_probe.txt), the lanes' copies leave together
            const uint32_t ub = SHARED_U ? 0u : (uint32_t)valid * 20u, tb = SHARED_TGT ? 0u : (uint32_t)valid * 4u,
                           cb = CTRL ? (uint32_t)valid * 4u : 0u;
            const uint32_t tx = sh_tx + (ANY_ROWS ? (ub + tb + cb) * (uint32_t)rows : 0u);
            if (lane == 0) {
                if (tx == 0) {
                    ptx::mbar_arrive(&full[s]);  // everything was staged with plain stores
                } else {
                    ptx::mbar_arrive_expect_tx(&full[s], tx);
                    if (SHARED_U && su_bulk) ptx::bulk_g2s(du, u_base + k0 * 5, (uint32_t)rows * 20u, &full[s], pol_keep);
                    if (SHARED_TGT && st_bulk) ptx::bulk_g2s(dtg, tgt_base + k0, (uint32_t)rows * 4u, &full[s], pol_keep);
                }
            }
            __syncwarp();
            for (int r = lane; ANY_ROWS && tx != 0 && r < rows; r += 32) {
                if constexpr (!SHARED_U) ptx::bulk_g2s(du + r * TN * 5, u_base + ((k0 + r) * N + n0) * 5, ub, &full[s], pol);
                if constexpr (!SHARED_TGT) ptx::bulk_g2s(dtg + r * TN, tgt_base + (k0 + r) * N + n0, tb, &full[s], pol);
                if constexpr (CTRL) ptx::bulk_g2s(dc + r * TN, P.ctrl + (k0 + r) * N + n0, cb, &full[s], pol);
            }
        };
        for (int c = 0; c < C && c < S; ++c) load_chunk(c);
        for (int c = 0; c + S < C; ++c) {
            ptx::mbar_wait(&empty[c % S], (uint32_t)((c / S) & 1));
            load_chunk(c + S);
        }
        return;
    }

    // --------------------------------- consumers: role = warp-uniform ---------------------------------
    const int role = SPLIT == 1 ? 0 : tid / TN, sys = SPLIT == 1 ? tid : tid % TN;
#define DYNAX_FS_ROLE(R)                                                                                         \
    fs_consumer<TN, KF, S, SPLIT, R, SHARED_U, SHARED_TGT, CTRL, GX0, CHUNK, UNR, SHIFT>(P, smem_raw, s_in, full, empty, sys, n0, \
                                                                                  valid, cidx, T, C, self_rows)
    if constexpr (SPLIT == 1) {
        DYNAX_FS_ROLE(0);
    } else if constexpr (SPLIT == 2) {
        if (role == 0) DYNAX_FS_ROLE(0);
        else DYNAX_FS_ROLE(1);
    } else {
        if (role == 0) DYNAX_FS_ROLE(0);
        else if (role == 1) DYNAX_FS_ROLE(1);
        else if (role == 2) DYNAX_FS_ROLE(2);
        else DYNAX_FS_ROLE(3);
    }
#undef DYNAX_FS_ROLE
}

}  // namespace dynax
