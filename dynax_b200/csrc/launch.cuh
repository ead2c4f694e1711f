// Kernel launch helpers shared by rc_api.cu and lti_api.cu.
#pragma once
#include <stdlib.h>

#include "api_common.cuh"
#include "rollout_adj.cuh"
#include "rc_fwdsens.cuh"
#include "rollout_fwd.cuh"
#include "tmap.cuh"

namespace dynax {

// Scratch layout: [0, kAccBytes) fp64 accumulators for shared-parameter sums, then checkpoints.
constexpr size_t kAccBytes = 1024;
constexpr int kMinKC = 8;  // smallest checkpoint interval any instantiated config uses

inline size_t adj_workspace_bytes(int64_t N, int64_t T, int n) {
    const int64_t C = (T + kMinKC - 1) / kMinKC;
    return (kAccBytes + sizeof(float) * (size_t)C * (size_t)n * (size_t)N + 15) / 16 * 16;
}
// VJP with a SHARED input series: per-CTA partial sums of g_u ([ceil(N/128)][T][m] floats) behind the checkpoints
constexpr int kGuPartTN = 128;
inline size_t gu_part_bytes(int64_t N, int64_t T, int m) {
    return sizeof(float) * (size_t)((N + kGuPartTN - 1) / kGuPartTN) * (size_t)T * (size_t)m;
}

template <class Model, int TN, int KF, int S, bool SHARED_U, int MINB, bool SHIFT = false, int NO = 2>
int launch_fwd(const FwdParams<Model> &P0, cudaStream_t st) {
    using Cfg = FwdCfg<Model, TN, KF, S, SHARED_U, SHIFT, NO>;
    auto kern = rollout_fwd_kernel<Model, TN, KF, S, SHARED_U, MINB, SHIFT, NO>;
    if (SHIFT && P0.bulk_in) return fail(DYNAX_ERR_INVALID_ARGUMENT, "shifted-row instantiation launched on rows TMA can move");
    DYNAX_NOTE_LAUNCH();
    int rc = set_smem(kern, Cfg::SMEM_BYTES);
    if (rc != DYNAX_OK) return rc;
    // tensor maps (one copy per array and stage) wherever the layout allows; DYNAX_NO_TENSOR=1 keeps per-row copies
    FwdParams<Model> P = P0;
    P.tensor_in = P.tensor_out = 0;
    if (!getenv("DYNAX_NO_TENSOR") && P.T > 0) {
        constexpr int n = Model::n, m = Model::m, p = Model::p;
        if (P.bulk_in && (!SHARED_U || P.ctrl)) {
            const int iu = SHARED_U ? 1 : row_map_inner(P.N, m, TN), ic = P.ctrl ? row_map_inner(P.N, 1, TN) : 1;
            if (iu > 0 && ic > 0) {
                if (!SHARED_U) rc = make_row_map(&P.map_u, P.u, P.N, m, P.T, TN, KF, iu);
                if (rc == DYNAX_OK && P.ctrl) rc = make_row_map(&P.map_c, P.ctrl, P.N, 1, P.T, TN, KF, ic);
                if (rc != DYNAX_OK) return rc;
                P.tensor_in = 1; P.inner_u = iu; P.inner_c = ic;
            }
        }
        if (P.bulk_out && (P.xsol || P.ysol)) {
            const int ix = P.xsol ? row_map_inner(P.N, n, TN) : 1, iy = P.ysol ? row_map_inner(P.N, p, TN) : 1;
            if (ix > 0 && iy > 0) {
                if (P.xsol) rc = make_row_map(&P.map_x, P.xsol, P.N, n, P.T, TN, KF, ix);
                if (rc == DYNAX_OK && P.ysol) rc = make_row_map(&P.map_y, P.ysol, P.N, p, P.T, TN, KF, iy);
                if (rc != DYNAX_OK) return rc;
                P.tensor_out = 1; P.inner_x = ix; P.inner_y = iy;
            }
        }
    }
    const int64_t grid = (P.N + TN - 1) / TN;
    if (grid > 0x7fffffffLL) return fail(DYNAX_ERR_INVALID_ARGUMENT, "N too large for one launch");
    const int64_t chunks = P.chunk_len > 0 ? (P.T + P.chunk_len - 1) / P.chunk_len : 1;
    if (chunks > 65535) return fail(DYNAX_ERR_INVALID_ARGUMENT, "too many time chunks");
    kern<<<dim3((unsigned)grid, (unsigned)chunks), Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(P);
    DYNAX_CUDA_TRY(cudaGetLastError());
    return DYNAX_OK;
}

template <class Model, int TN, int KC, int S, int MODE, bool SHARED_U, bool SHARED_TGT, int MINB, bool SHIFT = false>
int launch_adj(const AdjParams<Model> &P0, cudaStream_t st) {
    using Cfg = AdjCfg<Model, TN, KC, S, MODE, SHARED_U, SHARED_TGT, SHIFT>;
    static_assert(KC >= kMinKC, "workspace sizing assumes KC >= kMinKC");
    auto kern = rollout_adj_kernel<Model, TN, KC, S, MODE, SHARED_U, SHARED_TGT, MINB, SHIFT>;
    if (SHIFT && (P0.bulk || P0.ctrl || P0.chunk_len > 0))
        return fail(DYNAX_ERR_INVALID_ARGUMENT, "shifted-row instantiation launched on rows TMA can move / a chunked launch");
    DYNAX_NOTE_LAUNCH();
    int rc = set_smem(kern, Cfg::SMEM_BYTES);
    if (rc != DYNAX_OK) return rc;
    // tensor maps (one copy per array and stage) when every per-system array of this launch allows one;
    // DYNAX_NO_TENSOR=1 keeps the per-row copies
    AdjParams<Model> P = P0;
    P.tensor = 0;
    if (MODE == ADJ_VJP && SHARED_U) P.g_u = nullptr;   // a shared series gets ONE g_u row per step: gu_part + gu_reduce_kernel
    if (P.bulk && !getenv("DYNAX_NO_TENSOR")) {
        constexpr int n = Model::n, m = Model::m, p = Model::p;
        const float *a1 = MODE == ADJ_MSE ? (SHARED_TGT ? nullptr : P.target) : P.g_ysol;
        const float *a2 = MODE == ADJ_MSE ? nullptr : P.g_xsol;
        const int w1 = MODE == ADJ_MSE ? 1 : p;
        const int iu = SHARED_U ? 1 : row_map_inner(P.N, m, TN), ic = (SHARED_U && P.ctrl) ? row_map_inner(P.N, 1, TN) : 1;
        const int i1 = a1 ? row_map_inner(P.N, w1, TN) : 1, i2 = a2 ? row_map_inner(P.N, n, TN) : 1;
        const bool any = !SHARED_U || (SHARED_U && P.ctrl) || a1 || a2;
        if (any && iu > 0 && ic > 0 && i1 > 0 && i2 > 0) {
            if (!SHARED_U) rc = make_row_map(&P.map_u, P.u, P.N, m, P.T, TN, KC, iu);
            if (rc == DYNAX_OK && SHARED_U && P.ctrl) rc = make_row_map(&P.map_c, P.ctrl, P.N, 1, P.T, TN, KC, ic);
            if (rc == DYNAX_OK && a1) rc = make_row_map(&P.map_a1, a1, P.N, w1, P.T, TN, KC, i1);
            if (rc == DYNAX_OK && a2) rc = make_row_map(&P.map_a2, a2, P.N, n, P.T, TN, KC, i2);
            if (rc != DYNAX_OK) return rc;
            P.tensor = 1; P.inner_u = iu; P.inner_c = ic; P.inner_a1 = i1; P.inner_a2 = i2;
        }
    }
    const int64_t grid = (P.N + TN - 1) / TN;
    if (grid > 0x7fffffffLL) return fail(DYNAX_ERR_INVALID_ARGUMENT, "N too large for one launch");
    if (P.chunk_len > 0) {  // time-chunked launch: the caller (rc_chunked_adj.cu) owns zeroing and finalisation
        const int64_t chunks = (P.T + P.chunk_len - 1) / P.chunk_len;
        if (chunks > 65535) return fail(DYNAX_ERR_INVALID_ARGUMENT, "too many time chunks");
        kern<<<dim3((unsigned)grid, (unsigned)chunks), Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(P);
        DYNAX_CUDA_TRY(cudaGetLastError());
        return DYNAX_OK;
    }
    if (P.shared_theta) DYNAX_CUDA_TRY(cudaMemsetAsync(P.shared_acc, 0, kAccBytes, st));
    static_assert(!(MODE == ADJ_VJP && SHARED_U) || TN == kGuPartTN, "gu_part is sized for 128-system CTAs");
    kern<<<(unsigned)grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(P);
    DYNAX_CUDA_TRY(cudaGetLastError());
    if (MODE == ADJ_VJP && SHARED_U && P.gu_part != nullptr && P0.g_u != nullptr) {
        const int64_t elems = P.T * Model::m;
        gu_reduce_kernel<<<(unsigned)((elems + 255) / 256), 256, 0, st>>>(P.gu_part, P0.g_u, elems, grid);
        DYNAX_CUDA_TRY(cudaGetLastError());
    }
    if (P.shared_theta) {
        shared_finalize_kernel<Model, MODE><<<1, 32, 0, st>>>(P.shared_acc, P.mp, P.gp, P.lb, P.ub, P.loss);
        DYNAX_CUDA_TRY(cudaGetLastError());
    }
    return DYNAX_OK;
}

// Forward-mode (tangent) fused loss + gradient of the 4R3C model (rc_fwdsens.cuh).
template <int TN, int KF, int S, int MINB, int SPLIT = 1, bool SU = false, bool STG = false, bool CTRL = false,
          bool GX0 = false, bool CHUNK = false, int UNR = KF, bool SHIFT = false>
int launch_fwdsens(const FwdSensParams &P0, cudaStream_t st) {
    using Cfg = FwdSensCfg<TN, KF, S, SPLIT, SU, STG, CTRL, GX0, SHIFT>;
    auto kern = rc_fwdsens_kernel<TN, KF, S, MINB, SPLIT, SU, STG, CTRL, GX0, CHUNK, UNR, SHIFT>;
    if (SHIFT && P0.bulk) return fail(DYNAX_ERR_INVALID_ARGUMENT, "shifted-row instantiation launched on rows TMA can move");
    DYNAX_NOTE_LAUNCH();
    int rc = set_smem(kern, Cfg::SMEM_BYTES);
    if (rc != DYNAX_OK) return rc;
    // Per-system rows through tensor maps (one copy per array and stage) whenever the layout allows it: rows that 1-D
    // bulk copies can move (P0.bulk) and an inner extent of >= 32 floats dividing N*w and TN*w.  DYNAX_NO_TENSOR=1
    // keeps the per-row copies (A/B timing, tests).
    FwdSensParams P = P0;
    P.tensor = 0;
    if (P.bulk && (!SU || !STG || CTRL) && !getenv("DYNAX_NO_TENSOR")) {
        const int iu = SU ? 1 : row_map_inner(P.N, 5, TN), it = (STG && !CTRL) ? 1 : row_map_inner(P.N, 1, TN);
        if (iu > 0 && it > 0) {
            if (!SU) rc = make_row_map(&P.map_u, P.u, P.N, 5, P.T, TN, KF, iu);
            if (rc == DYNAX_OK && !STG) rc = make_row_map(&P.map_t, P.target, P.N, 1, P.T, TN, KF, it);
            if (rc == DYNAX_OK && CTRL) rc = make_row_map(&P.map_c, P.ctrl, P.N, 1, P.T, TN, KF, it);
            if (rc != DYNAX_OK) return rc;
            P.tensor = 1; P.inner_u = iu; P.inner_t = it;
        }
    }
    const int64_t grid = (P.N + TN - 1) / TN;
    if (grid > 0x7fffffffLL) return fail(DYNAX_ERR_INVALID_ARGUMENT, "N too large for one launch");
    if (P.T > 0x7fffffffLL) return fail(DYNAX_ERR_INVALID_ARGUMENT, "T too large");
    if (P.shared_acc) DYNAX_CUDA_TRY(cudaMemsetAsync(P.shared_acc, 0, kAccBytes, st));
    if constexpr (CHUNK) {  // time-chunked launch: the caller (lti_chunked.cu) owns the stitch and the outputs
        const int64_t chunks = (P.T + P.chunk_len - 1) / P.chunk_len;
        if (chunks > 65535) return fail(DYNAX_ERR_INVALID_ARGUMENT, "too many time chunks");
        kern<<<dim3((unsigned)grid, (unsigned)chunks), Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(P);
        DYNAX_CUDA_TRY(cudaGetLastError());
        return DYNAX_OK;
    }
    kern<<<(unsigned)grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(P);
    DYNAX_CUDA_TRY(cudaGetLastError());
    if (P.shared_acc) {
        RCModel::GradPtrs gp;
        gp.g_theta = P.g_theta;
        shared_finalize_kernel<RCModel, ADJ_MSE><<<1, 32, 0, st>>>(P.shared_acc, P.mp, gp, P.lb, P.ub, P.loss);
        DYNAX_CUDA_TRY(cudaGetLastError());
    }
    return DYNAX_OK;
}


}  // namespace dynax
