// Time-parallel ("chunked associative scan") forward rollout for linear block SSMs.
//
// The nn.scan of /root/reference/dynax/simulators/simulator.py:36-53 is serial in time, but for a
// linear model  x_{k+1} = x_k + dt (A x_k + B u_k)  is affine in x, so the horizon can be cut into
// chunks of L steps that are rolled IN PARALLEL and stitched (validated recipe):
//
//   psi      per system  Psi_L = (I + dt A)^L - I   by L residual steps  Psi <- Psi + dt A (I + Psi)
//                        (I + dt A is never formed: SURVEY finding #8)
//   pass 1   every chunk from a ZERO initial state            -> v_c   (rollout_fwd_kernel, gridDim.y = chunks)
//   stitch   serial over the T/L chunk boundaries             x_{c+1} = x_c + Psi_L x_c + v_c
//   pass 2   every chunk from its true boundary state, emitting xsol / ysol
//
// u is read twice, so this is for launches that cannot fill the machine with systems alone
// (N per GPU below ~one wave: e.g. config 3 sharded over 8 GPUs, or a single building over a year).
#include <string.h>

#include "launch.cuh"

namespace dynax {

template <class Model>
__global__ void psi_kernel(typename Model::Ptrs mp, float *psi /*[N,n,n] column-major per system*/, int64_t N,
                           int64_t L, float dt) {
    constexpr int n = Model::n, m = Model::m;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    Model M;
    M.load(mp, i);
    float P[n][n];  // P[j] = column j of Psi
    float zero_u[m];
#pragma unroll
    for (int d = 0; d < m; ++d) zero_u[d] = 0.f;
#pragma unroll
    for (int j = 0; j < n; ++j)
#pragma unroll
        for (int d = 0; d < n; ++d) P[j][d] = 0.f;
    for (int64_t k = 0; k < L; ++k) {
#pragma unroll
        for (int j = 0; j < n; ++j) {
            float t[n], r[n];
#pragma unroll
            for (int d = 0; d < n; ++d) t[d] = P[j][d] + (d == j ? 1.f : 0.f);
            M.rhs(t, zero_u, r);  // A (e_j + Psi_j)   (B*0 contributes exact zeros)
#pragma unroll
            for (int d = 0; d < n; ++d) P[j][d] = fmaf(dt, r[d], P[j][d]);
        }
    }
#pragma unroll
    for (int j = 0; j < n; ++j)
#pragma unroll
        for (int d = 0; d < n; ++d) psi[(i * n + j) * n + d] = P[j][d];
}

// xb[0] = x0;  xb[c+1] = xb[c] + Psi xb[c] + v[c]
template <int n>
__global__ void stitch_kernel(const float *psi, const float *x0, const float *v /*[C,N,n]*/, float *xb /*[C,N,n]*/,
                              int64_t N, int64_t C) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float P[n][n], x[n];
#pragma unroll
    for (int j = 0; j < n; ++j)
#pragma unroll
        for (int d = 0; d < n; ++d) P[j][d] = psi[(i * n + j) * n + d];
#pragma unroll
    for (int d = 0; d < n; ++d) x[d] = x0[i * n + d];
    for (int64_t c = 0; c < C; ++c) {
#pragma unroll
        for (int d = 0; d < n; ++d) xb[(c * N + i) * n + d] = x[d];
        if (c + 1 == C) break;
        float y[n];
#pragma unroll
        for (int d = 0; d < n; ++d) {
            float acc = v[(c * N + i) * n + d];
#pragma unroll
            for (int j = 0; j < n; ++j) acc = fmaf(P[j][d], x[j], acc);
            y[d] = x[d] + acc;
        }
#pragma unroll
        for (int d = 0; d < n; ++d) x[d] = y[d];
    }
}

inline int64_t auto_chunk_len(int64_t N, int64_t T, int tile) {
    const int64_t blocks = (N + tile - 1) / tile;
    int64_t chunks = (2 * 148 + blocks - 1) / blocks;  // aim at two CTAs per SM
    if (chunks > T / 64) chunks = T / 64;              // keep chunks long enough to amortise the stitching
    if (chunks < 1) chunks = 1;
    return (T + chunks - 1) / chunks;
}

inline size_t chunked_ws_bytes(int64_t N, int64_t T, int n, int64_t L) {
    const int64_t C = (T + L - 1) / L;
    return sizeof(float) * ((size_t)N * n * n + 2 * (size_t)C * N * n) + 64;
}

template <class Model, class Launch>
int forward_chunked(FwdParams<Model> P, int64_t L, void *ws, size_t ws_bytes, cudaStream_t st, Launch launch) {
    constexpr int n = Model::n;
    const int64_t N = P.N, T = P.T;
    const int64_t C = (T + L - 1) / L;
    const size_t need = chunked_ws_bytes(N, T, n, L);
    if (!ws || ws_bytes < need)
        return fail(DYNAX_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", need, ws ? ws_bytes : (size_t)0);
    float *psi = reinterpret_cast<float *>(ws);
    float *v = psi + ((N * n * n + 3) & ~(int64_t)3);
    float *xb = v + C * N * n;
    const unsigned tb = 128, gb = (unsigned)((N + tb - 1) / tb);
    psi_kernel<Model><<<gb, tb, 0, st>>>(P.mp, psi, N, L, P.dt);
    DYNAX_CUDA_TRY(cudaGetLastError());
    // pass 1: zero-state response of every chunk (nothing emitted)
    FwdParams<Model> P1 = P;
    P1.xsol = nullptr; P1.ysol = nullptr; P1.x_final = v;
    P1.chunk_len = L; P1.x0_zero = 1; P1.x0_chunk_stride = 0;
    int rc = launch(P1, st);
    if (rc != DYNAX_OK) return rc;
    stitch_kernel<n><<<gb, tb, 0, st>>>(psi, P.x0, v, xb, N, C);
    DYNAX_CUDA_TRY(cudaGetLastError());
    // pass 2: every chunk from its boundary state, emitting outputs; v is reused for the chunk-end states
    FwdParams<Model> P2 = P;
    P2.x0 = xb; P2.x0_chunk_stride = N * n; P2.x0_zero = 0; P2.chunk_len = L; P2.x_final = v;
    rc = launch(P2, st);
    if (rc != DYNAX_OK) return rc;
    if (P.x_final)
        DYNAX_CUDA_TRY(cudaMemcpyAsync(P.x_final, v + (C - 1) * N * n, sizeof(float) * N * n, cudaMemcpyDeviceToDevice, st));
    return DYNAX_OK;
}

// a_end[C-1] = 0;  a_end[c-1] = a_start_true[c] = w[c] + a_end[c] + Psi^T a_end[c]      (c = C-1 .. 1)
// where w[c] is chunk c's start adjoint for a ZERO terminal adjoint.  ((I+dtA)^T)^L = (I+Psi_L)^T; the
// last chunk may be shorter but its terminal adjoint is zero, so only Psi_L of full chunks is needed.
template <int n>
__global__ void stitch_adjoint_kernel(const float *psi, const float *w /*[C,N,n]*/, float *aend /*[C,N,n]*/, int64_t N,
                                      int64_t C) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float P[n][n], a[n];
#pragma unroll
    for (int j = 0; j < n; ++j)
#pragma unroll
        for (int d = 0; d < n; ++d) P[j][d] = psi[(i * n + j) * n + d];  // P[j][d] = Psi[d][j]
#pragma unroll
    for (int d = 0; d < n; ++d) a[d] = 0.f;
    for (int64_t c = C - 1; c >= 0; --c) {
#pragma unroll
        for (int d = 0; d < n; ++d) aend[(c * N + i) * n + d] = a[d];
        if (c == 0) break;
        float b[n];
#pragma unroll
        for (int j = 0; j < n; ++j) {
            float acc = w[(c * N + i) * n + j];
#pragma unroll
            for (int d = 0; d < n; ++d) acc = fmaf(P[j][d], a[d], acc);  // (Psi^T a)_j
            b[j] = a[j] + acc;
        }
#pragma unroll
        for (int j = 0; j < n; ++j) a[j] = b[j];
    }
}

struct ChunkedAdjWs {
    float *psi, *v, *xb, *w, *aend, *ckpt;
    double *chunk_acc, *shared_acc;
    size_t bytes;
};

inline ChunkedAdjWs chunked_adj_layout(void *ws, int64_t N, int64_t T, int n, int nout, int64_t L) {
    const int64_t C = (T + L - 1) / L;
    const int64_t segs = (L + kMinKC - 1) / kMinKC;
    char *p = reinterpret_cast<char *>(ws);
    auto take = [&](size_t bytes) {
        char *r = p;
        p += (bytes + 255) & ~(size_t)255;
        return r;
    };
    ChunkedAdjWs o;
    o.shared_acc = reinterpret_cast<double *>(take(kAccBytes));
    o.chunk_acc = reinterpret_cast<double *>(take(sizeof(double) * (size_t)N * (nout + 1)));
    o.psi = reinterpret_cast<float *>(take(sizeof(float) * (size_t)N * n * n));
    o.v = reinterpret_cast<float *>(take(sizeof(float) * (size_t)C * N * n));
    o.xb = reinterpret_cast<float *>(take(sizeof(float) * (size_t)C * N * n));
    o.w = reinterpret_cast<float *>(take(sizeof(float) * (size_t)C * N * n));
    o.aend = reinterpret_cast<float *>(take(sizeof(float) * (size_t)C * N * n));
    o.ckpt = reinterpret_cast<float *>(take(sizeof(float) * (size_t)C * segs * n * N));
    o.bytes = (size_t)(p - reinterpret_cast<char *>(ws));
    return o;
}

}  // namespace dynax

using namespace dynax;

extern "C" {

size_t dynax_rc_forward_chunked_workspace(int64_t N, int64_t T, int64_t chunk_len) {
    if (N <= 0 || T <= 0) return 0;
    const int64_t L = chunk_len > 0 ? chunk_len : auto_chunk_len(N, T, 128);
    return chunked_ws_bytes(N, T, 3, L);
}

int dynax_rc_forward_chunked_f32(const float *theta, const float *x0, const float *u, float *xsol, float *ysol,
                                 float *x_final, int64_t N, int64_t T, float dt, uint32_t flags, int64_t chunk_len,
                                 void *ws, size_t ws_bytes, void *stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (N < 0 || T < 1) return fail(DYNAX_ERR_INVALID_ARGUMENT, "need N >= 0 and T >= 1");
    if (N == 0) return DYNAX_OK;
    if (!theta || !x0 || !u) return fail(DYNAX_ERR_INVALID_ARGUMENT, "theta, x0 and u must be non-NULL");
    if (flags & ~(DYNAX_SHARED_THETA | DYNAX_SHARED_U))
        return fail(DYNAX_ERR_INVALID_ARGUMENT, "unsupported flags 0x%x for rc_forward_chunked (Euler residual form only)", flags);
    int rc = require_sm100();
    if (rc != DYNAX_OK) return rc;
    const int64_t L = chunk_len > 0 ? (chunk_len < T ? chunk_len : T) : auto_chunk_len(N, T, 128);
    FwdParams<RCModel> P;
    P.mp.theta = theta;
    P.mp.stride = (flags & DYNAX_SHARED_THETA) ? 0 : 7;
    P.x0 = x0; P.u = u; P.xsol = xsol; P.ysol = ysol; P.x_final = x_final;
    P.N = N; P.T = T; P.dt = dt; P.discrete = 0;
    const bool rows16 = (N % 4) == 0;
    P.bulk_in = rows16 && aligned16(u);
    P.bulk_out = rows16 && (!xsol || aligned16(xsol)) && (!ysol || aligned16(ysol));
    if (flags & DYNAX_SHARED_U)
        return forward_chunked<RCModel>(P, L, ws, ws_bytes, st, [](const FwdParams<RCModel> &Q, cudaStream_t s) {
            return launch_fwd<RCModel, 128, 4, 4, true, 4>(Q, s);
        });
    return forward_chunked<RCModel>(P, L, ws, ws_bytes, st, [](const FwdParams<RCModel> &Q, cudaStream_t s) {
        return launch_fwd<RCModel, 128, 4, 4, false, 3>(Q, s);
    });
}

size_t dynax_rc_loss_grad_chunked_workspace(int64_t N, int64_t T, int64_t chunk_len) {
    if (N <= 0 || T <= 0) return 0;
    const int64_t L = chunk_len > 0 ? (chunk_len < T ? chunk_len : T) : auto_chunk_len(N, T, 128);
    return chunked_adj_layout(nullptr, N, T, 3, 7, L).bytes;
}

int dynax_rc_loss_grad_chunked_f32(const float *theta, const float *x0, const float *u, const float *target,
                                   const float *lb, const float *ub, float *loss, float *g_theta, float *g_x0,
                                   int64_t N, int64_t T, float dt, uint32_t flags, int64_t chunk_len, void *ws,
                                   size_t ws_bytes, void *stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (N < 0 || T < 1) return fail(DYNAX_ERR_INVALID_ARGUMENT, "need N >= 0 and T >= 1");
    if (N == 0) return DYNAX_OK;
    if (!theta || !x0 || !u || !target || !loss) return fail(DYNAX_ERR_INVALID_ARGUMENT, "theta, x0, u, target, loss must be non-NULL");
    if ((lb == nullptr) != (ub == nullptr)) return fail(DYNAX_ERR_INVALID_ARGUMENT, "lb and ub must both be given or both be NULL");
    if (flags & ~(DYNAX_SHARED_THETA | DYNAX_SHARED_U | DYNAX_SHARED_TARGET))
        return fail(DYNAX_ERR_INVALID_ARGUMENT, "unsupported flags 0x%x for rc_loss_grad_chunked", flags);
    int rc = require_sm100();
    if (rc != DYNAX_OK) return rc;
    const int64_t L = chunk_len > 0 ? (chunk_len < T ? chunk_len : T) : auto_chunk_len(N, T, 128);
    const int64_t C = (T + L - 1) / L;
    if (!ws || !aligned16(ws)) return fail(DYNAX_ERR_WORKSPACE, "workspace must be non-NULL and 16-byte aligned");
    const ChunkedAdjWs W = chunked_adj_layout(ws, N, T, 3, 7, L);
    if (ws_bytes < W.bytes) return fail(DYNAX_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", W.bytes, ws_bytes);
    const bool sth = flags & DYNAX_SHARED_THETA, su = flags & DYNAX_SHARED_U, stg = flags & DYNAX_SHARED_TARGET;
    const bool bulk = (N % 4) == 0 && aligned16(u) && aligned16(target);
    const unsigned tb = 128, gb = (unsigned)((N + tb - 1) / tb);

    // ---- forward: chunk boundary states xb[c]  (psi, zero-state responses, stitch)
    FwdParams<RCModel> F;
    F.mp.theta = theta; F.mp.stride = sth ? 0 : 7;
    F.x0 = x0; F.u = u; F.xsol = nullptr; F.ysol = nullptr; F.x_final = W.v;
    F.N = N; F.T = T; F.dt = dt; F.discrete = 0; F.bulk_in = bulk; F.bulk_out = 0;
    F.chunk_len = L; F.x0_zero = 1;
    psi_kernel<RCModel><<<gb, tb, 0, st>>>(F.mp, W.psi, N, L, dt);
    DYNAX_CUDA_TRY(cudaGetLastError());
    rc = su ? launch_fwd<RCModel, 128, 4, 4, true, 4>(F, st) : launch_fwd<RCModel, 128, 4, 4, false, 3>(F, st);
    if (rc != DYNAX_OK) return rc;
    stitch_kernel<3><<<gb, tb, 0, st>>>(W.psi, x0, W.v, W.xb, N, C);
    DYNAX_CUDA_TRY(cudaGetLastError());

    // ---- adjoint pass B1 (zero terminal adjoint, no gradient sums) -> w[c]; backward stitch -> aend[c];
    //      pass B2 (true terminal adjoints) -> per-system fp64 sums over chunks -> outputs
    AdjParams<RCModel> A;
    memset(&A, 0, sizeof(A));
    A.mp = F.mp; A.gp.g_theta = g_theta;
    A.x0 = W.xb; A.x0_chunk_stride = N * 3;
    A.u = u; A.target = target; A.lb = lb; A.ub = ub; A.loss = loss;
    A.ckpt = W.ckpt; A.ckpt_chunk_stride = ((L + kMinKC - 1) / kMinKC) * 3 * N;
    A.shared_acc = W.shared_acc; A.chunk_acc = nullptr;
    A.N = N; A.T = T; A.dt = dt; A.bulk = bulk; A.shared_theta = sth ? 1 : 0; A.chunk_len = L;
    auto launch = [&](const AdjParams<RCModel> &Q) {
        if (su && stg) return launch_adj<RCModel, 128, 16, 4, ADJ_MSE, true, true, 2>(Q, st);
        if (su) return launch_adj<RCModel, 128, 16, 4, ADJ_MSE, true, false, 2>(Q, st);
        if (stg) return launch_adj<RCModel, 128, 12, 2, ADJ_MSE, false, true, 3>(Q, st);
        return launch_adj<RCModel, 128, 12, 2, ADJ_MSE, false, false, 3>(Q, st);
    };
    AdjParams<RCModel> B1 = A;
    B1.a_term = nullptr; B1.a_start = W.w; B1.skip_grads = 1;
    rc = launch(B1);
    if (rc != DYNAX_OK) return rc;
    stitch_adjoint_kernel<3><<<gb, tb, 0, st>>>(W.psi, W.w, W.aend, N, C);
    DYNAX_CUDA_TRY(cudaGetLastError());
    DYNAX_CUDA_TRY(cudaMemsetAsync(W.chunk_acc, 0, sizeof(double) * (size_t)N * 8, st));
    if (sth) DYNAX_CUDA_TRY(cudaMemsetAsync(W.shared_acc, 0, kAccBytes, st));
    AdjParams<RCModel> B2 = A;
    B2.a_term = W.aend; B2.a_start = nullptr; B2.skip_grads = 0; B2.chunk_acc = W.chunk_acc; B2.g_x0 = g_x0;
    rc = launch(B2);
    if (rc != DYNAX_OK) return rc;
    chunk_finalize_kernel<RCModel><<<gb, tb, 0, st>>>(W.chunk_acc, A.mp, A.gp, lb, ub, loss, W.shared_acc, N, sth ? 1 : 0);
    DYNAX_CUDA_TRY(cudaGetLastError());
    if (sth) {
        shared_finalize_kernel<RCModel, ADJ_MSE><<<1, 32, 0, st>>>(W.shared_acc, A.mp, A.gp, lb, ub, loss);
        DYNAX_CUDA_TRY(cudaGetLastError());
    }
    return DYNAX_OK;
}

}  // extern "C"
