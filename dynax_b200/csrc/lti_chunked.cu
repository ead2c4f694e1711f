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
#include <stdlib.h>
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
    // the serial recurrence is short; what costs is one dependent global round trip per chunk, so the zero-state
    // responses are fetched G chunks at a time (N=1, T=8760 has 136 chunks: 9 round trips instead of 136)
    constexpr int G = 16;
    for (int64_t c0 = 0; c0 < C; c0 += G) {
        float vv[G][n];
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
            for (int d = 0; d < n; ++d) vv[g][d] = (c0 + g + 1 < C) ? __ldg(v + ((c0 + g) * N + i) * n + d) : 0.f;
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const int64_t c = c0 + g;
            if (c >= C) break;
#pragma unroll
            for (int d = 0; d < n; ++d) xb[(c * N + i) * n + d] = x[d];
            if (c + 1 == C) break;
            float y[n];
#pragma unroll
            for (int d = 0; d < n; ++d) {
                float acc = vv[g][d];
#pragma unroll
                for (int j = 0; j < n; ++j) acc = fmaf(P[j][d], x[j], acc);
                y[d] = x[d] + acc;
            }
#pragma unroll
            for (int d = 0; d < n; ++d) x[d] = y[d];
        }
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
    // like stitch_kernel: one dependent global round trip per GROUP of chunks, not per chunk
    constexpr int G = 16;
    for (int64_t c0 = C - 1; c0 >= 0; c0 -= G) {
        float wv[G][n];
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
            for (int j = 0; j < n; ++j) wv[g][j] = (c0 - g >= 1) ? __ldg(w + ((c0 - g) * N + i) * n + j) : 0.f;
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const int64_t c = c0 - g;
            if (c < 0) break;
#pragma unroll
            for (int d = 0; d < n; ++d) aend[(c * N + i) * n + d] = a[d];
            if (c == 0) break;
            float b[n];
#pragma unroll
            for (int j = 0; j < n; ++j) {
                float acc = wv[g][j];
#pragma unroll
                for (int d = 0; d < n; ++d) acc = fmaf(P[j][d], a[d], acc);  // (Psi^T a)_j
                b[j] = a[j] + acc;
            }
#pragma unroll
            for (int j = 0; j < n; ++j) a[j] = b[j];
        }
    }
}

// Tangent path: assemble loss and gradient from the chunks' local results.  With S_c = dx/dtheta at the start of
// chunk c (S_0 = 0) and X_c = dx/dx0 there (X_0 = I):
//     grad  += g_loc[c] + w[c]^T S_c          g_x0 += w[c]^T X_c
//     S_c+1  = S_c + Psi S_c + s_end[c]       X_c+1 = X_c + Psi X_c           (Psi = (I + dt A)^L - I, full chunks)
// The ten columns (7 of S, 3 of X) evolve independently: one thread per (column, system), columns outermost so that a
// warp holds 32 systems of one column (coalesced records, warp-level sums for shared theta).  Columns in fp32 like
// the serial kernel's S, sums in fp64; records fetched G chunks at a time (the recurrence is short, the dependent
// global round trips are what cost).  Thread (column 0) also owns the loss.
__global__ void tangent_stitch_kernel(const float *psi, const float *s_end, const double *g_loc, const float *w,
                                      RCModel::Ptrs mp, const float *lb, const float *ub, float *loss, float *g_theta,
                                      float *g_x0, double *shared_acc, int64_t N, int64_t C, int64_t T) {
    const int64_t N32 = (N + 31) / 32 * 32;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int j = (int)(t / N32);  // column: 0..6 = theta, 7..9 = x0
    const int64_t i = t % N32;
    const bool active = i < N && j < 10;
    const int64_t ii = i < N ? i : N - 1;
    const int jj = j < 10 ? j : 9;
    float Ps[3][3];  // Ps[d][e] = Psi[d][e]   (psi is stored column-major per system: psi[(i*3 + e)*3 + d])
#pragma unroll
    for (int e = 0; e < 3; ++e)
#pragma unroll
        for (int d = 0; d < 3; ++d) Ps[d][e] = __ldg(psi + (ii * 3 + e) * 3 + d);
    float v[3] = {jj == 7 ? 1.f : 0.f, jj == 8 ? 1.f : 0.f, jj == 9 ? 1.f : 0.f};
    double g = 0.0, sq = 0.0;
    constexpr int G = 8;  // 16 costs 200 registers and is slower from N=128 up (99 vs 95 us at T=8760)
    for (int64_t c0 = 0; c0 < C; c0 += G) {
        float wv[G][3], se[G][3];
        double gl[G], sl[G];
#pragma unroll
        for (int q = 0; q < G; ++q) {
            const int64_t o = ((c0 + q < C) ? (c0 + q) : (C - 1)) * N + ii;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                wv[q][d] = __ldg(w + o * 3 + d);
                se[q][d] = jj < 7 ? __ldg(s_end + o * 21 + d * 7 + jj) : 0.f;
            }
            gl[q] = jj < 7 ? __ldg(g_loc + o * 8 + jj) : 0.0;
            sl[q] = jj == 0 ? __ldg(g_loc + o * 8 + 7) : 0.0;
        }
#pragma unroll
        for (int q = 0; q < G; ++q) {
            if (c0 + q >= C) break;
            g += gl[q] + ((double)wv[q][0] * (double)v[0] + (double)wv[q][1] * (double)v[1] + (double)wv[q][2] * (double)v[2]);
            sq += sl[q];
            float vn[3];
#pragma unroll
            for (int d = 0; d < 3; ++d)
                vn[d] = v[d] + (fmaf(Ps[d][2], v[2], fmaf(Ps[d][1], v[1], Ps[d][0] * v[0])) + se[q][d]);
#pragma unroll
            for (int d = 0; d < 3; ++d) v[d] = vn[d];
        }
    }
    if (jj >= 7) {  // d loss / d x0
        if (active && g_x0 != nullptr) g_x0[i * 3 + (jj - 7)] = (float)g;
        return;
    }
    double L = sq / (double)T;  // meaningful in column 0 only
    if (shared_acc != nullptr) {  // ONE parameter vector: sum over systems; penalty added once by shared_finalize_kernel
        const int lane = threadIdx.x & 31;
        const double vs = warp_sum(active ? g : 0.0);
        if (lane == 0) atomicAdd(shared_acc + jj, vs);
        if (jj == 0) {
            const double ls = warp_sum(active ? L : 0.0);
            if (lane == 0) atomicAdd(shared_acc + 7, ls);
        }
        return;
    }
    if (!active) return;
    if (lb != nullptr && ub != nullptr) {
        const float *th = mp.theta + i * mp.stride;
        {  // 3-inverse.py:103-107, normaliser = ub
            const double tv = (double)th[jj], lo = (double)lb[jj], hi = (double)ub[jj];
            if (tv > hi) g += 1.0 / hi;
            if (tv < lo) g -= 1.0 / hi;
        }
        if (jj == 0) {
#pragma unroll
            for (int q = 0; q < 7; ++q) {
                const double tv = (double)th[q], lo = (double)lb[q], hi = (double)ub[q];
                if (tv > hi) L += (tv - hi) / hi;
                if (tv < lo) L += (lo - tv) / hi;
            }
        }
    }
    if (jj == 0) loss[i] = (float)L;
    if (g_theta != nullptr) g_theta[i * 7 + jj] = (float)g;   // (NULL: loss only, as dynax_rc_loss_grad_f32)
}

struct ChunkedAdjWs {
    float *psi, *v, *xb, *w, *aend, *ckpt;
    double *chunk_acc, *shared_acc;
    float *s_end;   // tangent path: [C,N,21]
    double *g_loc;  // tangent path: [C,N,8]
    size_t bytes;
};

inline ChunkedAdjWs chunked_adj_layout(void *ws, int64_t N, int64_t T, int n, int nout, int64_t L,
                                       bool tangent_bufs = false) {
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
    o.s_end = tangent_bufs ? reinterpret_cast<float *>(take(sizeof(float) * (size_t)C * N * n * nout)) : nullptr;
    o.g_loc = tangent_bufs ? reinterpret_cast<double *>(take(sizeof(double) * (size_t)C * N * (nout + 1))) : nullptr;
    o.bytes = (size_t)(p - reinterpret_cast<char *>(ws));
    return o;
}

// Time-parallel general-cotangent VJP, any linear block model (see dynax_rc_vjp_chunked_f32 below).
template <class Model, int MINB>
int vjp_chunked(typename Model::Ptrs mp, typename Model::GradPtrs gp, const float *x0, const float *u,
                const float *g_xsol, const float *g_ysol, float *g_x0, float *g_u, int64_t N, int64_t T, float dt,
                bool sth, int64_t chunk_len, void *ws, size_t ws_bytes, cudaStream_t st) {
    constexpr int n = Model::n, NOUT = Model::NOUT;
    int rc = require_sm100();
    if (rc != DYNAX_OK) return rc;
    const int64_t L = chunk_len > 0 ? (chunk_len < T ? chunk_len : T) : auto_chunk_len(N, T, 128);
    const int64_t C = (T + L - 1) / L;
    if (!ws || !aligned16(ws)) return fail(DYNAX_ERR_WORKSPACE, "workspace must be non-NULL and 16-byte aligned");
    const ChunkedAdjWs W = chunked_adj_layout(ws, N, T, n, NOUT, L);
    if (ws_bytes < W.bytes) return fail(DYNAX_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", W.bytes, ws_bytes);
    const bool bulk = (N % 4) == 0 && aligned16(u) && (!g_xsol || aligned16(g_xsol)) && (!g_ysol || aligned16(g_ysol));
    const unsigned tb = 128, gb = (unsigned)((N + tb - 1) / tb);

    // ---- forward: chunk boundary states xb[c]  (psi, zero-state responses, stitch)
    FwdParams<Model> F;
    F.mp = mp;
    F.x0 = x0; F.u = u; F.xsol = nullptr; F.ysol = nullptr; F.x_final = W.v;
    F.N = N; F.T = T; F.dt = dt; F.discrete = 0; F.bulk_in = (N % 4) == 0 && aligned16(u); F.bulk_out = 0;
    F.chunk_len = L; F.x0_zero = 1;
    psi_kernel<Model><<<gb, tb, 0, st>>>(mp, W.psi, N, L, dt);
    DYNAX_CUDA_TRY(cudaGetLastError());
    rc = launch_fwd<Model, 128, 8, 2, false, 2>(F, st);
    if (rc != DYNAX_OK) return rc;
    stitch_kernel<n><<<gb, tb, 0, st>>>(W.psi, x0, W.v, W.xb, N, C);
    DYNAX_CUDA_TRY(cudaGetLastError());

    // ---- adjoint pass B1 (zero terminal adjoint, no sums) -> w[c]; backward stitch -> aend[c]; pass B2 (true terminal
    //      adjoints): gradient sums per system in fp64 over the chunks, g_u in place
    AdjParams<Model> A;
    memset(&A, 0, sizeof(A));
    A.mp = mp; A.gp = gp;
    A.x0 = W.xb; A.x0_chunk_stride = N * n;
    A.u = u; A.g_xsol = g_xsol; A.g_ysol = g_ysol; A.g_u = g_u;
    A.ckpt = W.ckpt; A.ckpt_chunk_stride = ((L + kMinKC - 1) / kMinKC) * n * N;
    A.shared_acc = W.shared_acc;
    A.N = N; A.T = T; A.dt = dt; A.bulk = bulk; A.shared_theta = sth ? 1 : 0; A.chunk_len = L;
    AdjParams<Model> B1 = A;
    B1.a_start = W.w; B1.skip_grads = 1;
    rc = launch_adj<Model, 128, 8, 2, ADJ_VJP, false, false, MINB>(B1, st);
    if (rc != DYNAX_OK) return rc;
    stitch_adjoint_kernel<n><<<gb, tb, 0, st>>>(W.psi, W.w, W.aend, N, C);
    DYNAX_CUDA_TRY(cudaGetLastError());
    DYNAX_CUDA_TRY(cudaMemsetAsync(W.chunk_acc, 0, sizeof(double) * (size_t)N * (NOUT + 1), st));
    if (sth) DYNAX_CUDA_TRY(cudaMemsetAsync(W.shared_acc, 0, kAccBytes, st));
    AdjParams<Model> B2 = A;
    B2.a_term = W.aend; B2.chunk_acc = W.chunk_acc; B2.g_x0 = g_x0;
    rc = launch_adj<Model, 128, 8, 2, ADJ_VJP, false, false, MINB>(B2, st);
    if (rc != DYNAX_OK) return rc;
    chunk_finalize_kernel<Model><<<gb, tb, 0, st>>>(W.chunk_acc, mp, gp, nullptr, nullptr, nullptr, W.shared_acc, N, sth ? 1 : 0);
    DYNAX_CUDA_TRY(cudaGetLastError());
    if (sth) {
        shared_finalize_kernel<Model, ADJ_VJP><<<1, 32, 0, st>>>(W.shared_acc, mp, gp, nullptr, nullptr, nullptr);
        DYNAX_CUDA_TRY(cudaGetLastError());
    }
    return DYNAX_OK;
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
    DYNAX_API_RANGE();
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
        return launch_fwd<RCModel, 128, 8, 2, false, 3>(Q, s);  // 8-row chunks: fewer hand-shakes on the serial chain (16 rows: 2 CTAs/SM, slower here)
    });
}

// Dense linear block SSMs (Discrete/ContinuousLinearSSM stepped by the simulator): the same time-parallel forward.
size_t dynax_lti_forward_chunked_workspace(int64_t N, int64_t T, int n, int64_t chunk_len) {
    if (N <= 0 || T <= 0 || n <= 0) return 0;
    const int64_t L = chunk_len > 0 ? chunk_len : auto_chunk_len(N, T, 128);
    return chunked_ws_bytes(N, T, n, L);
}

int dynax_lti_forward_chunked_f32(const float *A, const float *B, const float *C, const float *D, const float *x0,
                                  const float *u, float *xsol, float *ysol, float *x_final, int64_t N, int64_t T, int n,
                                  int m, int p, float dt, uint32_t flags, int64_t chunk_len, void *ws, size_t ws_bytes,
                                  void *stream) {
    DYNAX_API_RANGE();
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (N < 0 || T < 0) return fail(DYNAX_ERR_INVALID_ARGUMENT, "negative N or T");
    if (N == 0) return DYNAX_OK;
    if (!A || !B || !C || !D || !x0 || (!u && T > 0)) return fail(DYNAX_ERR_INVALID_ARGUMENT, "A,B,C,D,x0,u must be non-NULL");
    if (flags & ~(DYNAX_SHARED_THETA | DYNAX_SHARED_U))
        return fail(DYNAX_ERR_INVALID_ARGUMENT, "unsupported flags 0x%x for lti_forward_chunked (Euler residual form only)", flags);
    if (!dynax_lti_supported(n, m, p))
        return fail(DYNAX_ERR_UNSUPPORTED_DIMS, "no dense-LTI kernel instantiated for (n,m,p)=(%d,%d,%d)", n, m, p);
    int rc = require_sm100();
    if (rc != DYNAX_OK) return rc;
    if (T == 0) {
        if (x_final) DYNAX_CUDA_TRY(cudaMemcpyAsync(x_final, x0, sizeof(float) * n * N, cudaMemcpyDeviceToDevice, st));
        return DYNAX_OK;
    }
    if (!ws || !aligned16(ws)) return fail(DYNAX_ERR_WORKSPACE, "workspace must be non-NULL and 16-byte aligned");
    const int64_t L = chunk_len > 0 ? (chunk_len < T ? chunk_len : T) : auto_chunk_len(N, T, 128);
    const bool rows16 = (N % 4) == 0;
    const int bulk_in = rows16 && aligned16(u);
    const int bulk_out = rows16 && (!xsol || aligned16(xsol)) && (!ysol || aligned16(ysol));
#define X(a, b, c)                                                                                          \
    if (n == a && m == b && p == c) {                                                                       \
        using M = DenseModel<a, b, c>;                                                                      \
        FwdParams<M> P;                                                                                     \
        P.mp.A = A; P.mp.B = B; P.mp.C = C; P.mp.D = D;                                                     \
        P.mp.shared = (flags & DYNAX_SHARED_THETA) ? 1 : 0;                                                 \
        P.x0 = x0; P.u = u; P.xsol = xsol; P.ysol = ysol; P.x_final = x_final;                              \
        P.N = N; P.T = T; P.dt = dt; P.discrete = 0; P.bulk_in = bulk_in; P.bulk_out = bulk_out;            \
        if (flags & DYNAX_SHARED_U)                                                                         \
            return forward_chunked<M>(P, L, ws, ws_bytes, st, [](const FwdParams<M> &Q, cudaStream_t s) {   \
                return launch_fwd<M, 128, 4, 2, true, 2>(Q, s);                                             \
            });                                                                                             \
        return forward_chunked<M>(P, L, ws, ws_bytes, st, [](const FwdParams<M> &Q, cudaStream_t s) {       \
            return launch_fwd<M, 128, 8, 2, false, 2>(Q, s);                                                \
        });                                                                                                 \
    }
    DYNAX_LTI_DIMS(X)
#undef X
    return fail(DYNAX_ERR_UNSUPPORTED_DIMS, "unreachable");
}

// Time-parallel general-cotangent VJP of the RC rollout (what jax.grad of a user loss on the simulator outputs needs)
// for launches with few systems and a long horizon: chunk boundary states as above; every chunk's adjoint for a ZERO
// terminal adjoint (no gradient sums) -> w[c]; backward stitch a_end[c-1] = w[c] + a_end[c] + Psi^T a_end[c]; every
// chunk's adjoint again from its true terminal adjoint, gradient sums added per system in fp64, g_u written in place.
size_t dynax_rc_vjp_chunked_workspace(int64_t N, int64_t T, int64_t chunk_len) {
    if (N <= 0 || T <= 0) return 0;
    const int64_t L = chunk_len > 0 ? (chunk_len < T ? chunk_len : T) : auto_chunk_len(N, T, 128);
    return chunked_adj_layout(nullptr, N, T, 3, 7, L).bytes;
}

int dynax_rc_vjp_chunked_f32(const float *theta, const float *x0, const float *u, const float *g_xsol,
                             const float *g_ysol, float *g_theta, float *g_x0, float *g_u, int64_t N, int64_t T,
                             float dt, uint32_t flags, int64_t chunk_len, void *ws, size_t ws_bytes, void *stream) {
    DYNAX_API_RANGE();
    if (N < 0 || T < 1) return fail(DYNAX_ERR_INVALID_ARGUMENT, "need N >= 0 and T >= 1");
    if (N == 0) return DYNAX_OK;
    if (!theta || !x0 || !u) return fail(DYNAX_ERR_INVALID_ARGUMENT, "theta, x0 and u must be non-NULL");
    if (!g_xsol && !g_ysol) return fail(DYNAX_ERR_INVALID_ARGUMENT, "need g_xsol or g_ysol");
    if (flags & ~DYNAX_SHARED_THETA)
        return fail(DYNAX_ERR_INVALID_ARGUMENT, "unsupported flags 0x%x for rc_vjp_chunked (per-system u, Euler residual form)", flags);
    RCModel::Ptrs mp;
    mp.theta = theta; mp.stride = (flags & DYNAX_SHARED_THETA) ? 0 : 7;
    RCModel::GradPtrs gp;
    gp.g_theta = g_theta;
    return vjp_chunked<RCModel, 2>(mp, gp, x0, u, g_xsol, g_ysol, g_x0, g_u, N, T, dt, (flags & DYNAX_SHARED_THETA) != 0,
                                   chunk_len, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}

size_t dynax_lti_vjp_chunked_workspace(int64_t N, int64_t T, int n, int m, int p, int64_t chunk_len) {
    if (N <= 0 || T <= 0 || n <= 0) return 0;
    const int64_t L = chunk_len > 0 ? (chunk_len < T ? chunk_len : T) : auto_chunk_len(N, T, 128);
    return chunked_adj_layout(nullptr, N, T, n, n * n + n * m + p * n + p * m, L).bytes;
}

int dynax_lti_vjp_chunked_f32(const float *A, const float *B, const float *C, const float *D, const float *x0,
                              const float *u, const float *g_xsol, const float *g_ysol, float *g_A, float *g_B,
                              float *g_C, float *g_D, float *g_x0, float *g_u, int64_t N, int64_t T, int n, int m,
                              int p, float dt, uint32_t flags, int64_t chunk_len, void *ws, size_t ws_bytes,
                              void *stream) {
    DYNAX_API_RANGE();
    if (N < 0 || T < 1) return fail(DYNAX_ERR_INVALID_ARGUMENT, "need N >= 0 and T >= 1");
    if (N == 0) return DYNAX_OK;
    if (!A || !B || !C || !D || !x0 || !u) return fail(DYNAX_ERR_INVALID_ARGUMENT, "A,B,C,D,x0,u must be non-NULL");
    if (!g_xsol && !g_ysol) return fail(DYNAX_ERR_INVALID_ARGUMENT, "need g_xsol or g_ysol");
    if (flags & ~DYNAX_SHARED_THETA)
        return fail(DYNAX_ERR_INVALID_ARGUMENT, "unsupported flags 0x%x for lti_vjp_chunked (per-system u, Euler residual form)", flags);
    if (!dynax_lti_supported(n, m, p))
        return fail(DYNAX_ERR_UNSUPPORTED_DIMS, "no dense-LTI kernel instantiated for (n,m,p)=(%d,%d,%d)", n, m, p);
#define X(a, b, c)                                                                                              \
    if (n == a && m == b && p == c) {                                                                           \
        using M = DenseModel<a, b, c>;                                                                          \
        M::Ptrs mp;                                                                                             \
        mp.A = A; mp.B = B; mp.C = C; mp.D = D; mp.shared = (flags & DYNAX_SHARED_THETA) ? 1 : 0;               \
        M::GradPtrs gp;                                                                                         \
        gp.gA = g_A; gp.gB = g_B; gp.gC = g_C; gp.gD = g_D;                                                     \
        return vjp_chunked<M, 1>(mp, gp, x0, u, g_xsol, g_ysol, g_x0, g_u, N, T, dt,                            \
                                 (flags & DYNAX_SHARED_THETA) != 0, chunk_len, ws, ws_bytes,                    \
                                 static_cast<cudaStream_t>(stream));                                            \
    }
    DYNAX_LTI_DIMS(X)
#undef X
    return fail(DYNAX_ERR_UNSUPPORTED_DIMS, "unreachable");
}

size_t dynax_rc_loss_grad_chunked_workspace(int64_t N, int64_t T, int64_t chunk_len) {
    if (N <= 0 || T <= 0) return 0;
    const int64_t L = chunk_len > 0 ? (chunk_len < T ? chunk_len : T) : auto_chunk_len(N, T, 128);
    return chunked_adj_layout(nullptr, N, T, 3, 7, L, true).bytes;
}

int dynax_rc_loss_grad_chunked_f32(const float *theta, const float *x0, const float *u, const float *target,
                                   const float *lb, const float *ub, float *loss, float *g_theta, float *g_x0,
                                   int64_t N, int64_t T, float dt, uint32_t flags, int64_t chunk_len, void *ws,
                                   size_t ws_bytes, void *stream) {
    DYNAX_API_RANGE();
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
    const ChunkedAdjWs W = chunked_adj_layout(ws, N, T, 3, 7, L, true);
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
    rc = su ? launch_fwd<RCModel, 128, 4, 4, true, 4>(F, st) : launch_fwd<RCModel, 128, 8, 2, false, 3>(F, st);
    if (rc != DYNAX_OK) return rc;
    stitch_kernel<3><<<gb, tb, 0, st>>>(W.psi, x0, W.v, W.xb, N, C);
    DYNAX_CUDA_TRY(cudaGetLastError());

    // ---- gradient, default: ONE forward-mode (tangent) pass over the chunks from their true boundary states with
    //      S = 0, then a stitch over chunks (tangent_stitch_kernel).  u is read twice in total (adjoint path below: 5x).
    {
        const char *algo = getenv("DYNAX_GRAD_ALGO");
        if (!(algo && algo[0] == '0')) {
            FwdSensParams Q;
            memset(&Q, 0, sizeof(Q));
            Q.mp = F.mp; Q.x0 = W.xb; Q.x0_chunk_stride = N * 3; Q.u = u; Q.target = target;
            Q.N = N; Q.T = T; Q.dt = dt; Q.chunk_len = L;
            Q.bulk = (N % 4) == 0 && (su || aligned16(u)) && (stg || aligned16(target));
            Q.s_end = W.s_end; Q.g_loc = W.g_loc; Q.w_out = W.w;
            if (su && stg) rc = launch_fwdsens<128, 8, 2, 4, 1, true, true, false, false, true>(Q, st);
            else if (su) rc = launch_fwdsens<128, 8, 2, 4, 1, true, false, false, false, true>(Q, st);
            else if (stg) rc = launch_fwdsens<128, 8, 2, 4, 1, false, true, false, false, true>(Q, st);
            else rc = launch_fwdsens<128, 8, 2, 4, 1, false, false, false, false, true>(Q, st);
            if (rc != DYNAX_OK) return rc;
            if (sth) DYNAX_CUDA_TRY(cudaMemsetAsync(W.shared_acc, 0, kAccBytes, st));
            const int64_t threads = 10 * ((N + 31) / 32 * 32);
            tangent_stitch_kernel<<<(unsigned)((threads + tb - 1) / tb), tb, 0, st>>>(
                W.psi, W.s_end, W.g_loc, W.w, F.mp, lb, ub, loss, g_theta, g_x0, sth ? W.shared_acc : nullptr, N, C, T);
            DYNAX_CUDA_TRY(cudaGetLastError());
            if (sth) {
                RCModel::GradPtrs gp;
                gp.g_theta = g_theta;
                shared_finalize_kernel<RCModel, ADJ_MSE><<<1, 32, 0, st>>>(W.shared_acc, F.mp, gp, lb, ub, loss);
                DYNAX_CUDA_TRY(cudaGetLastError());
            }
            return DYNAX_OK;
        }
    }

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
