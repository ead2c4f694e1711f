// C-ABI entry points for dense linear block SSMs (DiscreteLinearSSM / ContinuousLinearSSM,
// /root/reference/dynax/core/{discrete,continuous}_block_state_space.py:5-54).
// Kernels are instantiated for a fixed list of (n,m,p); dynax_lti_supported() reports it.
#include <stdlib.h>
#include <string.h>

#include "launch.cuh"

// X(n, m, p): (4,3,4) is the shape of the reference's own test
// (tests/test_forward_simulation.py:17-20); (3,5,1) is the 4R3C shape with free matrices.

namespace dynax {

template <int n, int m, int p>
static int lti_forward_t(const FwdParams<DenseModel<n, m, p>> &P, bool shared_u, cudaStream_t st) {
    using M = DenseModel<n, m, p>;
    DeviceInfo di;
    const int rc = device_info(&di);
    if (rc != DYNAX_OK) return rc;
    // a shared input series on launches of at most two 128-system CTAs per SM: 16-row chunks, 4-stage ring (rc_api.cu)
    if (shared_u) {
        if (P.N <= (int64_t)di.sm_count * 256) return launch_fwd<M, 128, 16, 4, true, 2>(P, st);
        return launch_fwd<M, 128, 4, 2, true, 2>(P, st);
    }
    // launches that fit one 128-system CTA per SM are bound by the per-chunk hand-shakes of the serial chain:
    // 16-row chunks (see rc_api.cu: 1.18 vs 1.91 ms per 8760 steps)
    if ((P.N + 127) / 128 <= di.sm_count) return launch_fwd<M, 128, 16, 2, false, 2>(P, st);
    return launch_fwd<M, 256, 4, 4, false, 1>(P, st);  // wide rows: 5 KB+ per bulk copy (see rc_api.cu)
}

template <class M>
constexpr int dense_vjp_tn() {
    constexpr size_t lim = 227 * 1024;
    if (AdjCfg<M, 192, 8, 2, ADJ_VJP, false, false>::SMEM_BYTES <= lim) return 192;
    if (AdjCfg<M, 160, 8, 2, ADJ_VJP, false, false>::SMEM_BYTES <= lim) return 160;
    return 128;
}

template <int n, int m, int p>
static int lti_vjp_t(const AdjParams<DenseModel<n, m, p>> &P, bool shared_u, cudaStream_t st) {
    using M = DenseModel<n, m, p>;
    if (shared_u) return launch_adj<M, 128, 8, 2, ADJ_VJP, true, false, 1>(P, st);
    // Mid-size models (17..40 accumulators, e.g. (3,5,1)): 64-system CTAs, three per SM -- 1.25 vs 1.83 ms on
    // N=113664 x T=1024.  Larger ones ((4,3,4): 56) would spill at the 168 registers that leaves: 128-system CTAs.
    if constexpr (M::NG > 16 && M::NG <= 40) return launch_adj<M, 64, 8, 2, ADJ_VJP, false, false, 3>(P, st);
    // Large models ((4,3,4): 56 accumulators, ~250 registers) run one CTA per SM whatever the tile: the widest tile whose
    // ring + fp64 totals still fit in shared memory gives the most warps per SM (192 systems: 1.88 vs 2.61 ms on
    // N=113664 x T=1024).  Launches of at most one 128-system CTA per SM keep the narrow tile (more SMs busy).
    if constexpr (M::NG > 40 && dense_vjp_tn<M>() > 128) {
        DeviceInfo di;
        const int rc = device_info(&di);
        if (rc != DYNAX_OK) return rc;
        const char *e = getenv("DYNAX_LTI_VJP_CFG");  // 0 forces the 128-system tile
        if ((P.N + 127) / 128 > di.sm_count && !(e && atoi(e) == 0))
            return launch_adj<M, dense_vjp_tn<M>(), 8, 2, ADJ_VJP, false, false, 1>(P, st);
    }
    return launch_adj<M, 128, 8, 2, ADJ_VJP, false, false, 1>(P, st);
}

}  // namespace dynax

using namespace dynax;

extern "C" {

int dynax_lti_supported(int n, int m, int p) {
#define X(a, b, c) \
    if (n == a && m == b && p == c) return 1;
    DYNAX_LTI_DIMS(X)
#undef X
    return 0;
}

int dynax_lti_forward_f32(const float *A, const float *B, const float *C, const float *D, const float *x0,
                          const float *u, float *xsol, float *ysol, float *x_final, int64_t N, int64_t T, int n,
                          int m, int p, float dt, uint32_t flags, void *stream) {
    DYNAX_API_RANGE();
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (N < 0 || T < 0) return fail(DYNAX_ERR_INVALID_ARGUMENT, "negative N or T");
    if (N == 0) return DYNAX_OK;
    if (!A || !B || !C || !D || !x0 || (!u && T > 0))
        return fail(DYNAX_ERR_INVALID_ARGUMENT, "A,B,C,D,x0,u must be non-NULL");
    if (flags & ~(DYNAX_SHARED_THETA | DYNAX_SHARED_U | DYNAX_DISCRETE))
        return fail(DYNAX_ERR_INVALID_ARGUMENT, "unsupported flags 0x%x for lti_forward", flags);
    if (!dynax_lti_supported(n, m, p))
        return fail(DYNAX_ERR_UNSUPPORTED_DIMS, "no dense-LTI kernel instantiated for (n,m,p)=(%d,%d,%d)", n, m, p);
    int rc = require_sm100();
    if (rc != DYNAX_OK) return rc;
    if (T == 0) {
        if (x_final) DYNAX_CUDA_TRY(cudaMemcpyAsync(x_final, x0, sizeof(float) * n * N, cudaMemcpyDeviceToDevice, st));
        return DYNAX_OK;
    }
    const bool rows16 = (N % 4) == 0;
    const int bulk_in = rows16 && aligned16(u);
    const int bulk_out = rows16 && (!xsol || aligned16(xsol)) && (!ysol || aligned16(ysol));
#define X(a, b, c)                                                                     \
    if (n == a && m == b && p == c) {                                                  \
        FwdParams<DenseModel<a, b, c>> P;                                              \
        P.mp.A = A; P.mp.B = B; P.mp.C = C; P.mp.D = D;                                \
        P.mp.shared = (flags & DYNAX_SHARED_THETA) ? 1 : 0;                            \
        P.x0 = x0; P.u = u; P.xsol = xsol; P.ysol = ysol; P.x_final = x_final;         \
        P.N = N; P.T = T; P.dt = dt; P.discrete = (flags & DYNAX_DISCRETE) ? 1 : 0;    \
        P.bulk_in = bulk_in; P.bulk_out = bulk_out;                                    \
        return lti_forward_t<a, b, c>(P, (flags & DYNAX_SHARED_U) != 0, st);           \
    }
    DYNAX_LTI_DIMS(X)
#undef X
    return fail(DYNAX_ERR_UNSUPPORTED_DIMS, "unreachable");
}

int dynax_lti_vjp_f32(const float *A, const float *B, const float *C, const float *D, const float *x0,
                      const float *u, const float *g_xsol, const float *g_ysol, float *g_A, float *g_B, float *g_C,
                      float *g_D, float *g_x0, float *g_u, int64_t N, int64_t T, int n, int m, int p, float dt,
                      uint32_t flags, void *ws, size_t ws_bytes, void *stream) {
    DYNAX_API_RANGE();
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (N < 0 || T < 1) return fail(DYNAX_ERR_INVALID_ARGUMENT, "need N >= 0 and T >= 1");
    if (N == 0) return DYNAX_OK;
    if (!A || !B || !C || !D || !x0 || !u) return fail(DYNAX_ERR_INVALID_ARGUMENT, "A,B,C,D,x0,u must be non-NULL");
    if (flags & ~(DYNAX_SHARED_THETA | DYNAX_SHARED_U | DYNAX_DISCRETE))
        return fail(DYNAX_ERR_INVALID_ARGUMENT, "unsupported flags 0x%x for lti_vjp", flags);
    if (!g_xsol && !g_ysol) return fail(DYNAX_ERR_INVALID_ARGUMENT, "need g_xsol or g_ysol");
    if (!dynax_lti_supported(n, m, p))
        return fail(DYNAX_ERR_UNSUPPORTED_DIMS, "no dense-LTI kernel instantiated for (n,m,p)=(%d,%d,%d)", n, m, p);
    int rc = require_sm100();
    if (rc != DYNAX_OK) return rc;
    const size_t need = adj_workspace_bytes(N, T, n) + ((flags & DYNAX_SHARED_U) ? gu_part_bytes(N, T, m) : 0);
    if (!ws || ws_bytes < need)
        return fail(DYNAX_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", need, ws ? ws_bytes : (size_t)0);
    if (!aligned16(ws)) return fail(DYNAX_ERR_WORKSPACE, "workspace must be 16-byte aligned");
    const int bulk = (N % 4) == 0 && aligned16(u) && (!g_xsol || aligned16(g_xsol)) && (!g_ysol || aligned16(g_ysol));
#define X(a, b, c)                                                                               \
    if (n == a && m == b && p == c) {                                                            \
        AdjParams<DenseModel<a, b, c>> P;                                                        \
        memset(&P, 0, sizeof(P));                                                                \
        P.mp.A = A; P.mp.B = B; P.mp.C = C; P.mp.D = D;                                          \
        P.mp.shared = (flags & DYNAX_SHARED_THETA) ? 1 : 0;                                      \
        P.gp.gA = g_A; P.gp.gB = g_B; P.gp.gC = g_C; P.gp.gD = g_D;                              \
        P.x0 = x0; P.u = u; P.g_xsol = g_xsol; P.g_ysol = g_ysol; P.g_x0 = g_x0; P.g_u = g_u;    \
        if (g_u && (flags & DYNAX_SHARED_U))   /* g_u is [T,m] then: summed over systems in the kernel */ \
            P.gu_part = reinterpret_cast<float *>(reinterpret_cast<char *>(ws) + adj_workspace_bytes(N, T, a)); \
        P.shared_acc = reinterpret_cast<double *>(ws);                                           \
        P.ckpt = reinterpret_cast<float *>(reinterpret_cast<char *>(ws) + kAccBytes);            \
        P.N = N; P.T = T; P.dt = dt; P.discrete = (flags & DYNAX_DISCRETE) ? 1 : 0;              \
        P.shared_theta = (flags & DYNAX_SHARED_THETA) ? 1 : 0; P.bulk = bulk;                    \
        return lti_vjp_t<a, b, c>(P, (flags & DYNAX_SHARED_U) != 0, st);                         \
    }
    DYNAX_LTI_DIMS(X)
#undef X
    return fail(DYNAX_ERR_UNSUPPORTED_DIMS, "unreachable");
}

}  // extern "C"
