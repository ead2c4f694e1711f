// Tabular agent / time-table interpolation (row f-1 of SURVEY.md section 8):
//   /root/reference/dynax/agents/tabular.py:33-44  ->  dynax/utils/interpolate.py:26-46 (constant), :48-68 (linear)
// Both modes first map the query time to a fractional ROW INDEX with jnp.interp(at, ts, arange(len(ts)))
// (piecewise-linear in ts, clamped at the ends) and then sample the table with
// jax.scipy.ndimage.map_coordinates: order=1 interpolates linearly between rows floor(idx) and
// floor(idx)+1, order=0 takes the NEAREST row, rounding halves away from zero
// (tests/test_agents.py:19: t=2.5 between rows 1 and 2 gives row 2).
#include "api_common.cuh"

namespace dynax {

__global__ void tabular_interp_kernel(const float *__restrict__ ts, const float *__restrict__ xs, int64_t n_rows,
                                      int n_cols, const float *__restrict__ at, int64_t n_at, int mode,
                                      float *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_at) return;
    const float t = at[i];
    // jnp.interp(t, ts, arange(n)): binary search for the bracketing knots, clamp outside
    float idx;
    if (!(t > ts[0])) {
        idx = 0.f;
    } else if (!(t < ts[n_rows - 1])) {
        idx = (float)(n_rows - 1);
    } else {
        int64_t lo = 0, hi = n_rows - 1;  // ts[lo] < t < ts[hi]... invariant ts[lo] <= t < ts[hi]
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (ts[mid] <= t) lo = mid; else hi = mid;
        }
        const float dt = ts[hi] - ts[lo];
        idx = (float)lo + (dt > 0.f ? (t - ts[lo]) / dt : 0.f);
    }
    if (mode == 1) {
        float fl = floorf(idx);
        int64_t r0 = (int64_t)fl;
        if (r0 > n_rows - 1) r0 = n_rows - 1;
        const int64_t r1 = r0 + 1 < n_rows ? r0 + 1 : n_rows - 1;
        const float w = idx - fl;
        for (int c = 0; c < n_cols; ++c) {
            const float a = xs[r0 * n_cols + c], b = xs[r1 * n_cols + c];
            out[i * n_cols + c] = a * (1.f - w) + b * w;   // map_coordinates order=1 weights
        }
    } else {
        int64_t r = (int64_t)roundf(idx);  // half away from zero
        if (r > n_rows - 1) r = n_rows - 1;
        if (r < 0) r = 0;
        for (int c = 0; c < n_cols; ++c) out[i * n_cols + c] = xs[r * n_cols + c];
    }
}

}  // namespace dynax

using namespace dynax;

extern "C" int dynax_tabular_interp_f32(const float *ts, const float *xs, int64_t n_rows, int n_cols, const float *at,
                                        int64_t n_at, int mode, float *out, void *stream) {
    DYNAX_API_RANGE();
    if (n_at < 0 || n_rows < 1 || n_cols < 1) return fail(DYNAX_ERR_INVALID_ARGUMENT, "need n_rows >= 1, n_cols >= 1, n_at >= 0");
    if (mode != 0 && mode != 1) return fail(DYNAX_ERR_INVALID_ARGUMENT, "mode must be 0 (constant/nearest) or 1 (linear)");
    if (n_at == 0) return DYNAX_OK;
    if (!ts || !xs || !at || !out) return fail(DYNAX_ERR_INVALID_ARGUMENT, "ts, xs, at, out must be non-NULL");
    int rc = require_sm100();
    if (rc != DYNAX_OK) return rc;
    const int tb = 256;
    tabular_interp_kernel<<<(unsigned)((n_at + tb - 1) / tb), tb, 0, static_cast<cudaStream_t>(stream)>>>(
        ts, xs, n_rows, n_cols, at, n_at, mode, out);
    DYNAX_CUDA_TRY(cudaGetLastError());
    return DYNAX_OK;
}
