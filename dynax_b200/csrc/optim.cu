// Rows f-3 / f-4 of SURVEY.md section 8: keep the calibration loop and the input resampling on the device.
//
// dynax_lamb_update_f32: the parameter update of /root/reference/examples/3-inverse.py:117-122
//   (optax.lamb(1e-3) applied by train_state.apply_gradients, :113), for N candidates at once.  Every
//   Flax leaf there is a SCALAR, so LAMB's per-leaf trust ratio is per element:
//     m <- b1 m + (1-b1) g;  v <- b2 v + (1-b2) g^2;  u = (m/(1-b1^t)) / (sqrt(v/(1-b2^t)) + eps)
//     r = |theta|/|u|  (1 if either is 0);  theta <- theta - lr * r * u
//   optax is a third-party dependency absent from /root/reference and from this image (unpinned:
//   Dockerfile.jax:66-67); this restates its published lamb = scale_by_adam(eps=1e-6) ->
//   add_decayed_weights(0) -> scale_by_trust_ratio -> scale(-lr).  PARITY UNPINNED BY THE REFERENCE.
//
// dynax_resample_mean_f32: the input pipeline of examples/1-forward.py:26-27 and 3-inverse.py:34-35,
//   data.groupby(index // dt).mean(): mean of every `factor` consecutive rows (last group may be short).
#include "api_common.cuh"

namespace dynax {

__global__ void lamb_update_kernel(float *theta, const float *grad, float *m, float *v, int64_t n, float lr, float b1,
                                   float b2, float eps, float bc1, float bc2) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float g = grad[i];
    const float mi = b1 * m[i] + (1.f - b1) * g;
    const float vi = b2 * v[i] + (1.f - b2) * g * g;
    m[i] = mi;
    v[i] = vi;
    const float u = (mi / bc1) / (sqrtf(vi / bc2) + eps);
    const float th = theta[i];
    const float pn = fabsf(th), un = fabsf(u);
    const float trust = (pn == 0.f || un == 0.f) ? 1.f : pn / un;
    theta[i] = th - lr * trust * u;
}

// Same update with the epoch number kept ON THE DEVICE (incremented by the kernel), so that a captured
// CUDA graph of {loss+grad, update} can be replayed without any host-side parameter changing.
__global__ void lamb_update_dev_kernel(float *theta, const float *grad, float *m, float *v, int64_t n, int64_t *step,
                                       float lr, float b1, float b2, float eps) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t t = *step + 1;  // 1-based epoch of this update
    if (i < n) {
        const float bc1 = 1.f - powf(b1, (float)t), bc2 = 1.f - powf(b2, (float)t);
        const float g = grad[i];
        const float mi = b1 * m[i] + (1.f - b1) * g;
        const float vi = b2 * v[i] + (1.f - b2) * g * g;
        m[i] = mi;
        v[i] = vi;
        const float u = (mi / bc1) / (sqrtf(vi / bc2) + eps);
        const float th = theta[i];
        const float pn = fabsf(th), un = fabsf(u);
        const float trust = (pn == 0.f || un == 0.f) ? 1.f : pn / un;
        theta[i] = th - lr * trust * u;
    }
}

__global__ void bump_step_kernel(int64_t *step) { *step += 1; }

__global__ void resample_mean_kernel(const float *in, float *out, int64_t rows, int cols, int64_t factor,
                                     int64_t out_rows) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= out_rows * cols) return;
    const int64_t r = idx / cols;
    const int c = (int)(idx % cols);
    const int64_t lo = r * factor, hi = (lo + factor < rows) ? lo + factor : rows;
    double acc = 0.0;
    for (int64_t k = lo; k < hi; ++k) acc += (double)in[k * cols + c];
    out[idx] = (float)(acc / (double)(hi - lo));
}

}  // namespace dynax

using namespace dynax;

namespace dynax {
// Per-system non-finite guard: x + dt*rhs can blow up inside the parameter box of
// examples/3-inverse.py:66-85 (dt*|lambda| > 2), and the caller should learn it from a flag, not from an inf loss
// three calls later.  flags[i] = 1 if any of the `width` values of row i is NaN/Inf; *count += number of such rows.
__global__ void nonfinite_rows_kernel(const float *a, int64_t rows, int width, int32_t *flags, int *count) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int bad = 0;
    if (i < rows) {
        for (int j = 0; j < width; ++j) bad |= !isfinite(a[i * width + j]);
        if (flags) flags[i] = bad;
    }
    const unsigned m = __ballot_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0 && m && count) atomicAdd(count, __popc(m));
}
}  // namespace dynax

extern "C" {

int dynax_lamb_update_f32(float *theta, const float *grad, float *m, float *v, int64_t n, int64_t step, float lr,
                          float b1, float b2, float eps, void *stream) {
    DYNAX_API_RANGE();
    if (n < 0 || step < 1) return fail(DYNAX_ERR_INVALID_ARGUMENT, "need n >= 0 and step >= 1");
    if (n == 0) return DYNAX_OK;
    if (!theta || !grad || !m || !v) return fail(DYNAX_ERR_INVALID_ARGUMENT, "theta, grad, m, v must be non-NULL");
    int rc = require_sm100();
    if (rc != DYNAX_OK) return rc;
    const float bc1 = 1.f - powf(b1, (float)step), bc2 = 1.f - powf(b2, (float)step);
    const int tb = 256;
    lamb_update_kernel<<<(unsigned)((n + tb - 1) / tb), tb, 0, static_cast<cudaStream_t>(stream)>>>(
        theta, grad, m, v, n, lr, b1, b2, eps, bc1, bc2);
    DYNAX_CUDA_TRY(cudaGetLastError());
    return DYNAX_OK;
}

int dynax_lamb_update_dev_f32(float *theta, const float *grad, float *m, float *v, int64_t n, int64_t *step_counter,
                              float lr, float b1, float b2, float eps, void *stream) {
    DYNAX_API_RANGE();
    if (n < 0) return fail(DYNAX_ERR_INVALID_ARGUMENT, "need n >= 0");
    if (n == 0) return DYNAX_OK;
    if (!theta || !grad || !m || !v || !step_counter) return fail(DYNAX_ERR_INVALID_ARGUMENT, "NULL argument");
    int rc = require_sm100();
    if (rc != DYNAX_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int tb = 256;
    lamb_update_dev_kernel<<<(unsigned)((n + tb - 1) / tb), tb, 0, st>>>(theta, grad, m, v, n, step_counter, lr, b1, b2, eps);
    DYNAX_CUDA_TRY(cudaGetLastError());
    bump_step_kernel<<<1, 1, 0, st>>>(step_counter);
    DYNAX_CUDA_TRY(cudaGetLastError());
    return DYNAX_OK;
}

int dynax_resample_mean_f32(const float *in, float *out, int64_t rows, int cols, int64_t factor, void *stream) {
    DYNAX_API_RANGE();
    if (rows < 0 || cols < 1 || factor < 1) return fail(DYNAX_ERR_INVALID_ARGUMENT, "need rows >= 0, cols >= 1, factor >= 1");
    if (rows == 0) return DYNAX_OK;
    if (!in || !out) return fail(DYNAX_ERR_INVALID_ARGUMENT, "in and out must be non-NULL");
    int rc = require_sm100();
    if (rc != DYNAX_OK) return rc;
    const int64_t out_rows = (rows + factor - 1) / factor;
    const int tb = 256;
    resample_mean_kernel<<<(unsigned)((out_rows * cols + tb - 1) / tb), tb, 0, static_cast<cudaStream_t>(stream)>>>(
        in, out, rows, cols, factor, out_rows);
    DYNAX_CUDA_TRY(cudaGetLastError());
    return DYNAX_OK;
}

int dynax_nonfinite_rows_f32(const float *a, int64_t rows, int width, int32_t *flags, int *count, void *stream) {
    DYNAX_API_RANGE();
    if (rows < 0 || width < 1) return fail(DYNAX_ERR_INVALID_ARGUMENT, "need rows >= 0 and width >= 1");
    if (!a && rows > 0) return fail(DYNAX_ERR_INVALID_ARGUMENT, "a must be non-NULL");
    if (!flags && !count) return fail(DYNAX_ERR_INVALID_ARGUMENT, "need flags or count");
    int rc = require_sm100();
    if (rc != DYNAX_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (count) DYNAX_CUDA_TRY(cudaMemsetAsync(count, 0, sizeof(int), st));
    if (rows == 0) return DYNAX_OK;
    const int tb = 256;
    nonfinite_rows_kernel<<<(unsigned)((rows + tb - 1) / tb), tb, 0, st>>>(a, rows, width, flags, count);
    DYNAX_CUDA_TRY(cudaGetLastError());
    return DYNAX_OK;
}

}  // extern "C"
