// Host-buffer entry points: what a reference user calls with NumPy / JAX-CPU arrays.
//
// The system axis is cut into tiles; tile i+1's host->device copy (a strided 2-D copy out of the
// caller's [T,N,m] array) runs on a second stream while tile i's kernel and its device->host copy
// run on the first, so with pinned host memory the call is PCIe-bound, not latency-bound.
// Device scratch is cached across calls (grow-only) and released by dynax_host_release().
#include <mutex>
#include <string.h>

#include "api_common.cuh"

using namespace dynax;

namespace {

struct Slot {
    void *ptr = nullptr;
    size_t bytes = 0;
};

struct HostCtx {
    std::mutex mu;
    cudaStream_t streams[2] = {nullptr, nullptr};
    cudaEvent_t done[2] = {nullptr, nullptr};
    Slot slots[2][8];
    int device = -1;
};

HostCtx &ctx() {
    static HostCtx c;
    return c;
}

int ensure_ctx(HostCtx &c) {
    int dev = 0;
    DYNAX_CUDA_TRY(cudaGetDevice(&dev));
    if (c.device != dev) {
        for (int b = 0; b < 2; ++b) {
            if (c.streams[b]) cudaStreamDestroy(c.streams[b]);
            if (c.done[b]) cudaEventDestroy(c.done[b]);
            for (auto &s : c.slots[b]) {
                if (s.ptr) cudaFree(s.ptr);
                s = Slot();
            }
            DYNAX_CUDA_TRY(cudaStreamCreateWithFlags(&c.streams[b], cudaStreamNonBlocking));
            DYNAX_CUDA_TRY(cudaEventCreateWithFlags(&c.done[b], cudaEventDisableTiming));
        }
        c.device = dev;
    }
    return DYNAX_OK;
}

int reserve(Slot &s, size_t bytes, void **out) {
    if (bytes > s.bytes) {
        if (s.ptr) DYNAX_CUDA_TRY(cudaFree(s.ptr));
        s = Slot();
        DYNAX_CUDA_TRY(cudaMalloc(&s.ptr, bytes));
        s.bytes = bytes;
    }
    *out = s.ptr;
    return DYNAX_OK;
}

int64_t pick_tile(int64_t N, int64_t T, int64_t tile_systems, int bytes_per_system_step) {
    int64_t nt = tile_systems;
    if (nt <= 0) {
        // ~2 GB of streamed input per tile: the strided 2-D copies approach the contiguous PCIe rate only with
        // long rows (probe on one B200, 32768 systems x 8760: 4096-system tiles 51.0 GB/s, 8192: 53.2, 16384: 54.3,
        // one contiguous copy 55.5), and the kernels are ~100x faster than the link, so little overlap is lost
        const int64_t budget = 2048ll << 20;
        nt = budget / (T > 0 ? T * bytes_per_system_step : 1);
        if (nt < 4096) nt = 4096;
    }
    if (nt > N) nt = N;
    if (nt < N) nt = (nt + 127) / 128 * 128;  // keep interior tiles TMA-aligned
    if (nt > N) nt = N;
    return nt;
}

// strided [T, nt, w] block out of a [T, N, w] host array (or back)
int copy_tile(float *dst, size_t dpitch_f, const float *src, size_t spitch_f, int64_t nt, int w, int64_t T,
              cudaMemcpyKind kind, cudaStream_t st) {
    const size_t width = (size_t)nt * w;
    if (dpitch_f == width && spitch_f == width) {  // the tile is the whole system axis: one contiguous block, not T rows
        DYNAX_CUDA_TRY(cudaMemcpyAsync(dst, src, width * (size_t)T * sizeof(float), kind, st));
        return DYNAX_OK;
    }
    DYNAX_CUDA_TRY(cudaMemcpy2DAsync(dst, dpitch_f * sizeof(float), src, spitch_f * sizeof(float), width * sizeof(float),
                                     (size_t)T, kind, st));
    return DYNAX_OK;
}

#define TRY(expr)                       \
    do {                                \
        int rc_ = (expr);               \
        if (rc_ != DYNAX_OK) return rc_; \
    } while (0)

}  // namespace

extern "C" {

int dynax_host_release(void) {
    DYNAX_API_RANGE();
    HostCtx &c = ctx();
    std::lock_guard<std::mutex> lk(c.mu);
    for (int b = 0; b < 2; ++b) {
        for (auto &s : c.slots[b]) {
            if (s.ptr) cudaFree(s.ptr);
            s = Slot();
        }
    }
    return DYNAX_OK;
}

int dynax_rc_forward_host_f32(const float *theta, const float *x0, const float *u, float *xsol, float *ysol,
                              float *x_final, int64_t N, int64_t T, float dt, uint32_t flags, int64_t tile_systems) {
    DYNAX_API_RANGE();
    if (N < 0 || T < 0) return fail(DYNAX_ERR_INVALID_ARGUMENT, "negative N or T");
    if (N == 0) return DYNAX_OK;
    if (!theta || !x0 || (!u && T > 0)) return fail(DYNAX_ERR_INVALID_ARGUMENT, "theta, x0 and u must be non-NULL");
    TRY(require_sm100());
    HostCtx &c = ctx();
    std::lock_guard<std::mutex> lk(c.mu);
    TRY(ensure_ctx(c));
    const bool sth = flags & DYNAX_SHARED_THETA, su = flags & DYNAX_SHARED_U;
    const int64_t nt = pick_tile(N, T, tile_systems, 36);
    int b = 0;
    for (int64_t n0 = 0; n0 < N; n0 += nt, b ^= 1) {
        const int64_t cur = (N - n0) < nt ? (N - n0) : nt;
        cudaStream_t st = c.streams[b];
        // the previous use of this buffer set finished its D2H when its event fired
        DYNAX_CUDA_TRY(cudaEventSynchronize(c.done[b]));
        float *d_th, *d_x0, *d_u, *d_xs = nullptr, *d_ys = nullptr, *d_xf = nullptr;
        TRY(reserve(c.slots[b][0], sizeof(float) * 7 * (sth ? 1 : cur), (void **)&d_th));
        TRY(reserve(c.slots[b][1], sizeof(float) * 3 * cur, (void **)&d_x0));
        TRY(reserve(c.slots[b][2], sizeof(float) * 5 * (size_t)T * (su ? 1 : cur) + 16, (void **)&d_u));
        if (xsol) TRY(reserve(c.slots[b][3], sizeof(float) * 3 * (size_t)T * cur + 16, (void **)&d_xs));
        if (ysol) TRY(reserve(c.slots[b][4], sizeof(float) * (size_t)T * cur + 16, (void **)&d_ys));
        if (x_final) TRY(reserve(c.slots[b][5], sizeof(float) * 3 * cur, (void **)&d_xf));
        DYNAX_CUDA_TRY(cudaMemcpyAsync(d_th, theta + (sth ? 0 : n0 * 7), sizeof(float) * 7 * (sth ? 1 : cur),
                                       cudaMemcpyHostToDevice, st));
        DYNAX_CUDA_TRY(cudaMemcpyAsync(d_x0, x0 + n0 * 3, sizeof(float) * 3 * cur, cudaMemcpyHostToDevice, st));
        if (T > 0) {
            if (su) DYNAX_CUDA_TRY(cudaMemcpyAsync(d_u, u, sizeof(float) * 5 * T, cudaMemcpyHostToDevice, st));
            else TRY(copy_tile(d_u, (size_t)cur * 5, u + n0 * 5, (size_t)N * 5, cur, 5, T, cudaMemcpyHostToDevice, st));
        }
        // tiles that cannot fill the GPU with systems alone take the time-parallel kernels (break-evens of ops.py)
        if (!(flags & DYNAX_DISCRETE) && cur <= 8192 && T >= 256) {
            void *d_ws;
            const size_t wsb = dynax_rc_forward_chunked_workspace(cur, T, 0);
            TRY(reserve(c.slots[b][6], wsb, &d_ws));
            TRY(dynax_rc_forward_chunked_f32(d_th, d_x0, d_u, d_xs, d_ys, d_xf, cur, T, dt, flags, 0, d_ws, wsb, st));
        } else {
            TRY(dynax_rc_forward_f32(d_th, d_x0, d_u, d_xs, d_ys, d_xf, cur, T, dt, flags, st));
        }
        if (T > 0 && xsol)
            TRY(copy_tile(xsol + n0 * 3, (size_t)N * 3, d_xs, (size_t)cur * 3, cur, 3, T, cudaMemcpyDeviceToHost, st));
        if (T > 0 && ysol)
            TRY(copy_tile(ysol + n0, (size_t)N, d_ys, (size_t)cur, cur, 1, T, cudaMemcpyDeviceToHost, st));
        if (x_final)
            DYNAX_CUDA_TRY(cudaMemcpyAsync(x_final + n0 * 3, d_xf, sizeof(float) * 3 * cur, cudaMemcpyDeviceToHost, st));
        DYNAX_CUDA_TRY(cudaEventRecord(c.done[b], st));
    }
    DYNAX_CUDA_TRY(cudaStreamSynchronize(c.streams[0]));
    DYNAX_CUDA_TRY(cudaStreamSynchronize(c.streams[1]));
    return DYNAX_OK;
}

static int loss_grad_host_impl(const float *theta, const float *x0, const float *u, const float *ctrl, int ctrl_col,
                               const float *target, const float *lb, const float *ub, float *loss, float *g_theta,
                               int64_t N, int64_t T, float dt, uint32_t flags, int64_t tile_systems) {
    if (ctrl) flags |= DYNAX_SHARED_U;  // u is the shared [T,5] series, ctrl the per-system channel [T,N]
    if (N < 0 || T < 1) return fail(DYNAX_ERR_INVALID_ARGUMENT, "need N >= 0 and T >= 1");
    if (N == 0) return DYNAX_OK;
    if (!theta || !x0 || !u || !target || !loss || !g_theta)
        return fail(DYNAX_ERR_INVALID_ARGUMENT, "theta, x0, u, target, loss, g_theta must be non-NULL");
    if ((lb == nullptr) != (ub == nullptr)) return fail(DYNAX_ERR_INVALID_ARGUMENT, "lb and ub go together");
    TRY(require_sm100());
    HostCtx &c = ctx();
    std::lock_guard<std::mutex> lk(c.mu);
    TRY(ensure_ctx(c));
    const bool sth = flags & DYNAX_SHARED_THETA, su = flags & DYNAX_SHARED_U, stg = flags & DYNAX_SHARED_TARGET;
    const int64_t nt = pick_tile(N, T, tile_systems, ctrl ? 8 : 24);
    const int64_t ntiles = (N + nt - 1) / nt;
    // shared-theta partial sums per tile land here and are reduced on the host in fp64
    float(*part)[8] = sth ? new float[ntiles][8] : nullptr;
    int b = 0;
    int rc = DYNAX_OK;
    int64_t tile = 0;
    for (int64_t n0 = 0; n0 < N && rc == DYNAX_OK; n0 += nt, b ^= 1, ++tile) {
        const int64_t cur = (N - n0) < nt ? (N - n0) : nt;
        cudaStream_t st = c.streams[b];
        auto body = [&]() -> int {
            DYNAX_CUDA_TRY(cudaEventSynchronize(c.done[b]));
            float *d_th, *d_x0, *d_u, *d_tg, *d_out, *d_bounds = nullptr, *d_ctrl = nullptr;
            void *d_ws;
            // tiles that cannot fill the GPU with systems alone take the time-parallel kernels (break-evens of ops.py)
            const bool chunked = !ctrl && !(flags & DYNAX_DISCRETE) && cur <= 4096 && T >= 512;
            const size_t ws_bytes = chunked ? dynax_rc_loss_grad_chunked_workspace(cur, T, 0)
                                            : dynax_workspace_bytes(DYNAX_OP_RC_LOSS_GRAD, cur, T, 3, 5, 1, flags);
            TRY(reserve(c.slots[b][0], sizeof(float) * 7 * (sth ? 1 : cur), (void **)&d_th));
            TRY(reserve(c.slots[b][1], sizeof(float) * 3 * cur, (void **)&d_x0));
            TRY(reserve(c.slots[b][2], sizeof(float) * 5 * (size_t)T * (su ? 1 : cur) + 16, (void **)&d_u));
            TRY(reserve(c.slots[b][3], sizeof(float) * (size_t)T * (stg ? 1 : cur) + 16, (void **)&d_tg));
            TRY(reserve(c.slots[b][4], sizeof(float) * 8 * cur, (void **)&d_out));  // loss[cur] | grad[cur,7]
            TRY(reserve(c.slots[b][5], ws_bytes, &d_ws));
            if (lb && !sth) {
                TRY(reserve(c.slots[b][6], sizeof(float) * 16, (void **)&d_bounds));
                DYNAX_CUDA_TRY(cudaMemcpyAsync(d_bounds, lb, sizeof(float) * 7, cudaMemcpyHostToDevice, st));
                DYNAX_CUDA_TRY(cudaMemcpyAsync(d_bounds + 8, ub, sizeof(float) * 7, cudaMemcpyHostToDevice, st));
            }
            DYNAX_CUDA_TRY(cudaMemcpyAsync(d_th, theta + (sth ? 0 : n0 * 7), sizeof(float) * 7 * (sth ? 1 : cur),
                                           cudaMemcpyHostToDevice, st));
            DYNAX_CUDA_TRY(cudaMemcpyAsync(d_x0, x0 + n0 * 3, sizeof(float) * 3 * cur, cudaMemcpyHostToDevice, st));
            if (su) DYNAX_CUDA_TRY(cudaMemcpyAsync(d_u, u, sizeof(float) * 5 * T, cudaMemcpyHostToDevice, st));
            else TRY(copy_tile(d_u, (size_t)cur * 5, u + n0 * 5, (size_t)N * 5, cur, 5, T, cudaMemcpyHostToDevice, st));
            if (stg) DYNAX_CUDA_TRY(cudaMemcpyAsync(d_tg, target, sizeof(float) * T, cudaMemcpyHostToDevice, st));
            else TRY(copy_tile(d_tg, (size_t)cur, target + n0, (size_t)N, cur, 1, T, cudaMemcpyHostToDevice, st));
            float *d_loss = d_out, *d_g = d_out + cur;
            if (ctrl) {
                TRY(reserve(c.slots[b][7], sizeof(float) * (size_t)T * cur + 16, (void **)&d_ctrl));
                TRY(copy_tile(d_ctrl, (size_t)cur, ctrl + n0, (size_t)N, cur, 1, T, cudaMemcpyHostToDevice, st));
                TRY(dynax_rc_loss_grad_ctrl_f32(d_th, d_x0, d_u, d_ctrl, ctrl_col, d_tg, d_bounds,
                                                d_bounds ? d_bounds + 8 : nullptr, d_loss, d_g, nullptr, cur, T, dt,
                                                flags, d_ws, ws_bytes, st));
            } else if (chunked) {
                TRY(dynax_rc_loss_grad_chunked_f32(d_th, d_x0, d_u, d_tg, d_bounds, d_bounds ? d_bounds + 8 : nullptr, d_loss,
                                                   d_g, nullptr, cur, T, dt, flags, 0, d_ws, ws_bytes, st));
            } else {
                TRY(dynax_rc_loss_grad_f32(d_th, d_x0, d_u, d_tg, d_bounds, d_bounds ? d_bounds + 8 : nullptr, d_loss,
                                           d_g, nullptr, cur, T, dt, flags, d_ws, ws_bytes, st));
            }
            if (sth) {
                DYNAX_CUDA_TRY(cudaMemcpyAsync(&part[tile][0], d_loss, sizeof(float), cudaMemcpyDeviceToHost, st));
                DYNAX_CUDA_TRY(cudaMemcpyAsync(&part[tile][1], d_g, sizeof(float) * 7, cudaMemcpyDeviceToHost, st));
            } else {
                DYNAX_CUDA_TRY(cudaMemcpyAsync(loss + n0, d_loss, sizeof(float) * cur, cudaMemcpyDeviceToHost, st));
                DYNAX_CUDA_TRY(cudaMemcpyAsync(g_theta + n0 * 7, d_g, sizeof(float) * 7 * cur, cudaMemcpyDeviceToHost, st));
            }
            DYNAX_CUDA_TRY(cudaEventRecord(c.done[b], st));
            return DYNAX_OK;
        };
        rc = body();
    }
    cudaError_t e0 = cudaStreamSynchronize(c.streams[0]);
    cudaError_t e1 = cudaStreamSynchronize(c.streams[1]);
    if (rc == DYNAX_OK && (e0 != cudaSuccess || e1 != cudaSuccess))
        rc = fail(DYNAX_ERR_CUDA, "stream sync failed: %s", cudaGetErrorString(e0 != cudaSuccess ? e0 : e1));
    if (rc == DYNAX_OK && sth) {
        double L = 0.0, g[7] = {0, 0, 0, 0, 0, 0, 0};
        for (int64_t t = 0; t < ntiles; ++t) {
            L += part[t][0];
            for (int j = 0; j < 7; ++j) g[j] += part[t][1 + j];
        }
        if (lb)  // the box penalty counts once for the shared parameter set (3-inverse.py:103-107)
            for (int j = 0; j < 7; ++j) {
                const double th = theta[j], lo = lb[j], hi = ub[j];
                if (th > hi) { L += (th - hi) / hi; g[j] += 1.0 / hi; }
                if (th < lo) { L += (lo - th) / hi; g[j] -= 1.0 / hi; }
            }
        loss[0] = (float)L;
        for (int j = 0; j < 7; ++j) g_theta[j] = (float)g[j];
    }
    delete[] part;
    return rc;
}

// ---- resident data set: u and target stay in device memory across calls -------------------------------------
// examples/3-inverse.py:133-138 evaluates the SAME measured series 5000 times with changing parameters: uploading
// u[T,N,5] and target[T,N] once leaves 40 bytes per system and call (theta, x0) on the way in and 32 on the way out.
namespace {
struct RcDataset {
    float *d_u = nullptr, *d_tg = nullptr, *d_th = nullptr, *d_x0 = nullptr, *d_out = nullptr, *d_bounds = nullptr;
    void *d_ws = nullptr;
    size_t ws_bytes = 0;
    int64_t N = 0, T = 0;
    uint32_t flags = 0;
    int device = -1;
    cudaStream_t st = nullptr;
};
void dataset_free(RcDataset *d) {
    if (!d) return;
    cudaFree(d->d_u); cudaFree(d->d_tg); cudaFree(d->d_th); cudaFree(d->d_x0); cudaFree(d->d_out); cudaFree(d->d_bounds);
    cudaFree(d->d_ws);
    if (d->st) cudaStreamDestroy(d->st);
    delete d;
}
}  // namespace

int dynax_rc_dataset_create_f32(const float *u, const float *target, int64_t N, int64_t T, uint32_t flags, void **handle) {
    DYNAX_API_RANGE();
    if (!handle) return fail(DYNAX_ERR_INVALID_ARGUMENT, "handle must be non-NULL");
    *handle = nullptr;
    if (N < 1 || T < 1 || !u || !target) return fail(DYNAX_ERR_INVALID_ARGUMENT, "need N >= 1, T >= 1, u and target");
    if (flags & ~(DYNAX_SHARED_U | DYNAX_SHARED_TARGET)) return fail(DYNAX_ERR_INVALID_ARGUMENT, "flags: DYNAX_SHARED_U / DYNAX_SHARED_TARGET");
    TRY(require_sm100());
    RcDataset *d = new RcDataset;
    d->N = N; d->T = T; d->flags = flags;
    const size_t ub = sizeof(float) * 5 * (size_t)T * ((flags & DYNAX_SHARED_U) ? 1 : N);
    const size_t tb = sizeof(float) * (size_t)T * ((flags & DYNAX_SHARED_TARGET) ? 1 : N);
    auto body = [&]() -> int {
        DYNAX_CUDA_TRY(cudaGetDevice(&d->device));
        DYNAX_CUDA_TRY(cudaStreamCreateWithFlags(&d->st, cudaStreamNonBlocking));
        DYNAX_CUDA_TRY(cudaMalloc(&d->d_u, ub));
        DYNAX_CUDA_TRY(cudaMalloc(&d->d_tg, tb));
        DYNAX_CUDA_TRY(cudaMalloc(&d->d_th, sizeof(float) * 7 * N));
        DYNAX_CUDA_TRY(cudaMalloc(&d->d_x0, sizeof(float) * 3 * N));
        DYNAX_CUDA_TRY(cudaMalloc(&d->d_out, sizeof(float) * 8 * N));
        DYNAX_CUDA_TRY(cudaMalloc(&d->d_bounds, sizeof(float) * 16));
        // scratch for whichever algorithm a call may take (shared parameters: 1 KB; discrete stepping: checkpoints)
        d->ws_bytes = dynax_workspace_bytes(DYNAX_OP_RC_LOSS_GRAD, N, T, 3, 5, 1, DYNAX_DISCRETE);
        DYNAX_CUDA_TRY(cudaMalloc(&d->d_ws, d->ws_bytes));
        DYNAX_CUDA_TRY(cudaMemcpyAsync(d->d_u, u, ub, cudaMemcpyHostToDevice, d->st));
        DYNAX_CUDA_TRY(cudaMemcpyAsync(d->d_tg, target, tb, cudaMemcpyHostToDevice, d->st));
        DYNAX_CUDA_TRY(cudaStreamSynchronize(d->st));
        return DYNAX_OK;
    };
    const int rc = body();
    if (rc != DYNAX_OK) {
        dataset_free(d);
        return rc;
    }
    *handle = d;
    return DYNAX_OK;
}

int dynax_rc_dataset_loss_grad_f32(void *handle, const float *theta, const float *x0, const float *lb, const float *ub,
                                   float dt, uint32_t flags, float *loss, float *g_theta) {
    DYNAX_API_RANGE();
    RcDataset *d = static_cast<RcDataset *>(handle);
    if (!d || !theta || !x0 || !loss || !g_theta) return fail(DYNAX_ERR_INVALID_ARGUMENT, "handle, theta, x0, loss, g_theta must be non-NULL");
    if ((lb == nullptr) != (ub == nullptr)) return fail(DYNAX_ERR_INVALID_ARGUMENT, "lb and ub go together");
    if (flags & ~(DYNAX_SHARED_THETA | DYNAX_DISCRETE)) return fail(DYNAX_ERR_INVALID_ARGUMENT, "flags: DYNAX_SHARED_THETA / DYNAX_DISCRETE");
    int dev = -1;
    DYNAX_CUDA_TRY(cudaGetDevice(&dev));
    if (dev != d->device) return fail(DYNAX_ERR_INVALID_ARGUMENT, "data set lives on device %d, current device is %d", d->device, dev);
    const bool sth = flags & DYNAX_SHARED_THETA;
    const int64_t N = d->N, nout = sth ? 1 : N;
    cudaStream_t st = d->st;
    DYNAX_CUDA_TRY(cudaMemcpyAsync(d->d_th, theta, sizeof(float) * 7 * (sth ? 1 : N), cudaMemcpyHostToDevice, st));
    DYNAX_CUDA_TRY(cudaMemcpyAsync(d->d_x0, x0, sizeof(float) * 3 * N, cudaMemcpyHostToDevice, st));
    if (lb) {
        DYNAX_CUDA_TRY(cudaMemcpyAsync(d->d_bounds, lb, sizeof(float) * 7, cudaMemcpyHostToDevice, st));
        DYNAX_CUDA_TRY(cudaMemcpyAsync(d->d_bounds + 8, ub, sizeof(float) * 7, cudaMemcpyHostToDevice, st));
    }
    float *d_loss = d->d_out, *d_g = d->d_out + N;
    const uint32_t f = d->flags | flags;
    const bool chunked = !(f & DYNAX_DISCRETE) && N <= 4096 && d->T >= 512;   // break-even of ops.py
    if (chunked) {
        const size_t need = dynax_rc_loss_grad_chunked_workspace(N, d->T, 0);
        if (need > d->ws_bytes) {
            DYNAX_CUDA_TRY(cudaFree(d->d_ws));
            d->d_ws = nullptr; d->ws_bytes = 0;
            DYNAX_CUDA_TRY(cudaMalloc(&d->d_ws, need));
            d->ws_bytes = need;
        }
        TRY(dynax_rc_loss_grad_chunked_f32(d->d_th, d->d_x0, d->d_u, d->d_tg, lb ? d->d_bounds : nullptr, lb ? d->d_bounds + 8 : nullptr,
                                           d_loss, d_g, nullptr, N, d->T, dt, f, 0, d->d_ws, d->ws_bytes, st));
    } else {
        TRY(dynax_rc_loss_grad_f32(d->d_th, d->d_x0, d->d_u, d->d_tg, lb ? d->d_bounds : nullptr, lb ? d->d_bounds + 8 : nullptr, d_loss,
                                   d_g, nullptr, N, d->T, dt, f, d->d_ws, d->ws_bytes, st));
    }
    DYNAX_CUDA_TRY(cudaMemcpyAsync(loss, d_loss, sizeof(float) * nout, cudaMemcpyDeviceToHost, st));
    DYNAX_CUDA_TRY(cudaMemcpyAsync(g_theta, d_g, sizeof(float) * 7 * nout, cudaMemcpyDeviceToHost, st));
    DYNAX_CUDA_TRY(cudaStreamSynchronize(st));
    return DYNAX_OK;
}

int dynax_rc_dataset_destroy(void *handle) {
    DYNAX_API_RANGE();
    dataset_free(static_cast<RcDataset *>(handle));
    return DYNAX_OK;
}

// ---- page-locking a caller's long-lived host buffers (NumPy arrays are pageable: the driver then stages every copy
// through its own pinned buffer at a fraction of the link rate).  Thin wrappers so that a host program that never
// links the CUDA runtime can pin the arrays it passes to the *_host entry points again and again.
int dynax_host_register(const void *ptr, size_t bytes) {
    DYNAX_API_RANGE();
    if (!ptr || bytes == 0) return fail(DYNAX_ERR_INVALID_ARGUMENT, "ptr and bytes must be non-zero");
    TRY(require_sm100());
    DYNAX_CUDA_TRY(cudaHostRegister(const_cast<void *>(ptr), bytes, cudaHostRegisterDefault));
    return DYNAX_OK;
}

int dynax_host_unregister(const void *ptr) {
    DYNAX_API_RANGE();
    if (!ptr) return fail(DYNAX_ERR_INVALID_ARGUMENT, "ptr must be non-NULL");
    DYNAX_CUDA_TRY(cudaHostUnregister(const_cast<void *>(ptr)));
    return DYNAX_OK;
}

int dynax_rc_loss_grad_host_f32(const float *theta, const float *x0, const float *u, const float *target,
                                const float *lb, const float *ub, float *loss, float *g_theta, int64_t N,
                                int64_t T, float dt, uint32_t flags, int64_t tile_systems) {
    DYNAX_API_RANGE();
    return loss_grad_host_impl(theta, x0, u, nullptr, 0, target, lb, ub, loss, g_theta, N, T, dt, flags, tile_systems);
}

int dynax_rc_loss_grad_ctrl_host_f32(const float *theta, const float *x0, const float *u_shared, const float *ctrl,
                                     int ctrl_col, const float *target, const float *lb, const float *ub, float *loss,
                                     float *g_theta, int64_t N, int64_t T, float dt, uint32_t flags,
                                     int64_t tile_systems) {
    DYNAX_API_RANGE();
    if (!ctrl && N > 0) return fail(DYNAX_ERR_INVALID_ARGUMENT, "ctrl must be non-NULL");
    return loss_grad_host_impl(theta, x0, u_shared, ctrl, ctrl_col, target, lb, ub, loss, g_theta, N, T, dt, flags,
                               tile_systems);
}

}  // extern "C"
