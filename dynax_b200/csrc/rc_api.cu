// C-ABI entry points for the 4R3C RC model (device-pointer flavour) + library-wide helpers.
// See include/dynax_b200.h for the contract and the reference code each entry replaces.
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <string>
#include <vector>

#include "launch.cuh"

namespace dynax {

char *last_error_buf() {
    static thread_local char buf[kErrBufLen] = {0};
    return buf;
}

// ---- launch registry (api_common.cuh) ----
namespace {
struct Registry {
    std::mutex mu;
    std::vector<std::string> tags;
    std::vector<long long> launches;
};
Registry &registry() {
    static Registry *r = new Registry;  // never destroyed: registrations run during static initialisation
    return *r;
}
thread_local int last_launch_id = -1;
}  // namespace

int registry_add(const char *tag) {
    Registry &r = registry();
    std::lock_guard<std::mutex> lk(r.mu);
    r.tags.emplace_back(tag);
    r.launches.push_back(0);
    return (int)r.tags.size() - 1;
}

void registry_note(int id) {
    Registry &r = registry();
    std::lock_guard<std::mutex> lk(r.mu);
    r.launches[id] += 1;
    last_launch_id = id;
}

int require_sm100() {
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess)
        return fail(DYNAX_ERR_NO_DEVICE, "no CUDA device: %s (dynax_b200 has no CPU fallback)", cudaGetErrorString(e));
    int major = 0;
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess) return fail(DYNAX_ERR_NO_DEVICE, "cudaDeviceGetAttribute: %s", cudaGetErrorString(e));
    if (major != 10)
        return fail(DYNAX_ERR_NO_DEVICE, "device %d is compute capability %d.x; this library is sm_100a only", dev, major);
    return DYNAX_OK;
}

int device_info(DeviceInfo *out) {
    int dev = 0;
    DYNAX_CUDA_TRY(cudaGetDevice(&dev));
    DYNAX_CUDA_TRY(cudaDeviceGetAttribute(&out->sm_count, cudaDevAttrMultiProcessorCount, dev));
    DYNAX_CUDA_TRY(cudaDeviceGetAttribute(&out->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    return DYNAX_OK;
}

static int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return v ? atoi(v) : dflt;
}

// The fused loss+gradient runs the forward-mode (tangent) kernel unless the caller steps a discrete model or the
// environment holds DYNAX_GRAD_ALGO=0 (then, and for general cotangents: the checkpointed adjoint).
static bool rc_uses_tangent(uint32_t flags) { return env_int("DYNAX_GRAD_ALGO", 1) == 1 && !(flags & DYNAX_DISCRETE); }

static size_t rc_workspace_bytes(int op, int64_t N, int64_t T, uint32_t flags) {
    if (op == DYNAX_OP_RC_LOSS_GRAD && rc_uses_tangent(flags)) return (flags & DYNAX_SHARED_THETA) ? kAccBytes : 0;
    // general cotangents with a shared input series: + the per-CTA partial sums of g_u over systems
    return adj_workspace_bytes(N, T, 3) + ((op == DYNAX_OP_RC_VJP && (flags & DYNAX_SHARED_U)) ? gu_part_bytes(N, T, 5) : 0);
}

// ------------------------------------------------------------------------------------------
static int rc_forward(const float *theta, const float *x0, const float *u, float *xsol, float *ysol,
                      float *x_final, int64_t N, int64_t T, float dt, uint32_t flags, cudaStream_t st,
                      const float *ctrl = nullptr, int ctrl_col = 0) {
    if (N < 0 || T < 0) return fail(DYNAX_ERR_INVALID_ARGUMENT, "negative N or T");
    if (N == 0) return DYNAX_OK;
    if (!theta || !x0 || (!u && T > 0)) return fail(DYNAX_ERR_INVALID_ARGUMENT, "theta, x0 and u must be non-NULL");
    if (flags & ~(DYNAX_SHARED_THETA | DYNAX_SHARED_U | DYNAX_DISCRETE))
        return fail(DYNAX_ERR_INVALID_ARGUMENT, "unsupported flags 0x%x for rc_forward", flags);
    int rc = require_sm100();
    if (rc != DYNAX_OK) return rc;
    if (T == 0) {  // empty scan: nothing to emit, final state is the initial state
        if (x_final) DYNAX_CUDA_TRY(cudaMemcpyAsync(x_final, x0, sizeof(float) * 3 * N, cudaMemcpyDeviceToDevice, st));
        return DYNAX_OK;
    }
    FwdParams<RCModel> P;
    P.mp.theta = theta;
    P.mp.stride = (flags & DYNAX_SHARED_THETA) ? 0 : 7;
    P.x0 = x0; P.u = u; P.xsol = xsol; P.ysol = ysol; P.x_final = x_final;
    P.N = N; P.T = T; P.dt = dt;
    P.discrete = (flags & DYNAX_DISCRETE) ? 1 : 0;
    const bool rows16 = (N % 4) == 0;  // every [*,N,w] row starts and ends on a 16-byte boundary
    P.bulk_in = rows16 && aligned16(u) && (!ctrl || aligned16(ctrl));
    P.bulk_out = rows16 && (!xsol || aligned16(xsol)) && (!ysol || aligned16(ysol));
    if (env_int("DYNAX_NO_BULK", 0)) P.bulk_in = P.bulk_out = 0;
    if (env_int("DYNAX_FWD_DIRECT", 0)) P.bulk_out = 0;  // tuning knob: per-thread stores even where TMA could move the rows
    if (ctrl) {
        if (!(flags & DYNAX_SHARED_U)) return fail(DYNAX_ERR_INVALID_ARGUMENT, "a control channel needs a shared input series");
        if (ctrl_col < 0 || ctrl_col >= 5) return fail(DYNAX_ERR_INVALID_ARGUMENT, "ctrl_col must be in [0,5)");
        P.ctrl = ctrl;
        P.ctrl_col = ctrl_col;
    }
    if (flags & DYNAX_SHARED_U) {
        // A shared input series leaves almost nothing to stream in, so the short chunks that suit many CTAs per SM
        // (4 rows: the out-tiles stay small) make a launch of few CTAs wait on its own hand-shakes: up to two
        // 128-system CTAs per SM take 16-row chunks and a 4-stage ring (N = 128 .. 18 944, T = 8760, x+y: 1.70-1.72 ->
        // 0.97-1.07 ms; with a control channel 2.03-2.06 -> 1.17-1.26 ms; N = 65 536: 2.46 -> 2.23 ms but 1.47 -> 1.97 ms
        // without outputs, so the short chunks stay above the threshold).  DYNAX_FWD_CFG=0 forces the short chunks.
        DeviceInfo di;
        rc = device_info(&di);
        if (rc != DYNAX_OK) return rc;
        if (N <= (int64_t)di.sm_count * 256 && env_int("DYNAX_FWD_CFG", -1) != 0) return launch_fwd<RCModel, 128, 16, 4, true, 2>(P, st);
        return launch_fwd<RCModel, 128, 4, 4, true, 4>(P, st);
    }
    // Tile shapes measured on B200 (tools/tune.py, N=132608, T=8760): wide rows win because each
    // cp.async.bulk then moves 5 KB: 256-system CTAs 6.45 TB/s (x+y), 128-system 5.43, 64-system 5.31.
    // Launches that leave at most one 256-system CTA per SM are bound by the per-chunk hand-shakes of the serial
    // chain, not by HBM: longer chunks amortise them -- 16 rows while one 128-system CTA per SM suffices (N=16384:
    // 1.18 ms; 8 rows 1.37, 4 rows 1.91), 8 rows up to one 256-system CTA per SM (N=32768: 1.65 vs 1.91 ms).
    int cfg = env_int("DYNAX_FWD_CFG", -1);
    if (cfg < 0) {
        DeviceInfo di;
        rc = device_info(&di);
        if (rc != DYNAX_OK) return rc;
        // (one 128-system CTA per SM on more than half of the SMs: a 4-stage ring, N = 16 384: 0.98 -> 0.92 ms, 18 944: 1.05 ->
        // 0.96 ms; below that the ring depth makes no difference)
        cfg = ((N + 127) / 128 <= di.sm_count) ? (N > (int64_t)di.sm_count * 64 ? 5 : 4) : ((N + 255) / 256 <= di.sm_count) ? 3 : 0;
    }
    // rows TMA cannot move (N % 4 != 0) on launches of more than one wave: the shifted-row instantiation (16-byte cp.async)
    if (cfg == 0 && !P.bulk_in && !env_int("DYNAX_NO_SHIFT", 0)) return launch_fwd<RCModel, 256, 4, 4, false, 1, true>(P, st);
    switch (cfg) {
        default:
        case 0: return launch_fwd<RCModel, 256, 4, 4, false, 1>(P, st);
        case 3: return launch_fwd<RCModel, 128, 8, 2, false, 3>(P, st);
        case 4: return launch_fwd<RCModel, 128, 16, 2, false, 2>(P, st);
        case 5: return launch_fwd<RCModel, 128, 16, 4, false, 1>(P, st);
    }
}

template <int MODE>
static int rc_adjoint(const float *theta, const float *x0, const float *u, const float *target, const float *lb,
                      const float *ub, const float *g_xsol, const float *g_ysol, float *loss, float *g_theta,
                      float *g_x0, float *g_u, int64_t N, int64_t T, float dt, uint32_t flags, void *ws,
                      size_t ws_bytes, cudaStream_t st, const float *ctrl = nullptr, int ctrl_col = 0) {
    if (N < 0 || T < 1) return fail(DYNAX_ERR_INVALID_ARGUMENT, "need N >= 0 and T >= 1");
    if (N == 0) return DYNAX_OK;
    if (!theta || !x0 || !u) return fail(DYNAX_ERR_INVALID_ARGUMENT, "theta, x0 and u must be non-NULL");
    const uint32_t allowed = DYNAX_SHARED_THETA | DYNAX_SHARED_U | DYNAX_DISCRETE | (MODE == ADJ_MSE ? DYNAX_SHARED_TARGET : 0u);
    if (flags & ~allowed) return fail(DYNAX_ERR_INVALID_ARGUMENT, "unsupported flags 0x%x", flags);
    if (MODE == ADJ_MSE && (!target || !loss)) return fail(DYNAX_ERR_INVALID_ARGUMENT, "target and loss must be non-NULL");
    if (MODE == ADJ_MSE && ((lb == nullptr) != (ub == nullptr)))
        return fail(DYNAX_ERR_INVALID_ARGUMENT, "lb and ub must both be given or both be NULL");
    if (MODE == ADJ_VJP && !g_xsol && !g_ysol) return fail(DYNAX_ERR_INVALID_ARGUMENT, "need g_xsol or g_ysol");
    int rc = require_sm100();
    if (rc != DYNAX_OK) return rc;
    // the forward-mode kernel needs no scratch beyond the 1 KB of fp64 sums of a shared parameter vector; the
    // checkpointed adjoint needs its checkpoints (rc_workspace_bytes is what dynax_workspace_bytes reports)
    const size_t need = rc_workspace_bytes(MODE == ADJ_MSE ? DYNAX_OP_RC_LOSS_GRAD : DYNAX_OP_RC_VJP, N, T, flags);
    if (need > 0 && (!ws || ws_bytes < need))
        return fail(DYNAX_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", need, ws ? ws_bytes : (size_t)0);
    if (need > 0 && !aligned16(ws)) return fail(DYNAX_ERR_WORKSPACE, "workspace must be 16-byte aligned");

    AdjParams<RCModel> P;
    memset(&P, 0, sizeof(P));
    P.mp.theta = theta;
    P.mp.stride = (flags & DYNAX_SHARED_THETA) ? 0 : 7;
    P.gp.g_theta = g_theta;
    P.x0 = x0; P.u = u; P.target = target; P.lb = lb; P.ub = ub; P.loss = loss;
    P.g_xsol = g_xsol; P.g_ysol = g_ysol; P.g_x0 = g_x0; P.g_u = g_u;
    if (MODE == ADJ_VJP && g_u && (flags & DYNAX_SHARED_U))   // g_u is [T,5] then: summed over systems inside the kernel
        P.gu_part = reinterpret_cast<float *>(reinterpret_cast<char *>(ws) + adj_workspace_bytes(N, T, 3));
    P.shared_acc = reinterpret_cast<double *>(ws);
    P.ckpt = ws ? reinterpret_cast<float *>(reinterpret_cast<char *>(ws) + kAccBytes) : nullptr;
    P.N = N; P.T = T; P.dt = dt;
    P.discrete = (flags & DYNAX_DISCRETE) ? 1 : 0;
    P.shared_theta = (flags & DYNAX_SHARED_THETA) ? 1 : 0;
    P.bulk = (N % 4) == 0 && aligned16(u) && (!target || aligned16(target)) && (!g_xsol || aligned16(g_xsol)) &&
             (!g_ysol || aligned16(g_ysol)) && (!ctrl || aligned16(ctrl));
    if (ctrl) {
        if (!(flags & DYNAX_SHARED_U)) return fail(DYNAX_ERR_INVALID_ARGUMENT, "a control channel needs a shared input series");
        if (ctrl_col < 0 || ctrl_col >= 5) return fail(DYNAX_ERR_INVALID_ARGUMENT, "ctrl_col must be in [0,5)");
        P.ctrl = ctrl;
        P.ctrl_col = ctrl_col;
    }
    if (env_int("DYNAX_NO_BULK", 0)) P.bulk = 0;
    const bool su = flags & DYNAX_SHARED_U, stg = flags & DYNAX_SHARED_TARGET;
    // Forward-mode (tangent) single sweep for the fused 4R3C loss+gradient with per-system theta (rc_fwdsens.cuh):
    // one pass over u and target (24 B per system-step; nothing at all when both are shared series), measured
    // 2.2-2.3e11 system-steps/s against 1.3-1.4e11 for the checkpointed adjoint below on the same box.
    // DYNAX_GRAD_ALGO=0 forces the adjoint; the adjoint also serves every case the tangent kernel does not cover
    // (discrete stepping, general cotangents).  d loss / d x0 rides along as three more sensitivity columns.
    if (MODE == ADJ_MSE && rc_uses_tangent(flags)) {
        FwdSensParams F;
        memset(&F, 0, sizeof(F));
        F.mp = P.mp; F.x0 = x0; F.u = u; F.target = target; F.lb = lb; F.ub = ub; F.loss = loss; F.g_theta = g_theta;
        F.ctrl = ctrl; F.ctrl_col = ctrl_col; F.g_x0 = g_x0;
        F.shared_acc = P.shared_theta ? P.shared_acc : nullptr;
        F.N = N; F.T = T; F.dt = dt;
        F.bulk = (N % 4) == 0 && (su || aligned16(u)) && (stg || aligned16(target)) && (!ctrl || aligned16(ctrl)) &&
                 !env_int("DYNAX_NO_BULK", 0);
        // Tile shape.  Every CTA runs the whole horizon, so a launch costs (waves x systems per CTA):
        //   * launches that fit one 64-system CTA per SM (N <= 9472; BASELINE.json config 3 sharded over 8 GPUs is 8192
        //     systems per rank) give every system 2 threads, each owning a subset of the sensitivity columns (SPLIT,
        //     rc_fwdsens.cuh), with 64-row chunks rolled 8 steps at a time: such a launch runs ONE warp per scheduler,
        //     so the serial chain of 8760 steps and the per-chunk hand-shakes bound it (N = 8192, T = 8760, same box:
        //     round-1 kernel 0.87 ms; + tensor maps 0.73; + 2 threads per system 0.57; + 64-row chunks 0.50 ms).
        //     Measured and NOT kept (profiles/r02_small_launch_sweep.txt): 4 threads per system (x is stepped 4 times
        //     per system: 0.61-0.70 ms), row prefetch into registers (0.63-0.70), unpredicated chunk bodies (0.66),
        //     32-system CTAs (0.81);
        //   * up to sm_count x 128 systems with a control channel or d loss / d x0: 128-system CTAs, 16-row chunks;
        //   * 224-system CTAs when they save a wave (N = 65536: 293 CTAs fill the 296 slots once, 256 CTAs of 256
        //     leave 40 SMs half empty);   * else 256-system CTAs, 2 per SM (5 KB per TMA row copy).
        // DYNAX_FS_CFG: 2 = always 128-wide, 3 = always 256-wide; DYNAX_FS_SPLIT: 1 or 2 forces the threads per system.
        DeviceInfo di;
        rc = device_info(&di);
        if (rc != DYNAX_OK) return rc;
        const int64_t slots = 2 * (int64_t)di.sm_count;
        auto cost = [&](int64_t tn) { return ((N + tn - 1) / tn + slots - 1) / slots * tn; };
        const int cfg = env_int("DYNAX_FS_CFG", 0);
        int split = env_int("DYNAX_FS_SPLIT", 0);
        if (ctrl || g_x0) split = 1;
        else if (split != 1 && split != 2) split = (cfg != 0) ? 1 : (N <= (int64_t)di.sm_count * 64) ? 2 : 1;
        int tn = 256;
        if (cfg == 2 || (cfg == 0 && (N + 127) / 128 <= di.sm_count)) tn = 128;
        else if (cfg == 0 && cost(224) < cost(256)) tn = 224;
        // (with g_x0: 11 fp64 totals per system in shared memory -> 128-system CTAs x3 or 224-system CTAs x2 per SM)
#define DYNAX_FS_LAUNCH(SU, STG, CTRL)                                                                \
    (g_x0 ? (tn == 128 ? launch_fwdsens<128, 16, 2, 2, 1, SU, STG, CTRL, true>(F, st)                   \
                       : launch_fwdsens<224, 8, 2, 2, 1, SU, STG, CTRL, true>(F, st))                   \
          : tn == 128 ? launch_fwdsens<128, 16, 2, 2, 1, SU, STG, CTRL>(F, st)                          \
                      : tn == 224 ? launch_fwdsens<224, 8, 2, 2, 1, SU, STG, CTRL>(F, st)               \
                                  : launch_fwdsens<256, 8, 2, 2, 1, SU, STG, CTRL>(F, st))
#define DYNAX_FS_LAUNCH_SPLIT(SU, STG)                                                                \
    (split == 2 ? launch_fwdsens<64, 64, 2, 1, 2, SU, STG, false, false, false, 8>(F, st)       \
                : DYNAX_FS_LAUNCH(SU, STG, false))
        // rows TMA cannot move (N % 4 != 0), per-system u and target, more than one 128-system CTA per SM: the shifted-row
        // instantiation (one bulk copy per row; N = 151 553, T = 2000: 3.05 ms with 4-byte copies by the producer lanes,
        // 2.04 ms with every thread fetching its own rows, 1.60 ms shifted; (224,4,4) 2.08, (256,8,2) 1.66, (128,8,4) 2.40,
        // (128,16,2) 1.87 ms; profiles/r02_unaligned_ab.txt)
        if (!F.bulk && !su && !stg && !ctrl && !g_x0 && split == 1 && tn != 128 && !env_int("DYNAX_NO_SHIFT", 0)) {
            return launch_fwdsens<224, 8, 2, 2, 1, false, false, false, false, false, 8, true>(F, st);
        }
        if (ctrl && stg) return DYNAX_FS_LAUNCH(true, true, true);
        if (ctrl) return DYNAX_FS_LAUNCH(true, false, true);
        if (su && stg) return DYNAX_FS_LAUNCH_SPLIT(true, true);
        if (su) return DYNAX_FS_LAUNCH_SPLIT(true, false);
        if (stg) return DYNAX_FS_LAUNCH_SPLIT(false, true);
        return DYNAX_FS_LAUNCH_SPLIT(false, false);
#undef DYNAX_FS_LAUNCH
#undef DYNAX_FS_LAUNCH_SPLIT
    }
    if constexpr (MODE == ADJ_VJP) {
        if (su) return launch_adj<RCModel, 128, 8, 2, ADJ_VJP, true, false, 2>(P, st);
        // launches that fit one CTA per SM are bound by the per-segment hand-shakes of the serial chain: 16-step segments
        if (env_int("DYNAX_VJP_CFG", 0) == 0) {
            DeviceInfo di;
            rc = device_info(&di);
            if (rc != DYNAX_OK) return rc;
            if ((N + 127) / 128 <= di.sm_count) return launch_adj<RCModel, 128, 16, 2, ADJ_VJP, false, false, 1>(P, st);
        }
        // full cotangents in and g_u out (79 B per system-step): 12-step segments, 6.24 vs 6.00 TB/s; with g_ysol only
        // (47 B) 8-step segments are faster (5.66 vs 4.99 TB/s)
        // rows TMA cannot move (N % 4 != 0), more than one CTA per SM: the shifted-row instantiation (16-byte cp.async)
        if (!P.bulk && (N + 127) / 128 > 148 && !env_int("DYNAX_NO_SHIFT", 0))
            return launch_adj<RCModel, 128, 8, 2, ADJ_VJP, false, false, 2, true>(P, st);
        if (g_xsol && g_u) return launch_adj<RCModel, 128, 12, 2, ADJ_VJP, false, false, 2>(P, st);
        return launch_adj<RCModel, 128, 8, 2, ADJ_VJP, false, false, 2>(P, st);
    } else {
        // shared-u stages are small (<= 17 KB): a 4-deep ring keeps the producer ahead of the consumers
        if (su && stg) return launch_adj<RCModel, 128, 16, 4, ADJ_MSE, true, true, 2>(P, st);
        if (su) return launch_adj<RCModel, 128, 16, 4, ADJ_MSE, true, false, 2>(P, st);
        if (stg) return launch_adj<RCModel, 128, 12, 2, ADJ_MSE, false, true, 3>(P, st);
        // Tile shapes measured on B200 (bench.py, N=2^20, T=8760, same box): KC=12 keeps 3 CTAs per SM
        // (74 KB smem, 128 regs) and is ~4 % faster than KC=16 with 2 CTAs per SM when the grid fills whole
        // waves -- but every CTA runs for the whole horizon, so a launch costs ceil(CTAs/slots) full waves.
        // Pick the shape whose wave quantisation wastes less (N=65536: 512 CTAs = 1.15 waves of 444 slots
        // but 1.73 waves of 296; measured 5.6 ms vs 4.7 ms).  DYNAX_ADJ_CFG >= 0 forces a shape.
        int cfg = env_int("DYNAX_ADJ_CFG", -1);
        if (cfg < 0) {
            DeviceInfo di;
            rc = device_info(&di);
            if (rc != DYNAX_OK) return rc;
            const int64_t ctas = (N + 127) / 128, s3 = 3ll * di.sm_count, s2 = 2ll * di.sm_count;
            const double cost3 = (double)((ctas + s3 - 1) / s3 * s3) / 1.04, cost2 = (double)((ctas + s2 - 1) / s2 * s2);
            cfg = (ctas <= s3 || cost3 <= cost2) ? 0 : 1;
        }
        // rows TMA cannot move (N % 4 != 0) on launches of more than one CTA per SM: the shifted-row instantiation
        if (!P.bulk && (N + 127) / 128 > 148 && !env_int("DYNAX_NO_SHIFT", 0))
            return launch_adj<RCModel, 128, 12, 2, ADJ_MSE, false, false, 3, true>(P, st);
        switch (cfg) {
            default:
            case 0: return launch_adj<RCModel, 128, 12, 2, ADJ_MSE, false, false, 3>(P, st);
            case 1: return launch_adj<RCModel, 128, 16, 2, ADJ_MSE, false, false, 2>(P, st);
        }
    }
}

}  // namespace dynax

using namespace dynax;

extern "C" {

int dynax_version(void) { return DYNAX_ABI_VERSION; }

const char *dynax_last_error(void) { return last_error_buf(); }

int dynax_device_check(void) { return require_sm100(); }

int dynax_debug_kernel_count(void) {
    std::lock_guard<std::mutex> lk(registry().mu);
    return (int)registry().tags.size();
}

const char *dynax_debug_kernel_tag(int i) {
    std::lock_guard<std::mutex> lk(registry().mu);
    return (i >= 0 && i < (int)registry().tags.size()) ? registry().tags[i].c_str() : "";
}

long long dynax_debug_kernel_launches(int i) {
    std::lock_guard<std::mutex> lk(registry().mu);
    return (i >= 0 && i < (int)registry().launches.size()) ? registry().launches[i] : -1;
}

const char *dynax_debug_last_kernel(void) { return dynax_debug_kernel_tag(last_launch_id); }

size_t dynax_workspace_bytes(int op, int64_t N, int64_t T, int n, int m, int p, uint32_t flags) {
    (void)p;
    switch (op) {
        case DYNAX_OP_RC_LOSS_GRAD:
        case DYNAX_OP_RC_VJP: return rc_workspace_bytes(op, N, T, flags);
        case DYNAX_OP_LTI_VJP: return adj_workspace_bytes(N, T, n) + ((flags & DYNAX_SHARED_U) ? gu_part_bytes(N, T, m) : 0);
        default: return 0;
    }
}

int dynax_rc_forward_f32(const float *theta, const float *x0, const float *u, float *xsol, float *ysol,
                         float *x_final, int64_t N, int64_t T, float dt, uint32_t flags, void *stream) {
    DYNAX_API_RANGE();
    return rc_forward(theta, x0, u, xsol, ysol, x_final, N, T, dt, flags, static_cast<cudaStream_t>(stream));
}

int dynax_rc_loss_grad_f32(const float *theta, const float *x0, const float *u, const float *target,
                           const float *lb, const float *ub, float *loss, float *g_theta, float *g_x0, int64_t N,
                           int64_t T, float dt, uint32_t flags, void *ws, size_t ws_bytes, void *stream) {
    DYNAX_API_RANGE();
    return rc_adjoint<ADJ_MSE>(theta, x0, u, target, lb, ub, nullptr, nullptr, loss, g_theta, g_x0, nullptr, N, T, dt,
                               flags, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}

int dynax_rc_forward_ctrl_f32(const float *theta, const float *x0, const float *u_shared, const float *ctrl,
                              int ctrl_col, float *xsol, float *ysol, float *x_final, int64_t N, int64_t T, float dt,
                              uint32_t flags, void *stream) {
    DYNAX_API_RANGE();
    if (!ctrl && N > 0 && T > 0) return fail(DYNAX_ERR_INVALID_ARGUMENT, "ctrl must be non-NULL");
    return rc_forward(theta, x0, u_shared, xsol, ysol, x_final, N, T, dt, flags | DYNAX_SHARED_U,
                      static_cast<cudaStream_t>(stream), ctrl, ctrl_col);
}

int dynax_rc_loss_grad_ctrl_f32(const float *theta, const float *x0, const float *u_shared, const float *ctrl,
                                int ctrl_col, const float *target, const float *lb, const float *ub, float *loss,
                                float *g_theta, float *g_x0, int64_t N, int64_t T, float dt, uint32_t flags, void *ws,
                                size_t ws_bytes, void *stream) {
    DYNAX_API_RANGE();
    if (!ctrl && N > 0) return fail(DYNAX_ERR_INVALID_ARGUMENT, "ctrl must be non-NULL");
    return rc_adjoint<ADJ_MSE>(theta, x0, u_shared, target, lb, ub, nullptr, nullptr, loss, g_theta, g_x0, nullptr, N, T,
                               dt, flags | DYNAX_SHARED_U, ws, ws_bytes, static_cast<cudaStream_t>(stream), ctrl, ctrl_col);
}

int dynax_rc_vjp_f32(const float *theta, const float *x0, const float *u, const float *g_xsol, const float *g_ysol,
                     float *g_theta, float *g_x0, float *g_u, int64_t N, int64_t T, float dt, uint32_t flags,
                     void *ws, size_t ws_bytes, void *stream) {
    DYNAX_API_RANGE();
    return rc_adjoint<ADJ_VJP>(theta, x0, u, nullptr, nullptr, nullptr, g_xsol, g_ysol, nullptr, g_theta, g_x0, g_u, N,
                               T, dt, flags, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
