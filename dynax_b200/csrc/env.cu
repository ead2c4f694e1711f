// Batched RC-env step (row f-2 of SURVEY.md section 8): K consecutive `env.step` calls for N
// independent environments in one launch.
//
// Follows /root/reference/dynax/wrapper/envs/rc.py:
//   :166      control = action*(u_high-u_low)/(num_actions-1) + u_low
//   :128      dist    = Tabular(ts, xs, 'linear')(states.time)           (4 disturbance columns)
//   :131-132  inputs  = [dist0, dist1, control, dist2, dist3];  x_next = x + dt*(A x + B u)  (one Euler step)
//   :135-138  time   += dt
//   :200-222  obs     = [time_next, y (PRE-update x[0]), dist0, dist3, -control/COP]
//   :238-250  h = int((int(obs0) % 86400)/3600); energy = obs4*dt/3600; cost = price[h]*energy;
//             dTz = relu(obs1 - T_high[h]) + relu(T_low[h] - obs1);  reward = -w0*cost - w1*dTz
//   :255-264  terminated = 1 if sum(relu(x_next-40)) + sum(relu(12-x_next)) > 0 else 0
// Pinned by the reference's known-answer test tests/test_gymwrapper_basics.py:7-12.
//
// One thread == one environment; model coefficients and state in registers across the K steps.
#include "api_common.cuh"
#include "models.cuh"

namespace dynax {

struct EnvParams {
    const float *theta;  // [7] or [N,7]
    int64_t theta_stride;
    const float *x, *t;        // [N,3], [N]
    const int32_t *action;     // [K,N]
    const float *ts, *dist;    // [R], [R,4]
    int64_t R;
    const float *price, *t_low, *t_high;  // [24]
    float cop, w_energy, w_comfort, u_low, u_high;
    int num_actions;
    float *x_next, *t_next;               // [N,3], [N]
    float *obs, *reward, *terminated;     // [K,N,5], [K,N], [K,N]
    float *cost, *dTz;                    // [K,N] or NULL
    int64_t N, K;
    float dt;
};

// jnp.interp(t, ts, arange(R)) -> fractional row index, with a hint from the previous lookup
__device__ __forceinline__ float frac_index(const float *ts, int64_t R, float t, int64_t &hint) {
    if (!(t > ts[0])) return 0.f;
    if (!(t < ts[R - 1])) return (float)(R - 1);
    int64_t lo = hint;
    if (!(lo >= 0 && lo + 1 < R && ts[lo] <= t && t < ts[lo + 1])) {
        if (lo + 2 < R && lo + 1 >= 0 && ts[lo + 1] <= t && t < ts[lo + 2]) {
            lo = lo + 1;
        } else {
            int64_t a = 0, b = R - 1;
            while (b - a > 1) {
                const int64_t mid = (a + b) >> 1;
                if (ts[mid] <= t) a = mid; else b = mid;
            }
            lo = a;
        }
    }
    hint = lo;
    const float w = ts[lo + 1] - ts[lo];
    return (float)lo + (w > 0.f ? (t - ts[lo]) / w : 0.f);
}

// rc.py:238  h = int((int(t) % 86400) / 3600) with Python/jnp semantics: the remainder is non-negative also for
// negative times (C's % is not), times beyond 2^31 s do not overflow, and the table index stays in [0, 23]
__device__ __forceinline__ int hour_of_day(float t) {
    const long long ti = (long long)t;
    const long long r = ((ti % 86400) + 86400) % 86400;
    const int h = (int)((float)r / 3600.0f);
    return h < 0 ? 0 : (h > 23 ? 23 : h);
}

__global__ void rc_env_step_kernel(const EnvParams P) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.N) return;
    RCModel M;
    RCModel::Ptrs mp{P.theta, P.theta_stride};
    M.load(mp, i);
    float x[3] = {P.x[i * 3], P.x[i * 3 + 1], P.x[i * 3 + 2]};
    float t = P.t[i];
    int64_t hint = 0;
    for (int64_t k = 0; k < P.K; ++k) {
        // rc.py:166  action*(u_high-u_low)/(num_actions-1)+u_low, in the reference's order of operations
        const float control = ((float)P.action[k * P.N + i] * (P.u_high - P.u_low)) / (float)(P.num_actions - 1) + P.u_low;
        const float idx = frac_index(P.ts, P.R, t, hint);                               // interpolate.py:63
        const float fl = floorf(idx);
        int64_t r0 = (int64_t)fl;
        if (r0 > P.R - 1) r0 = P.R - 1;
        const int64_t r1 = r0 + 1 < P.R ? r0 + 1 : P.R - 1;
        const float w = idx - fl;
        float d[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) d[c] = P.dist[r0 * 4 + c] * (1.f - w) + P.dist[r1 * 4 + c] * w;  // order-1 map_coordinates
        const float u[5] = {d[0], d[1], control, d[2], d[3]};                           // rc.py:131
        const float y = x[0];                                                           // pre-update output
        M.step(x, u, P.dt, false);                                                      // simulator.py:38-40
        t = t + P.dt;                                                                   // rc.py:137
        const float power = -control / P.cop;                                           // rc.py:221
        const int h = hour_of_day(t);                                                   // rc.py:238
        const float energy = power * P.dt / 3600.0f;                                    // rc.py:241
        const float cst = P.price[h] * energy;                                          // rc.py:244
        const float dT = fmaxf(y - P.t_high[h], 0.f) + fmaxf(P.t_low[h] - y, 0.f);      // rc.py:247
        const float rew = -P.w_energy * cst - P.w_comfort * dT;                         // rc.py:250
        const float viol = (fmaxf(x[0] - 40.f, 0.f) + fmaxf(x[1] - 40.f, 0.f) + fmaxf(x[2] - 40.f, 0.f)) +
                           (fmaxf(12.f - x[0], 0.f) + fmaxf(12.f - x[1], 0.f) + fmaxf(12.f - x[2], 0.f));  // rc.py:258
        const int64_t o = k * P.N + i;
        float *ob = P.obs + o * 5;
        ob[0] = t; ob[1] = y; ob[2] = d[0]; ob[3] = d[3]; ob[4] = power;                // rc.py:218-222
        P.reward[o] = rew;
        P.terminated[o] = viol > 0.f ? 1.f : 0.f;
        if (P.cost) P.cost[o] = cst;
        if (P.dTz) P.dTz[o] = dT;
    }
    P.x_next[i * 3] = x[0]; P.x_next[i * 3 + 1] = x[1]; P.x_next[i * 3 + 2] = x[2];
    P.t_next[i] = t;
}

}  // namespace dynax

using namespace dynax;

extern "C" int dynax_rc_env_step_f32(const float *theta, const float *x, const float *t, const int32_t *action,
                                     const float *ts, const float *dist, int64_t R, const float *price,
                                     const float *t_low, const float *t_high, float cop, float w_energy,
                                     float w_comfort, float u_low, float u_high, int num_actions, float *x_next,
                                     float *t_next, float *obs, float *reward, float *terminated, float *cost,
                                     float *dTz, int64_t N, int64_t K, float dt, uint32_t flags, void *stream) {
    DYNAX_API_RANGE();
    if (N < 0 || K < 1) return fail(DYNAX_ERR_INVALID_ARGUMENT, "need N >= 0 and K >= 1");
    if (N == 0) return DYNAX_OK;
    if (!theta || !x || !t || !action || !ts || !dist || !price || !t_low || !t_high || !x_next || !t_next || !obs ||
        !reward || !terminated)
        return fail(DYNAX_ERR_INVALID_ARGUMENT, "NULL argument");
    if (R < 1 || num_actions < 2 || cop == 0.f) return fail(DYNAX_ERR_INVALID_ARGUMENT, "need R >= 1, num_actions >= 2, COP != 0");
    if (flags & ~DYNAX_SHARED_THETA) return fail(DYNAX_ERR_INVALID_ARGUMENT, "unsupported flags 0x%x for rc_env_step", flags);
    int rc = require_sm100();
    if (rc != DYNAX_OK) return rc;
    EnvParams P;
    P.theta = theta; P.theta_stride = (flags & DYNAX_SHARED_THETA) ? 0 : 7;
    P.x = x; P.t = t; P.action = action; P.ts = ts; P.dist = dist; P.R = R;
    P.price = price; P.t_low = t_low; P.t_high = t_high;
    P.cop = cop; P.w_energy = w_energy; P.w_comfort = w_comfort; P.u_low = u_low; P.u_high = u_high;
    P.num_actions = num_actions;
    P.x_next = x_next; P.t_next = t_next; P.obs = obs; P.reward = reward; P.terminated = terminated;
    P.cost = cost; P.dTz = dTz; P.N = N; P.K = K; P.dt = dt;
    const int tb = 128;
    rc_env_step_kernel<<<(unsigned)((N + tb - 1) / tb), tb, 0, static_cast<cudaStream_t>(stream)>>>(P);
    DYNAX_CUDA_TRY(cudaGetLastError());
    return DYNAX_OK;
}
