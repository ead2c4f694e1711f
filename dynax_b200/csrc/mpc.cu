// Sampling-based MPC: cost of N candidate control sequences through ONE shared 4R3C model
// (BASELINE.json config 4).  Each candidate is the rollout of DifferentiableSimulator
// (/root/reference/dynax/simulators/simulator.py:36-43) with inputs assembled as in the RC env
//     inputs_k = [d_k0, d_k1, u_k, d_k2, d_k3]                    (dynax/wrapper/envs/rc.py:131)
// and the per-step cost is the negated env reward (rc.py:227-252):
//     cost_k = w0 * price[h] * (-u_k/COP) * dt/3600 + w1 * (relu(y_k - T_high[h]) + relu(T_low[h] - y_k))
// with y_k = x_k[0] taken from the PRE-update state and the hour index h from the POST-update time
// (rc.py:135-144,238).
//
// One thread == one candidate; theta, x0, the disturbance series and the schedules are shared, so
// everything that does not depend on the candidate (B d_k, the price/comfort bands per step) is
// folded into a per-step table in shared memory once per CTA.  HBM traffic is 4 B per
// candidate-step (u) + 4/H (cost): the kernel sits at the FP32 issue roof, not the HBM roof.
#include <float.h>
#include <stdlib.h>
#include <string.h>

#include "api_common.cuh"
#include "models.cuh"

namespace dynax {

struct MpcParams {
    const float *theta;   // [7]
    const float *x0;      // [3]
    const float *d;       // [H,4]  disturbances: Tao, qCon_i, qRad_e, qRad_i
    const float *u;       // [H,N]  candidate HVAC heat flows
    const float *price, *t_low, *t_high;  // [24] hourly schedules
    float *cost;          // [N]
    unsigned long long *best;  // packed (ordered cost bits << 32 | candidate index) or NULL
    int64_t N, H;
    float dt, cop, w_energy, w_comfort, start_time;
};

struct MpcStep {  // per-step shared constants, 32 bytes
    float bd0, bd1, bd2;  // (B d_k) without the control column
    float ce;             // w0 * price[h] * (-1/COP) * dt/3600
    float lo, hi;         // comfort band at hour h
    float pad0, pad1;
};

__device__ __forceinline__ unsigned int order_bits(float f) {
    unsigned int b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);  // monotone map float -> uint
}

constexpr int kMpcThreads = 256;
constexpr int kMpcUnroll = 8;

// V candidates per thread.  V=4: one 16-byte load per thread and step (4 KB contiguous per CTA row instead of 1 KB)
// and one read of the step table per four candidates; needs N % 4 == 0 and 16-byte aligned u and cost.  Every candidate
// sees the same arithmetic as with V=1 (bit-identical costs).
template <int V>
struct MpcVec;
template <>
struct MpcVec<1> {
    using T = float;
    __device__ static void get(const T &v, float *o) { o[0] = v; }
    __device__ static T put(const float *o) { return o[0]; }
};
template <>
struct MpcVec<4> {
    using T = float4;
    __device__ static void get(const T &v, float *o) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
    __device__ static T put(const float *o) { return make_float4(o[0], o[1], o[2], o[3]); }
};

template <int V>
__global__ void __launch_bounds__(kMpcThreads) rc_mpc_cost_kernel(const MpcParams P) {
    using VT = typename MpcVec<V>::T;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MpcStep *tab = reinterpret_cast<MpcStep *>(smem_raw);
    const int64_t H = P.H, N = P.N;
    RCModel M;
    RCModel::Ptrs mp{P.theta, 0};
    M.load(mp, 0);
    for (int64_t k = threadIdx.x; k < H; k += blockDim.x) {
        const float *dk = P.d + k * 4;
        MpcStep s;
        // B @ [d0, d1, u, d2, d3] without the u column (RC.py:222-230)
        s.bd0 = M.B00 * dk[0] + M.B01 * dk[1];
        s.bd1 = M.B10 * dk[0] + M.B13 * dk[2];
        s.bd2 = M.B24 * dk[3];
        const float t_next = P.start_time + (float)(k + 1) * P.dt;      // rc.py:137 states.time + dt
        const long long ti = (long long)t_next;                         // rc.py:238, Python/jnp remainder (>= 0), h in [0,23]
        int h = (int)((float)(((ti % 86400) + 86400) % 86400) / 3600.0f);
        h = h < 0 ? 0 : (h > 23 ? 23 : h);
        s.ce = P.w_energy * (__ldg(P.price + h) * (P.dt / 3600.0f)) * (-1.0f / P.cop);
        s.lo = __ldg(P.t_low + h);
        s.hi = __ldg(P.t_high + h);
        s.pad0 = s.pad1 = 0.f;
        tab[k] = s;
    }
    __syncthreads();

    // first candidate of this thread; with V=4 the host guarantees N % 4 == 0, so a thread is all in or all out
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * V;
    const bool active = i < N;
    const int64_t ii = active ? i : (N - V);
    float x[V][3], cost[V];
#pragma unroll
    for (int c = 0; c < V; ++c) {
        x[c][0] = __ldg(P.x0); x[c][1] = __ldg(P.x0 + 1); x[c][2] = __ldg(P.x0 + 2);
        cost[c] = 0.f;
    }
    const float dt = P.dt, w1 = P.w_comfort;
    const float *up = P.u + ii;
    // one step of this thread's candidates; `s` = this step's shared constants
    auto step = [&](const MpcStep *s, const VT &uv) {
        const float4 a = *reinterpret_cast<const float4 *>(s);        // bd0 bd1 bd2 ce
        const float2 b = *reinterpret_cast<const float2 *>(&s->lo);   // lo hi
        float uk[V];
        MpcVec<V>::get(uv, uk);
#pragma unroll
        for (int c = 0; c < V; ++c) {
            float *xc = x[c];
            const float y = xc[0];  // pre-update output (simulator.py:38)
            const float dTz = fmaxf(y - b.y, 0.f) + fmaxf(b.x - y, 0.f);
            cost[c] = fmaf(a.w, uk[c], cost[c]);   // energy: w0*price*(-u/COP)*dt/3600
            cost[c] = fmaf(w1, dTz, cost[c]);
            // rhs = A x + B u as single FMA chains seeded with the shared (B d_k) term: 8 FMAs instead of 11 ops
            // (this kernel is bound by FP32 issue slots; the summation order differs from fxx(x)+fxu(u) by
            // rounding only, far inside the 1e-5 tolerance)
            const float r0 = fmaf(M.A02, xc[2], fmaf(M.A00, xc[0], fmaf(M.B02, uk[c], a.x)));
            const float r1 = fmaf(M.A12, xc[2], fmaf(M.A11, xc[1], a.y));
            const float r2 = fmaf(M.A22, xc[2], fmaf(M.A21, xc[1], fmaf(M.A20, xc[0], a.z)));
            xc[0] = fmaf(dt, r0, xc[0]);
            xc[1] = fmaf(dt, r1, xc[1]);
            xc[2] = fmaf(dt, r2, xc[2]);
        }
    };
    // full blocks of kMpcUnroll steps carry no per-step predicate (8 loads in flight, then 8 steps) ...
    const int64_t Hfull = H - (H % kMpcUnroll);
    for (int64_t k0 = 0; k0 < Hfull; k0 += kMpcUnroll) {
        VT uu[kMpcUnroll];
#pragma unroll
        for (int j = 0; j < kMpcUnroll; ++j) uu[j] = __ldcs(reinterpret_cast<const VT *>(up + (k0 + j) * N));
#pragma unroll
        for (int j = 0; j < kMpcUnroll; ++j) step(&tab[k0 + j], uu[j]);
    }
    // ... and the ragged tail goes step by step
    for (int64_t k = Hfull; k < H; ++k) step(&tab[k], __ldcs(reinterpret_cast<const VT *>(up + k * N)));
    if (active) *reinterpret_cast<VT *>(P.cost + i) = MpcVec<V>::put(cost);
    if (P.best != nullptr) {
        unsigned long long key = ~0ull;
#pragma unroll
        for (int c = 0; c < V; ++c) {
            const unsigned long long kc = ((unsigned long long)order_bits(cost[c]) << 32) | (unsigned long long)(uint32_t)(i + c);
            key = (active && kc < key) ? kc : key;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
            key = other < key ? other : key;
        }
        if ((threadIdx.x & 31) == 0) atomicMin(P.best, key);
    }
}

__global__ void rc_mpc_decode_kernel(const unsigned long long *best, const float *cost, long long *argmin,
                                     float *min_cost) {
    const unsigned long long key = *best;
    const long long idx = (long long)(key & 0xffffffffull);
    if (argmin) *argmin = idx;
    if (min_cost) *min_cost = cost[idx];
}

}  // namespace dynax

using namespace dynax;

extern "C" int dynax_rc_mpc_cost_f32(const float *theta, const float *x0, const float *d, const float *u,
                                     const float *price, const float *t_low, const float *t_high, float cop,
                                     float w_energy, float w_comfort, float start_time, float *cost,
                                     long long *argmin, float *min_cost, int64_t N, int64_t H, float dt, void *ws,
                                     size_t ws_bytes, void *stream) {
    DYNAX_API_RANGE();
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (N < 0 || H < 1) return fail(DYNAX_ERR_INVALID_ARGUMENT, "need N >= 0 and H >= 1");
    if (N == 0) return DYNAX_OK;
    if (!theta || !x0 || !d || !u || !price || !t_low || !t_high || !cost)
        return fail(DYNAX_ERR_INVALID_ARGUMENT, "theta,x0,d,u,price,t_low,t_high,cost must be non-NULL");
    if (N > 0xffffffffLL) return fail(DYNAX_ERR_INVALID_ARGUMENT, "N must fit 32 bits (argmin packing)");
    if (cop == 0.f) return fail(DYNAX_ERR_INVALID_ARGUMENT, "COP must be non-zero");
    const size_t smem = sizeof(MpcStep) * (size_t)H;
    if (smem > 200 * 1024) return fail(DYNAX_ERR_INVALID_ARGUMENT, "horizon too long for the shared-memory step table (H <= 6400)");
    const bool want_min = argmin != nullptr || min_cost != nullptr;
    if (want_min && (!ws || ws_bytes < 16 || !aligned16(ws)))
        return fail(DYNAX_ERR_WORKSPACE, "argmin needs a 16-byte aligned workspace of >= 16 bytes");
    int rc = require_sm100();
    if (rc != DYNAX_OK) return rc;
    MpcParams P;
    P.theta = theta; P.x0 = x0; P.d = d; P.u = u; P.price = price; P.t_low = t_low; P.t_high = t_high;
    P.cost = cost; P.best = want_min ? reinterpret_cast<unsigned long long *>(ws) : nullptr;
    P.N = N; P.H = H; P.dt = dt; P.cop = cop; P.w_energy = w_energy; P.w_comfort = w_comfort; P.start_time = start_time;
    if (want_min) DYNAX_CUDA_TRY(cudaMemsetAsync(ws, 0xff, 8, st));
    // four candidates per thread when the layout allows 16-byte accesses and there are enough candidates to keep
    // every SM busy with the wider CTAs (1024 candidates each); DYNAX_MPC_VEC=1 forces one candidate per thread
    DeviceInfo di;
    rc = device_info(&di);
    if (rc != DYNAX_OK) return rc;
    const char *e = getenv("DYNAX_MPC_VEC");
    const bool vec4 = (N % 4) == 0 && aligned16(u) && aligned16(cost) && N >= (int64_t)di.sm_count * 2 * 4 * kMpcThreads &&
                      !(e && atoi(e) == 1);
    if (vec4) {
        rc = set_smem(rc_mpc_cost_kernel<4>, smem);
        if (rc != DYNAX_OK) return rc;
        const int64_t grid = (N / 4 + kMpcThreads - 1) / kMpcThreads;
        rc_mpc_cost_kernel<4><<<(unsigned)grid, kMpcThreads, smem, st>>>(P);
    } else {
        rc = set_smem(rc_mpc_cost_kernel<1>, smem);
        if (rc != DYNAX_OK) return rc;
        const int64_t grid = (N + kMpcThreads - 1) / kMpcThreads;
        rc_mpc_cost_kernel<1><<<(unsigned)grid, kMpcThreads, smem, st>>>(P);
    }
    DYNAX_CUDA_TRY(cudaGetLastError());
    if (want_min) {
        rc_mpc_decode_kernel<<<1, 1, 0, st>>>(P.best, cost, argmin, min_cost);
        DYNAX_CUDA_TRY(cudaGetLastError());
    }
    return DYNAX_OK;
}
