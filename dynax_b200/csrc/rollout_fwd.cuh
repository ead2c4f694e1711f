// Forward rollout kernel: DifferentiableSimulator.__call__ for N systems at once
// (/root/reference/dynax/simulators/simulator.py:36-53 -- the nn.scan body and its lift).
//
// Mapping: one thread == one system; the model coefficients and the state live in registers
// for the whole horizon.  A CTA owns TN consecutive systems and streams the time axis:
//
//   producer warp (1 elected lane)  u[k0:k0+KF, n0:n0+TN, :]  --cp.async.bulk (TMA 1-D)-->  smem ring (S stages)
//   TN consumer threads             read their m inputs per step from smem (stride m words: conflict-free
//                                   for odd m), step the state, write x_{k+1}/y_k into an smem out-tile
//   producer warp                   out-tile --cp.async.bulk--> xsol/ysol rows (fully coalesced 16B-multiple rows)
//
// Synchronisation is mbarrier-only (full/empty per in-stage, ofull/ofree per out-stage); warps of a
// CTA drift freely within the ring.  If N is not a multiple of 4 (rows not 16-byte multiples) the
// producer warp falls back to plain coalesced ld/st for that launch -- same pipeline, no TMA.
#pragma once
#include <cuda.h>

#include "models.cuh"
#include "ptx.cuh"

namespace dynax {

template <class Model>
struct FwdParams {
    typename Model::Ptrs mp;
    const float *x0;  // [N,n]
    const float *u;   // [T,N,m] or [T,1,m]
    float *xsol;      // [T,N,n] or NULL
    float *ysol;      // [T,N,p] or NULL
    float *x_final;   // [N,n] or NULL
    int64_t N, T;
    float dt;
    int discrete;
    int bulk_in, bulk_out;
    // Time-chunked launches (gridDim.y = number of chunks; lti_chunked.cu): chunk c rolls steps
    // [c*chunk_len, min(T,(c+1)*chunk_len)) from x0 + c*x0_chunk_stride (or from zero) and writes its
    // final state to x_final + c*N*n.  chunk_len == 0: one chunk, the whole horizon.
    int64_t chunk_len = 0;
    int64_t x0_chunk_stride = 0;
    int x0_zero = 0;
    // Shared input series + ONE per-system channel (SHARED_U kernels only): u_k = u_shared[k] with column
    // ctrl_col replaced by ctrl[k, system].  This is how the RC env assembles its inputs
    // (dynax/wrapper/envs/rc.py:128-131: shared disturbance table + per-env control) and what
    // sampling-based control needs; it cuts the streamed input from 4m to 4 bytes per system-step.
    const float *ctrl = nullptr;  // [T,N] or NULL
    int ctrl_col = 0;
    // 3-D tensor maps over the row arrays (tmap.cuh): ONE cp.async.bulk.tensor per array and stage instead of one 1-D
    // bulk copy per row.  tensor_in: u (per-system) and ctrl; tensor_out: xsol and ysol.  inner_* = the maps' inner
    // extents in floats (the second coordinate of a CTA's box is n0*w/inner).
    int tensor_in = 0, tensor_out = 0, inner_u = 0, inner_c = 0, inner_x = 0, inner_y = 0;
    alignas(64) CUtensorMap map_u, map_c, map_x, map_y;
};

__host__ __device__ constexpr int align4i(int x) { return (x + 3) & ~3; }
__host__ __device__ constexpr int align32i(int x) { return (x + 31) & ~31; }   // 128 bytes: tensor-copy destinations

// SHIFT: the instantiation for per-system rows TMA cannot move (N % 4 != 0): rows are staged at their global offset modulo
// 16 bytes, so that the interior of a row moves as one bulk copy (ptx::copy_rows_shift_bulk), the row pitch grows by 4
// floats, and the computing threads add the (warp-uniform) shift of each row to their shared address.
// NO: out-tiles (x and y of one stage each).  Two suffice when several CTAs share an SM; a launch of one CTA per SM is a
// serial chain that stalls whenever the TMA unit has not finished reading the tile written two stages ago: four tiles.
template <class Model, int TN, int KF, int S, bool SHARED_U, bool SHIFT = false, int NO = 2>
struct FwdCfg {
    static_assert(!SHIFT || !SHARED_U, "shifted rows are per-system rows");
    static_assert(!SHIFT || KF <= 32, "copy_rows_shift_bulk: one lane per row of a stage");
    static constexpr int n = Model::n, m = Model::m, p = Model::p;
    static constexpr int IN_ROW = SHARED_U ? m : TN * m + (SHIFT ? 4 : 0);   // row pitch in the stage
    static constexpr int CTRL_OFF = align32i(KF * IN_ROW);              // per-system channel rows [KF][TN]
    static constexpr int IN_STAGE = align32i(CTRL_OFF + (SHARED_U ? KF * TN : 0));
    static constexpr int XO_STAGE = KF * TN * n;
    static constexpr int YO_STAGE = KF * TN * p;
    static constexpr int NBAR = 2 * S + 2 * NO;
    static constexpr int MODEL_FLOATS = align4i(Model::SMEM_FLOATS);  // block-shared model data (MLP weights)
    static constexpr size_t SMEM_BYTES =
        sizeof(float) * (size_t)(S * IN_STAGE + NO * XO_STAGE + NO * YO_STAGE + MODEL_FLOATS) + sizeof(uint64_t) * NBAR;
    static constexpr int THREADS = TN + 32;
};

template <class Model, int TN, int KF, int S, bool SHARED_U, int MINB, bool SHIFT = false, int NO = 2>
__global__ void __launch_bounds__(TN + 32, MINB) rollout_fwd_kernel(const __grid_constant__ FwdParams<Model> P) {
    using Cfg = FwdCfg<Model, TN, KF, S, SHARED_U, SHIFT, NO>;
    static_assert(NO >= 2, "at least two out-tiles");
    constexpr int n = Cfg::n, m = Cfg::m, p = Cfg::p;
    constexpr int IN_ROW = Cfg::IN_ROW, IN_STAGE = Cfg::IN_STAGE, XO_STAGE = Cfg::XO_STAGE,
                  YO_STAGE = Cfg::YO_STAGE;
    static_assert((TN % 32) == 0 && (S & (S - 1)) == 0, "TN multiple of 32, S power of two");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *s_in = reinterpret_cast<float *>(smem_raw);
    float *s_xo = s_in + S * IN_STAGE;
    float *s_yo = s_xo + NO * XO_STAGE;
    float *s_model = s_yo + NO * YO_STAGE;
    uint64_t *full = reinterpret_cast<uint64_t *>(s_model + Cfg::MODEL_FLOATS);
    uint64_t *empty = full + S;
    uint64_t *ofull = empty + S;
    uint64_t *ofree = ofull + NO;

    const int tid = threadIdx.x;
    const int64_t N = P.N;
    const int64_t n0 = (int64_t)blockIdx.x * TN;
    const int valid = (int)((N - n0) < TN ? (N - n0) : TN);
    // this CTA's slice of the time axis (the whole horizon unless the launch is time-chunked)
    const int64_t t_begin = P.chunk_len > 0 ? (int64_t)blockIdx.y * P.chunk_len : 0;
    const int64_t T = P.chunk_len > 0 ? ((P.T - t_begin) < P.chunk_len ? (P.T - t_begin) : P.chunk_len) : P.T;
    const float *const u_base = P.u + t_begin * (SHARED_U ? m : N * m);
    float *const xsol_base = P.xsol ? P.xsol + t_begin * N * n : nullptr;
    float *const ysol_base = P.ysol ? P.ysol + t_begin * N * p : nullptr;
    const bool emit_x = P.xsol != nullptr, emit_y = P.ysol != nullptr;
    // outputs that TMA cannot move (N % 4 != 0 or unaligned bases) are stored by the computing threads themselves,
    // straight from registers: no out-tile, no hand-shake with the producer warp (which would otherwise copy every row
    // with its 32 lanes)
    const bool emit = (emit_x || emit_y) && P.bulk_out;
    const bool direct = (emit_x || emit_y) && !P.bulk_out;
    const int64_t C = (T + KF - 1) / KF;

    // per-system input rows that TMA cannot move (N % 4 != 0) go through 4-byte cp.async from all producer lanes:
    // every lane then arrives on the stage's barrier (count 32) instead of one elected lane
    const bool async_rows = !P.bulk_in && (!SHARED_U || P.ctrl != nullptr);
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < S; ++s) {
            ptx::mbar_init(&full[s], async_rows ? 32 : 1);
            ptx::mbar_init(&empty[s], TN / 32);
        }
#pragma unroll
        for (int o = 0; o < NO; ++o) {
            ptx::mbar_init(&ofull[o], TN / 32);
            ptx::mbar_init(&ofree[o], 1);
        }
        ptx::fence_mbar_init();
    }
    Model::block_init(s_model, P.mp, tid, TN + 32);  // no-op unless the model keeps shared data
    __syncthreads();

    if (tid >= TN) {
        // ------------------------------- producer warp -------------------------------
        const int lane = tid - TN;
        const uint64_t pol = ptx::policy_evict_first(), pol_keep = ptx::policy_evict_last();

        auto load_chunk = [&](int64_t c) {
            const int s = (int)(c % S);
            const int64_t k0 = c * KF;
            const int rows = (int)((T - k0) < KF ? (T - k0) : KF);
            float *dst = s_in + s * IN_STAGE;
            if (SHARED_U) {
                const float *src = u_base + k0 * m;  // [T,1,m] is contiguous in k
                // the shared rows of a chunk are one contiguous block: a single bulk copy when it is 16-byte
                // aligned and sized (asynchronous, so the producer runs S chunks ahead instead of paying an L2
                // round trip per chunk), plain loads otherwise (ragged last chunk, odd offsets)
                const uint32_t sh_bytes = (uint32_t)(rows * m) * 4u;
                const bool sh_bulk = (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (sh_bytes & 15u) == 0;
                const float *csrc = P.ctrl ? P.ctrl + (t_begin + k0) * N + n0 : nullptr;
                float *cdst = dst + Cfg::CTRL_OFF;
                if (async_rows) {  // ctrl rows are not TMA-able: everything the lanes move goes through cp.async
                    if (!sh_bulk) ptx::copy_rows_async(dst, 0, src, 0, 1, rows * m, lane);
                    else if (lane == 0) {
                        ptx::mbar_expect_tx(&full[s], sh_bytes);
                        ptx::bulk_g2s(dst, src, sh_bytes, &full[s], pol_keep);
                    }
                    ptx::copy_rows_async(cdst, TN, csrc, N, rows, valid, lane);
                    ptx::cp_async_arrive(&full[s]);
                    return;
                }
                if (!sh_bulk)
                    for (int i = lane; i < rows * m; i += 32) dst[i] = __ldg(src + i);
                const bool c_bulk = csrc != nullptr;  // (not async_rows => bulk_in whenever there is a ctrl channel)
                __syncwarp();
                const uint32_t row_bytes = (uint32_t)valid * 4u;
                if (lane == 0) {
                    if (sh_bulk || c_bulk) {
                        const uint32_t c_bytes = P.tensor_in ? (uint32_t)(KF * TN * 4) : row_bytes * (uint32_t)rows;
                        ptx::mbar_arrive_expect_tx(&full[s], (sh_bulk ? sh_bytes : 0u) + (c_bulk ? c_bytes : 0u));
                        if (sh_bulk) ptx::bulk_g2s(dst, src, sh_bytes, &full[s], pol_keep);
                        if (c_bulk && P.tensor_in)   // the whole stage of the channel in one tensor copy
                            ptx::tensor_g2s_3d(cdst, &P.map_c, 0, (int)(n0 / P.inner_c), (int)(t_begin + k0), &full[s], pol);
                    } else {
                        ptx::mbar_arrive(&full[s]);
                    }
                }
                __syncwarp();
                if (c_bulk && !P.tensor_in && lane < rows)   // per-row copies: lane r issues row r (they leave together)
                    ptx::bulk_g2s(cdst + lane * TN, csrc + lane * N, row_bytes, &full[s], pol);
            } else if (P.tensor_in) {
                // one tensor copy for the stage: KF rows x this CTA's slice (rows past the horizon and systems past N
                // arrive as zeros; the transaction count is always the full box)
                if (lane == 0) {
                    ptx::mbar_arrive_expect_tx(&full[s], (uint32_t)(KF * TN * m * 4));
                    ptx::tensor_g2s_3d(dst, &P.map_u, 0, (int)(n0 * m / P.inner_u), (int)(t_begin + k0), &full[s], pol);
                }
            } else if (P.bulk_in) {
                // per-row copies (N % 4 == 0 but no tensor map for this layout): lane r issues row r -- one thread gets a
                // bulk copy out every ~45 ns only (profiles/r02_tma_probe.txt), the lanes' copies leave together
                const uint32_t row_bytes = (uint32_t)valid * m * 4u;
                if (lane == 0) ptx::mbar_arrive_expect_tx(&full[s], row_bytes * (uint32_t)rows);
                __syncwarp();
                if (lane < rows) ptx::bulk_g2s(dst + lane * IN_ROW, u_base + ((k0 + lane) * N + n0) * m, row_bytes, &full[s], pol);
            } else {
                if constexpr (SHIFT) ptx::copy_rows_shift_bulk(dst, IN_ROW, u_base + (k0 * N + n0) * m, N * m, rows, valid * m, lane, &full[s], pol);
                else ptx::copy_rows_async(dst, IN_ROW, u_base + (k0 * N + n0) * m, N * m, rows, valid * m, lane);
                ptx::cp_async_arrive(&full[s]);
            }
        };

        // The tile of chunk c is handed back once the TMA unit has READ it.  Two tiles: right after the store is issued.
        // More: one chunk later (the newest store group stays in flight), so the producer never sits in the wait.
        constexpr int LAG = NO > 2 ? 1 : 0;
        auto store_chunk = [&](int64_t c) {
            const int o = (int)(c % NO);
            const int64_t k0 = c * KF;
            const int rows = (int)((T - k0) < KF ? (T - k0) : KF);
            const float *sx = s_xo + o * XO_STAGE;
            const float *sy = s_yo + o * YO_STAGE;
            if (P.tensor_out && rows == KF) {
                // full stage: one tensor store per array (systems past N are clipped by the map).  A ragged stage goes
                // row by row: in a time-chunked launch the rows behind it belong to the next chunk's CTA.
                if (lane == 0) {
                    if (emit_x) ptx::tensor_s2g_3d(&P.map_x, 0, (int)(n0 * n / P.inner_x), (int)(t_begin + k0), sx);
                    if (emit_y) ptx::tensor_s2g_3d(&P.map_y, 0, (int)(n0 * p / P.inner_y), (int)(t_begin + k0), sy);
                    ptx::bulk_commit();
                    ptx::bulk_wait_read<LAG>();  // out-tile may be overwritten once TMA has read it
                }
            } else {
                // per-row stores: lane r stores row r (bulk groups are per thread: every lane waits for its own)
                if (lane < rows) {
                    if (emit_x)
                        ptx::bulk_s2g(xsol_base + ((k0 + lane) * N + n0) * n, sx + lane * TN * n, (uint32_t)valid * n * 4u);
                    if (emit_y)
                        ptx::bulk_s2g(ysol_base + ((k0 + lane) * N + n0) * p, sy + lane * TN * p, (uint32_t)valid * p * 4u);
                }
                ptx::bulk_commit();   // every lane, also without a row: the group counts stay in step
                ptx::bulk_wait_read<LAG>();
            }
            __syncwarp();
            if (lane == 0 && c >= LAG) ptx::mbar_arrive(&ofree[(c - LAG) % NO]);
        };

        for (int64_t c = 0; c < C && c < S; ++c) load_chunk(c);
        for (int64_t c = 0; c < C; ++c) {
            const int s = (int)(c % S);
            ptx::mbar_wait(&empty[s], (uint32_t)((c / S) & 1));  // consumers are done with chunk c
            if (c + S < C) load_chunk(c + S);
            if (emit) {
                ptx::mbar_wait(&ofull[c % NO], (uint32_t)((c / NO) & 1));
                store_chunk(c);
            }
        }
        ptx::bulk_wait<0>();  // every lane: its own store groups have landed
        return;
    }

    // --------------------------------- consumers ---------------------------------
    const int lane = tid & 31;
    const bool active = tid < valid;
    const int64_t i = active ? (n0 + tid) : (N - 1);  // idle lanes shadow the last system; never stored
    Model M;
    M.load(P.mp, i, s_model);
    float x[n];
#pragma unroll
    for (int d = 0; d < n; ++d) x[d] = P.x0_zero ? 0.f : __ldg(P.x0 + (int64_t)blockIdx.y * P.x0_chunk_stride + i * n + d);
    const float dt = P.dt;
    const bool discrete = P.discrete != 0;

    for (int64_t c = 0; c < C; ++c) {
        const int s = (int)(c % S);
        const int o = (int)(c % NO);
        const int64_t k0 = c * KF;
        const int rows = (int)((T - k0) < KF ? (T - k0) : KF);
        ptx::mbar_wait(&full[s], (uint32_t)((c / S) & 1));
        if (emit) ptx::mbar_wait(&ofree[o], (uint32_t)(((c / NO) & 1) ^ 1));
        const float *in = s_in + s * IN_STAGE + (SHARED_U ? 0 : tid * m);
        // SHIFT: row j of this stage sits at its global offset modulo 16 bytes (see ptx::copy_rows_shift_bulk)
        const unsigned sh0 = SHIFT ? (unsigned)((reinterpret_cast<uintptr_t>(u_base) >> 2) + (uint64_t)((k0 * N + n0) * m)) : 0u;
        const unsigned shd = SHIFT ? (unsigned)(N * m) : 0u;
        float *xo = s_xo + o * XO_STAGE + tid * n;
        float *yo = s_yo + o * YO_STAGE + tid * p;
#pragma unroll
        for (int j = 0; j < KF; ++j) {
            if (j < rows) {
                float u[m];
#pragma unroll
                for (int d = 0; d < m; ++d) u[d] = in[j * IN_ROW + (SHIFT ? (int)((sh0 + j * shd) & 3u) : 0) + d];
                if (SHARED_U && P.ctrl != nullptr) {
                    const float cv = s_in[s * IN_STAGE + Cfg::CTRL_OFF + j * TN + tid];
#pragma unroll
                    for (int d = 0; d < m; ++d) u[d] = (d == P.ctrl_col) ? cv : u[d];
                }
                if (emit_y) {
                    float y[p];
                    M.out(x, u, y);  // y_k from the PRE-update state (simulator.py:38)
                    if (direct) {
                        if (active) {
#pragma unroll
                            for (int d = 0; d < p; ++d) ysol_base[((k0 + j) * N + i) * p + d] = y[d];
                        }
                    } else {
#pragma unroll
                        for (int d = 0; d < p; ++d) yo[j * TN * p + d] = y[d];
                    }
                }
                M.step(x, u, dt, discrete);  // explicit Euler x += dt*rhs (simulator.py:40) unless the model says otherwise
                if (emit_x) {
                    if (direct) {
                        if (active) {
#pragma unroll
                            for (int d = 0; d < n; ++d) xsol_base[((k0 + j) * N + i) * n + d] = x[d];
                        }
                    } else {
#pragma unroll
                        for (int d = 0; d < n; ++d) xo[j * TN * n + d] = x[d];  // x_{k+1} (simulator.py:43)
                    }
                }
            }
        }
        if (emit) ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
            ptx::mbar_arrive(&empty[s]);
            if (emit) ptx::mbar_arrive(&ofull[o]);
        }
    }
    if (P.x_final != nullptr && active) {
#pragma unroll
        for (int d = 0; d < n; ++d) P.x_final[(P.chunk_len > 0 ? (int64_t)blockIdx.y * N * n : 0) + i * n + d] = x[d];
    }
}

}  // namespace dynax
