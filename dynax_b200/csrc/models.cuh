// Per-system model policies: coefficients live in registers, one thread == one system.
//
//   RCModel            Continuous4R3C / Discrete4R3C     /root/reference/dynax/models/RC.py:21-249
//   DenseModel<n,m,p>  Discrete/ContinuousLinearSSM      dynax/core/{discrete,continuous}_block_state_space.py:5-54
//
// Both implement the split form of BaseBlockSSM (dynax/core/base_block_state_space.py:59-71):
//   rhs = fxx(x) + fxu(u) = A x + B u          y = fyx(x) + fyu(u) = C x + D u
// and the pieces of its transpose that the discrete adjoint needs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dynax {

struct RCModel {
    static constexpr int n = 3, m = 5, p = 1;
    static constexpr int NG = 7;    // accumulators: one per parameter (see accum)
    static constexpr int NOUT = 7;  // d/d[Cai,Cwe,Cwi,Re,Ri,Rw,Rg]
    static constexpr bool HAS_BOX = true;  // parameter box penalty of examples/3-inverse.py:66-85,103-107

    struct Ptrs {
        const float *theta;  // [N,7] or [1,7]
        int64_t stride;      // 7, or 0 when shared
    };
    struct GradPtrs {
        float *g_theta;  // [N,7] (or [7] when shared)
    };

    float A00, A02, A11, A12, A20, A21, A22;
    float B00, B01, B02, B10, B13, B24;
    float iRe, iRi, iRw, iRg;  // reciprocals of the resistances (gradient accumulation only)

    // theta -> A, B with the reference's expression shapes so coefficient rounding matches
    // (RC.py:203-209, 224-229).  IEEE division: the library is built without fast-math.
    __device__ __forceinline__ void load(const Ptrs &P, int64_t i) {
        const float *th = P.theta + i * P.stride;
        const float Cai = __ldg(th + 0), Cwe = __ldg(th + 1), Cwi = __ldg(th + 2), Re = __ldg(th + 3),
                    Ri = __ldg(th + 4), Rw = __ldg(th + 5), Rg = __ldg(th + 6);
        A00 = -1.0f / Cai * (1.0f / Rg + 1.0f / Ri);
        A02 = 1.0f / (Cai * Ri);
        A11 = -1.0f / Cwe * (1.0f / Re + 1.0f / Rw);
        A12 = 1.0f / (Cwe * Rw);
        A20 = 1.0f / (Cwi * Ri);
        A21 = 1.0f / (Cwi * Rw);
        A22 = -1.0f / Cwi * (1.0f / Rw + 1.0f / Ri);
        B00 = 1.0f / (Cai * Rg);
        B01 = 1.0f / Cai;
        B02 = 1.0f / Cai;
        B10 = 1.0f / (Cwe * Re);
        B13 = 1.0f / Cwe;
        B24 = 1.0f / Cwi;
        iRe = 1.0f / Re; iRi = 1.0f / Ri; iRw = 1.0f / Rw; iRg = 1.0f / Rg;
    }

    // rhs = A x + B u  (structural zeros of A and B skipped: x + 0*y == x exactly for finite y).
    // The order of operations is spelled out with intrinsics (nvcc contracts a*b + c*d differently from one
    // inlining context to the next): every kernel that steps this model -- forward, adjoint, tangent (scalar and
    // packed) -- then produces bit-identical trajectories.  Each dot product = one product, then FMAs; the two
    // dot products are added last (fxx(x) + fxu(u), base_block_state_space.py:61).
    __device__ __forceinline__ void rhs(const float *x, const float *u, float *r) const {
        const float bu0 = fmaf(B02, u[2], fmaf(B00, u[0], __fmul_rn(B01, u[1])));
        const float bu1 = fmaf(B10, u[0], __fmul_rn(B13, u[3]));
        const float bu2 = __fmul_rn(B24, u[4]);
        const float ax0 = fmaf(A00, x[0], __fmul_rn(A02, x[2]));
        const float ax1 = fmaf(A11, x[1], __fmul_rn(A12, x[2]));
        const float ax2 = fmaf(A22, x[2], fmaf(A20, x[0], __fmul_rn(A21, x[1])));
        r[0] = __fadd_rn(ax0, bu0);
        r[1] = __fadd_rn(ax1, bu1);
        r[2] = __fadd_rn(ax2, bu2);
    }


    // ---- hooks used by the rollout kernels ----
    static constexpr int SMEM_FLOATS = 0;  // no block-shared model data
    __device__ __forceinline__ static void block_init(float *, const Ptrs &, int, int) {}
    __device__ __forceinline__ void load(const Ptrs &P, int64_t i, const float *) { load(P, i); }
    // explicit Euler in residual form (simulator.py:40); `discrete` steps x_{k+1} = rhs instead
    __device__ __forceinline__ void step(float *x, const float *u, float dt, bool discrete) const {
        float r[n];
        rhs(x, u, r);
#pragma unroll
        for (int d = 0; d < n; ++d) x[d] = discrete ? r[d] : fmaf(dt, r[d], x[d]);
    }

    // y = C x + D u with C=[1,0,0], D=0  (RC.py:236-249)
    __device__ __forceinline__ void out(const float *x, const float *, float *y) const { y[0] = x[0]; }

    // Gradient accumulators (s = cotangent of rhs), one per parameter:
    //   capacities   d rhs_i / d C_i = -rhs_i / C_i            ->  S s0 rhs0,  S s1 rhs1,  S s2 rhs2   with rhs AS COMPUTED
    //                by rhs() -- the increments the stored trajectory actually took; a more accurate re-evaluation
    //                of rhs is LESS accurate here (2.8e-4 vs 2e-5 on a year-long case: these sums cancel over time
    //                and must stay consistent with the trajectory they are evaluated on);
    //   resistances  in DIFFERENCE form:  S s1 (u0-x1) [Re],  S (s0/Cai - s2/Cwi)(x2-x0) [Ri],
    //                S (s1/Cwe - s2/Cwi)(x2-x1) [Rw],  S s0 (u0-x0) [Rg].
    // Accumulating dL/dA = S s x^T and dL/dB = S s u^T entry by entry and combining them afterwards (what XLA's
    // transpose of the scan does) subtracts two large fp32 sums -- dL/dRg ~ dL/dA00 - dL/dB00 = S s0 x0 - S s0 u0 --
    // and loses |x0| / |x0 - u0| in relative accuracy (7e-4 on a case with outdoor ~ zone temperature).
    __device__ __forceinline__ void accum(float *G, const float *s, const float *, const float *x,
                                          const float *u) const {
        float r[3];
        rhs(x, u, r);
        const float dg = u[0] - x[0], de = u[0] - x[1], di = x[2] - x[0], dw = x[2] - x[1];
        const float t0 = s[0] * B01, t1 = s[1] * B13, t2 = s[2] * B24;  // s_i / C_i
        G[0] = fmaf(s[0], r[0], G[0]);
        G[1] = fmaf(s[1], r[1], G[1]);
        G[2] = fmaf(s[2], r[2], G[2]);
        G[3] = fmaf(s[1], de, G[3]);
        G[4] = fmaf(t0 - t2, di, G[4]);
        G[5] = fmaf(t1 - t2, dw, G[5]);
        G[6] = fmaf(s[0], dg, G[6]);
    }

    // ax = A^T s + C^T gy
    __device__ __forceinline__ void adj_x(const float *s, const float *gy, float *ax) const {
        ax[0] = (A00 * s[0] + A20 * s[2]) + gy[0];
        ax[1] = A11 * s[1] + A21 * s[2];
        ax[2] = A02 * s[0] + A12 * s[1] + A22 * s[2];
    }

    // gu = B^T s + D^T gy
    __device__ __forceinline__ void adj_u(const float *s, const float *, float *gu) const {
        gu[0] = B00 * s[0] + B10 * s[1];
        gu[1] = B01 * s[0];
        gu[2] = B02 * s[0];
        gu[3] = B13 * s[1];
        gu[4] = B24 * s[2];
    }

    // dL/dtheta from the seven sums above (derivatives of RC.py:203-209,224-229), in fp64.
    __device__ __forceinline__ void finalize(const double *G, const Ptrs &P, int64_t i, double *g) const {
        const float *th = P.theta + i * P.stride;
        const double cai = 1.0 / (double)th[0], cwe = 1.0 / (double)th[1], cwi = 1.0 / (double)th[2];
        const double re = 1.0 / (double)th[3], ri = 1.0 / (double)th[4], rw = 1.0 / (double)th[5], rg = 1.0 / (double)th[6];
        g[0] = -cai * G[0];
        g[1] = -cwe * G[1];
        g[2] = -cwi * G[2];
        g[3] = -cwe * re * re * G[3];
        g[4] = -ri * ri * G[4];
        g[5] = -rw * rw * G[5];
        g[6] = -cai * rg * rg * G[6];
    }

    __device__ __forceinline__ static float *grad_addr(const GradPtrs &Gp, int64_t i, int j) {
        return Gp.g_theta ? Gp.g_theta + i * 7 + j : nullptr;
    }
};

template <int N_, int M_, int P_>
struct DenseModel {
    static constexpr int n = N_, m = M_, p = P_;
    static constexpr int NG = n * n + n * m + p * n + p * m;
    static constexpr int NOUT = NG;
    static constexpr bool HAS_BOX = false;

    struct Ptrs {
        const float *A, *B, *C, *D;  // [N|1, rows, cols] row-major matrices
        int64_t shared;              // 1 -> one set for all systems
    };
    struct GradPtrs {
        float *gA, *gB, *gC, *gD;  // each may be NULL
    };

    float A[n * n], B[n * m], C[p * n], D[p * m];

    __device__ __forceinline__ void load(const Ptrs &P, int64_t i) {
        const int64_t s = P.shared ? 0 : i;
#pragma unroll
        for (int k = 0; k < n * n; ++k) A[k] = __ldg(P.A + s * (n * n) + k);
#pragma unroll
        for (int k = 0; k < n * m; ++k) B[k] = __ldg(P.B + s * (n * m) + k);
#pragma unroll
        for (int k = 0; k < p * n; ++k) C[k] = __ldg(P.C + s * (p * n) + k);
#pragma unroll
        for (int k = 0; k < p * m; ++k) D[k] = __ldg(P.D + s * (p * m) + k);
    }

    // Even state and output counts: the gradient kernel pairs the ROWS of all four matrices (rhs through rows2, the
    // transposed products through adj_rows), so that one register layout serves its recomputation and its adjoint step.
    static constexpr bool ROWPAIR = (n % 2 == 0) && (p % 2 == 0);

    __device__ __forceinline__ void rhs(const float *x, const float *u, float *r) const {
        if constexpr (ROWPAIR) {
            rows2<n>(A, B, x, u, r);  // the same operations per row as below
        } else {
#pragma unroll
            for (int i = 0; i < n; ++i) {
                float ax = __fmul_rn(A[i * n], x[0]);  // (intrinsics: no context-dependent contraction, see RCModel::rhs)
#pragma unroll
                for (int j = 1; j < n; ++j) ax = fmaf(A[i * n + j], x[j], ax);
                float bu = __fmul_rn(B[i * m], u[0]);
#pragma unroll
                for (int j = 1; j < m; ++j) bu = fmaf(B[i * m + j], u[j], bu);
                r[i] = __fadd_rn(ax, bu);  // fxx(x) + fxu(u)   base_block_state_space.py:61
            }
        }
    }


    // ---- hooks used by the rollout kernels ----
    static constexpr int SMEM_FLOATS = 0;  // no block-shared model data
    __device__ __forceinline__ static void block_init(float *, const Ptrs &, int, int) {}
    __device__ __forceinline__ void load(const Ptrs &P, int64_t i, const float *) { load(P, i); }
    // out[i] = (Ma x)_i + (Mb u)_i for row-major Ma[R][n], Mb[R][m]: two ROWS per FFMA2.  Each row goes through exactly
    // the operations of rhs()/the scalar loops (product of column 0, fused multiply-adds over the other columns, one
    // addition), so the states are bit-identical to the scalar loops of rhs().  One register layout cannot serve row pairs
    // and column pairs of the same matrix: the gradient kernel pairs rows throughout for ROWPAIR models (rhs + adj_rows) and
    // columns otherwise (scalar rhs + adj_cols).
    template <int R>
    __device__ __forceinline__ static void rows2(const float *Ma, const float *Mb, const float *x, const float *u, float *out) {
#pragma unroll
        for (int i = 0; i + 1 < R; i += 2) {
            float2 a = __fmul2_rn(make_float2(Ma[i * n], Ma[(i + 1) * n]), make_float2(x[0], x[0]));
#pragma unroll
            for (int j = 1; j < n; ++j) a = __ffma2_rn(make_float2(Ma[i * n + j], Ma[(i + 1) * n + j]), make_float2(x[j], x[j]), a);
            float2 b = __fmul2_rn(make_float2(Mb[i * m], Mb[(i + 1) * m]), make_float2(u[0], u[0]));
#pragma unroll
            for (int j = 1; j < m; ++j) b = __ffma2_rn(make_float2(Mb[i * m + j], Mb[(i + 1) * m + j]), make_float2(u[j], u[j]), b);
            const float2 r = __fadd2_rn(a, b);
            out[i] = r.x;
            out[i + 1] = r.y;
        }
        if constexpr (R % 2 == 1) {
            constexpr int i = R - 1;
            float a = __fmul_rn(Ma[i * n], x[0]);
#pragma unroll
            for (int j = 1; j < n; ++j) a = fmaf(Ma[i * n + j], x[j], a);
            float b = __fmul_rn(Mb[i * m], u[0]);
#pragma unroll
            for (int j = 1; j < m; ++j) b = fmaf(Mb[i * m + j], u[j], b);
            out[i] = __fadd_rn(a, b);
        }
    }

    // explicit Euler in residual form (simulator.py:40); `discrete` steps x_{k+1} = rhs instead
    __device__ __forceinline__ void step(float *x, const float *u, float dt, bool discrete) const {
        float r[n];
        rows2<n>(A, B, x, u, r);  // fxx(x) + fxu(u)   base_block_state_space.py:61
#pragma unroll
        for (int d = 0; d < n; ++d) x[d] = discrete ? r[d] : fmaf(dt, r[d], x[d]);
    }

    // fyx(x) + fyu(u)   base_block_state_space.py:69
    __device__ __forceinline__ void out(const float *x, const float *u, float *y) const { rows2<p>(C, D, x, u, y); }

    // G[0..w) += a * v[0..w): two entries per FFMA2 (the same IEEE operations as w scalar fused multiply-adds)
    template <int w>
    __device__ __forceinline__ static void axpy_row(float *G, const float a, const float *v) {
        const float2 aa = make_float2(a, a);
#pragma unroll
        for (int j = 0; j + 1 < w; j += 2) {
            const float2 r = __ffma2_rn(aa, make_float2(v[j], v[j + 1]), make_float2(G[j], G[j + 1]));
            G[j] = r.x;
            G[j + 1] = r.y;
        }
        if constexpr (w % 2 == 1) G[w - 1] = fmaf(a, v[w - 1], G[w - 1]);
    }

    __device__ __forceinline__ void accum(float *G, const float *s, const float *gy, const float *x,
                                          const float *u) const {
#pragma unroll
        for (int i = 0; i < n; ++i) {
            axpy_row<n>(G + i * n, s[i], x);
            axpy_row<m>(G + n * n + i * m, s[i], u);
        }
#pragma unroll
        for (int i = 0; i < p; ++i) {
            axpy_row<n>(G + n * n + n * m + i * n, gy[i], x);
            axpy_row<m>(G + n * n + n * m + p * n + i * m, gy[i], u);
        }
    }

    // out[0..w) = Ma^T s + Mb^T gy for row-major Ma[n][w], Mb[p][w]: column pairs per FFMA2 (the same IEEE operations, in the
    // same order, as the scalar loops: product of row 0, fused multiply-adds over the other rows, one addition)
    template <int w>
    __device__ __forceinline__ static void adj_cols(const float *Ma, const float *Mb, const float *s, const float *gy,
                                                    float *out) {
#pragma unroll
        for (int j = 0; j + 1 < w; j += 2) {
            float2 v = __fmul2_rn(make_float2(Ma[j], Ma[j + 1]), make_float2(s[0], s[0]));
#pragma unroll
            for (int i = 1; i < n; ++i) v = __ffma2_rn(make_float2(Ma[i * w + j], Ma[i * w + j + 1]), make_float2(s[i], s[i]), v);
            float2 q = __fmul2_rn(make_float2(Mb[j], Mb[j + 1]), make_float2(gy[0], gy[0]));
#pragma unroll
            for (int i = 1; i < p; ++i) q = __ffma2_rn(make_float2(Mb[i * w + j], Mb[i * w + j + 1]), make_float2(gy[i], gy[i]), q);
            const float2 r = __fadd2_rn(v, q);
            out[j] = r.x;
            out[j + 1] = r.y;
        }
        if constexpr (w % 2 == 1) {
            constexpr int j = w - 1;
            float v = __fmul_rn(Ma[j], s[0]);
#pragma unroll
            for (int i = 1; i < n; ++i) v = fmaf(Ma[i * w + j], s[i], v);
            float q = __fmul_rn(Mb[j], gy[0]);
#pragma unroll
            for (int i = 1; i < p; ++i) q = fmaf(Mb[i * w + j], gy[i], q);
            out[j] = __fadd_rn(v, q);
        }
    }

    // out[j] = (Ma^T s)_j + (Mb^T gy)_j with ROW pairs (ROWPAIR models): the sum over rows runs as two interleaved partial
    // sums (even rows, odd rows) per column, added at the end
    template <int w>
    __device__ __forceinline__ static void adj_rows(const float *Ma, const float *Mb, const float *s, const float *gy,
                                                    float *out) {
#pragma unroll
        for (int j = 0; j < w; ++j) {
            float2 v = __fmul2_rn(make_float2(Ma[j], Ma[w + j]), make_float2(s[0], s[1]));
#pragma unroll
            for (int i = 2; i < n; i += 2)
                v = __ffma2_rn(make_float2(Ma[i * w + j], Ma[(i + 1) * w + j]), make_float2(s[i], s[i + 1]), v);
#pragma unroll
            for (int i = 0; i < p; i += 2)
                v = __ffma2_rn(make_float2(Mb[i * w + j], Mb[(i + 1) * w + j]), make_float2(gy[i], gy[i + 1]), v);
            out[j] = __fadd_rn(v.x, v.y);
        }
    }

    __device__ __forceinline__ void adj_x(const float *s, const float *gy, float *ax) const {
        if constexpr (ROWPAIR) adj_rows<n>(A, C, s, gy, ax);
        else adj_cols<n>(A, C, s, gy, ax);
    }

    __device__ __forceinline__ void adj_u(const float *s, const float *gy, float *gu) const {
        if constexpr (ROWPAIR) adj_rows<m>(B, D, s, gy, gu);
        else adj_cols<m>(B, D, s, gy, gu);
    }

    __device__ __forceinline__ void finalize(const double *G, const Ptrs &, int64_t, double *g) const {
#pragma unroll
        for (int k = 0; k < NG; ++k) g[k] = G[k];
    }

    __device__ __forceinline__ static float *grad_addr(const GradPtrs &Gp, int64_t i, int j) {
        if (j < n * n) return Gp.gA ? Gp.gA + i * (n * n) + j : nullptr;
        j -= n * n;
        if (j < n * m) return Gp.gB ? Gp.gB + i * (n * m) + j : nullptr;
        j -= n * m;
        if (j < p * n) return Gp.gC ? Gp.gC + i * (p * n) + j : nullptr;
        j -= p * n;
        return Gp.gD ? Gp.gD + i * (p * m) + j : nullptr;
    }
};

}  // namespace dynax
