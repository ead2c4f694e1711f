/* CPU restatement (plain C + OpenMP) of dynax's batched rollout hot path.
 *
 * TEST INFRASTRUCTURE ONLY: linked/loaded solely by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs.  The product (dynax_b200/) never
 * touches it.
 *
 * Parity status: forward pinned by the reference's golden vectors through
 * tests/test_oracle_c.py (this file == oracle/dynax_oracle.py == goldens); gradient is
 * "parity unpinned by the reference" and anchored on torch-autograd fp64 instead.
 *
 * Follows (paths relative to /root/reference):
 *   dynax/models/RC.py:201-210,222-230,236-239,245-246   theta -> A,B,C,D
 *   dynax/core/base_block_state_space.py:59-61,68-71     rhs = fxx(x)+fxu(u); y = fyx(x)+fyu(u)
 *   dynax/simulators/simulator.py:36-43                  x += dt*rhs; emit (x_{k+1}, y_k)
 *   examples/3-inverse.py:96-112                         MSE + bound penalty, reverse-mode grad
 *
 * Layouts (row-major, the same as the CUDA library): theta[N,7], x0[N,3], u[T,N,5],
 * xsol[T,N,3], ysol[T,N,1], target[T,N], loss[N], grad[N,7].
 * Build with -ffp-contract=off so the op order is the reference's un-fused one.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define W 16 /* systems per SIMD block */

typedef struct {
    float A00, A02, A11, A12, A20, A21, A22, B00, B01, B02, B10, B13, B24;
} rc_coef_f;

static inline rc_coef_f rc_coef(const float *th) {
    const float Cai = th[0], Cwe = th[1], Cwi = th[2], Re = th[3], Ri = th[4], Rw = th[5], Rg = th[6];
    rc_coef_f c;
    c.A00 = -1.0f / Cai * (1.0f / Rg + 1.0f / Ri); /* RC.py:203 */
    c.A02 = 1.0f / (Cai * Ri);                       /* RC.py:204 */
    c.A11 = -1.0f / Cwe * (1.0f / Re + 1.0f / Rw); /* RC.py:205 */
    c.A12 = 1.0f / (Cwe * Rw);                       /* RC.py:206 */
    c.A20 = 1.0f / (Cwi * Ri);                       /* RC.py:207 */
    c.A21 = 1.0f / (Cwi * Rw);                       /* RC.py:208 */
    c.A22 = -1.0f / Cwi * (1.0f / Rw + 1.0f / Ri); /* RC.py:209 */
    c.B00 = 1.0f / (Cai * Rg);                       /* RC.py:224 */
    c.B01 = 1.0f / Cai;                              /* RC.py:225 */
    c.B02 = 1.0f / Cai;                              /* RC.py:226 */
    c.B10 = 1.0f / (Cwe * Re);                       /* RC.py:227 */
    c.B13 = 1.0f / Cwe;                              /* RC.py:228 */
    c.B24 = 1.0f / Cwi;                              /* RC.py:229 */
    return c;
}

void oracle_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* One explicit-Euler step for a SIMD block; zeros of A/B are skipped (adding 0*x is exact). */
#define RC_STEP(l)                                                                                  \
    do {                                                                                            \
        const float *uu = urow + (size_t)(l) * 5;                                                   \
        const float ax0 = c[l].A00 * x0_[l] + c[l].A02 * x2_[l];                                    \
        const float ax1 = c[l].A11 * x1_[l] + c[l].A12 * x2_[l];                                    \
        const float ax2 = c[l].A20 * x0_[l] + c[l].A21 * x1_[l] + c[l].A22 * x2_[l];                \
        const float bu0 = c[l].B00 * uu[0] + c[l].B01 * uu[1] + c[l].B02 * uu[2];                   \
        const float bu1 = c[l].B10 * uu[0] + c[l].B13 * uu[3];                                      \
        const float bu2 = c[l].B24 * uu[4];                                                         \
        r0 = ax0 + bu0; r1 = ax1 + bu1; r2 = ax2 + bu2;                                             \
    } while (0)

/* A1: forward rollout.  xsol/ysol may be NULL.  Returns 0. */
int oracle_rc_forward_f32(const float *theta, const float *x0, const float *u, float *xsol, float *ysol,
                          float *x_final, long N, long T, float dt, int shared_theta, int shared_u) {
#pragma omp parallel for schedule(static)
    for (long b = 0; b < (N + W - 1) / W; ++b) {
        const long i0 = b * W;
        const int w = (int)((N - i0) < W ? (N - i0) : W);
        rc_coef_f c[W];
        float x0_[W], x1_[W], x2_[W];
        for (int l = 0; l < w; ++l) {
            c[l] = rc_coef(theta + (shared_theta ? 0 : (size_t)(i0 + l) * 7));
            x0_[l] = x0[(size_t)(i0 + l) * 3 + 0];
            x1_[l] = x0[(size_t)(i0 + l) * 3 + 1];
            x2_[l] = x0[(size_t)(i0 + l) * 3 + 2];
        }
        for (long k = 0; k < T; ++k) {
            const float *urow = shared_u ? (u + (size_t)k * 5) : (u + ((size_t)k * N + i0) * 5);
            float *xr = xsol ? xsol + ((size_t)k * N + i0) * 3 : NULL;
            float *yr = ysol ? ysol + ((size_t)k * N + i0) : NULL;
#pragma omp simd
            for (int l = 0; l < w; ++l) {
                float r0, r1, r2;
                const float *urow_l = shared_u ? urow - (size_t)l * 5 : urow;
                { const float *urow = urow_l; RC_STEP(l); }
                if (yr) yr[l] = x0_[l];                 /* y_k = C x_k, C=[1,0,0], D=0 (pre-update) */
                x0_[l] = x0_[l] + dt * r0;              /* simulator.py:40 */
                x1_[l] = x1_[l] + dt * r1;
                x2_[l] = x2_[l] + dt * r2;
                if (xr) { xr[l * 3 + 0] = x0_[l]; xr[l * 3 + 1] = x1_[l]; xr[l * 3 + 2] = x2_[l]; }
            }
        }
        if (x_final)
            for (int l = 0; l < w; ++l) {
                x_final[(size_t)(i0 + l) * 3 + 0] = x0_[l];
                x_final[(size_t)(i0 + l) * 3 + 1] = x1_[l];
                x_final[(size_t)(i0 + l) * 3 + 2] = x2_[l];
            }
    }
    return 0;
}

/* A5+A6: loss[N], grad[N,7] by BPTT the way XLA transposes the scan: the forward sweep stores
 * the per-step residuals x_k, a reverse loop accumulates -- everything in fp32 like the
 * reference's default dtype.  lb/ub may be NULL (no penalty).  target[T,N] (or [T] if
 * shared_target). */
int oracle_rc_loss_grad_f32(const float *theta, const float *x0, const float *u, const float *target,
                            const float *lb, const float *ub, float *loss, float *grad, long N, long T,
                            float dt, int shared_theta, int shared_u, int shared_target) {
    int fail = 0;
#pragma omp parallel
    {
        float *xs = (float *)malloc((size_t)T * 3 * W * sizeof(float)); /* residuals x_k, SoA */
        if (!xs) {
#pragma omp atomic write
            fail = 1;
        }
#pragma omp barrier
        if (!fail) {
#pragma omp for schedule(static)
        for (long b = 0; b < (N + W - 1) / W; ++b) {
            const long i0 = b * W;
            const int w = (int)((N - i0) < W ? (N - i0) : W);
            rc_coef_f c[W];
            float x0_[W], x1_[W], x2_[W];
            for (int l = 0; l < w; ++l) {
                c[l] = rc_coef(theta + (shared_theta ? 0 : (size_t)(i0 + l) * 7));
                x0_[l] = x0[(size_t)(i0 + l) * 3 + 0];
                x1_[l] = x0[(size_t)(i0 + l) * 3 + 1];
                x2_[l] = x0[(size_t)(i0 + l) * 3 + 2];
            }
            float sq[W];
            memset(sq, 0, sizeof sq);
            for (long k = 0; k < T; ++k) {
                const float *urow0 = shared_u ? (u + (size_t)k * 5) : (u + ((size_t)k * N + i0) * 5);
                const float *tr = shared_target ? (target + k) : (target + (size_t)k * N + i0);
                float *xk = xs + (size_t)k * 3 * W;
#pragma omp simd
                for (int l = 0; l < w; ++l) {
                    float r0, r1, r2;
                    const float *urow = shared_u ? urow0 - (size_t)l * 5 : urow0;
                    RC_STEP(l);
                    xk[l] = x0_[l]; xk[W + l] = x1_[l]; xk[2 * W + l] = x2_[l];
                    const float res = x0_[l] - tr[shared_target ? 0 : l];
                    sq[l] += res * res;
                    x0_[l] = x0_[l] + dt * r0;
                    x1_[l] = x1_[l] + dt * r1;
                    x2_[l] = x2_[l] + dt * r2;
                }
            }
            /* reverse sweep */
            float a0[W], a1[W], a2[W];
            float gA00[W], gA02[W], gA11[W], gA12[W], gA20[W], gA21[W], gA22[W];
            float gB00[W], gB01[W], gB02[W], gB10[W], gB13[W], gB24[W];
            memset(a0, 0, sizeof a0); memset(a1, 0, sizeof a1); memset(a2, 0, sizeof a2);
            memset(gA00, 0, sizeof gA00); memset(gA02, 0, sizeof gA02); memset(gA11, 0, sizeof gA11);
            memset(gA12, 0, sizeof gA12); memset(gA20, 0, sizeof gA20); memset(gA21, 0, sizeof gA21);
            memset(gA22, 0, sizeof gA22); memset(gB00, 0, sizeof gB00); memset(gB01, 0, sizeof gB01);
            memset(gB02, 0, sizeof gB02); memset(gB10, 0, sizeof gB10); memset(gB13, 0, sizeof gB13);
            memset(gB24, 0, sizeof gB24);
            const float two_over_T = 2.0f / (float)T;
            for (long k = T - 1; k >= 0; --k) {
                const float *urow0 = shared_u ? (u + (size_t)k * 5) : (u + ((size_t)k * N + i0) * 5);
                const float *tr = shared_target ? (target + k) : (target + (size_t)k * N + i0);
                const float *xk = xs + (size_t)k * 3 * W;
#pragma omp simd
                for (int l = 0; l < w; ++l) {
                    const float *uu = (shared_u ? urow0 : urow0 + (size_t)l * 5);
                    const float X0 = xk[l], X1 = xk[W + l], X2 = xk[2 * W + l];
                    const float gy = two_over_T * (X0 - tr[shared_target ? 0 : l]);
                    const float s0 = dt * a0[l], s1 = dt * a1[l], s2 = dt * a2[l]; /* cotangent of rhs */
                    gA00[l] += s0 * X0; gA02[l] += s0 * X2;
                    gA11[l] += s1 * X1; gA12[l] += s1 * X2;
                    gA20[l] += s2 * X0; gA21[l] += s2 * X1; gA22[l] += s2 * X2;
                    gB00[l] += s0 * uu[0]; gB01[l] += s0 * uu[1]; gB02[l] += s0 * uu[2];
                    gB10[l] += s1 * uu[0]; gB13[l] += s1 * uu[3];
                    gB24[l] += s2 * uu[4];
                    const float n0 = a0[l] + (c[l].A00 * s0 + c[l].A20 * s2) + gy; /* a + A^T s + C^T gy */
                    const float n1 = a1[l] + (c[l].A11 * s1 + c[l].A21 * s2);
                    const float n2 = a2[l] + (c[l].A02 * s0 + c[l].A12 * s1 + c[l].A22 * s2);
                    a0[l] = n0; a1[l] = n1; a2[l] = n2;
                }
            }
            for (int l = 0; l < w; ++l) {
                const float *th = theta + (shared_theta ? 0 : (size_t)(i0 + l) * 7);
                const float iCai = 1.0f / th[0], iCwe = 1.0f / th[1], iCwi = 1.0f / th[2];
                const float iRe = 1.0f / th[3], iRi = 1.0f / th[4], iRw = 1.0f / th[5], iRg = 1.0f / th[6];
                float g[7];
                g[0] = -iCai * (gA00[l] * c[l].A00 + gA02[l] * c[l].A02 + gB00[l] * c[l].B00 + (gB01[l] + gB02[l]) * c[l].B01);
                g[1] = -iCwe * (gA11[l] * c[l].A11 + gA12[l] * c[l].A12 + gB10[l] * c[l].B10 + gB13[l] * c[l].B13);
                g[2] = -iCwi * (gA20[l] * c[l].A20 + gA21[l] * c[l].A21 + gA22[l] * c[l].A22 + gB24[l] * c[l].B24);
                g[3] = iCwe * iRe * iRe * (gA11[l] - gB10[l]);
                g[4] = iRi * iRi * (iCai * (gA00[l] - gA02[l]) + iCwi * (gA22[l] - gA20[l]));
                g[5] = iRw * iRw * (iCwe * (gA11[l] - gA12[l]) + iCwi * (gA22[l] - gA21[l]));
                g[6] = iCai * iRg * iRg * (gA00[l] - gB00[l]);
                float L = sq[l] / (float)T;
                if (lb && ub)
                    for (int j = 0; j < 7; ++j) { /* 3-inverse.py:103-107, normaliser = ub */
                        const float over = th[j] - ub[j], under = lb[j] - th[j];
                        if (over > 0) { L += over / ub[j]; g[j] += 1.0f / ub[j]; }
                        if (under > 0) { L += under / ub[j]; g[j] -= 1.0f / ub[j]; }
                    }
                loss[i0 + l] = L;
                for (int j = 0; j < 7; ++j) grad[(size_t)(i0 + l) * 7 + j] = g[j];
            }
        }
        }
        free(xs);
    }
    return fail ? -1 : 0;
}

/* fp64 "truth" for the parity samples of full-size runs (bench.py's parity block, tests/): loss, gradient and the
 * conditioning budgets of oracle/dynax_oracle.py::rc_output_sensitivity_budget in ONE forward sweep per system,
 * all arithmetic in double on the fp32 inputs.  The gradient is evaluated in forward mode (S = dx/dtheta beside the
 * state, F = d rhs/d theta in closed form from RC.py:203-209,224-229), i.e. independently of the reverse-mode
 * restatements above and in dynax_oracle.py (rc_loss_grad = hand adjoint + chain rule); tests/test_oracle_c.py pins
 * the two against each other.  loss/grad include the box penalty of examples/3-inverse.py:103-107 when lb/ub are given.
 * Budgets (may be NULL): loss_budget_i = (2/T) sum_k |r_k| rtol |y_k|,  grad_budget_ij = (2/T) sum_k |dy_k/dtheta_j| rtol |y_k|. */
int oracle_rc_loss_grad_budget_f64(const float *theta, const float *x0, const float *u, const float *target,
                                   const double *lb, const double *ub, double *loss, double *grad,
                                   double *loss_budget, double *grad_budget, long N, long T, double dt,
                                   int shared_theta, int shared_u, int shared_target, double traj_rtol) {
#pragma omp parallel for schedule(static)
    for (long i = 0; i < N; ++i) {
        const float *th = theta + (shared_theta ? 0 : (size_t)i * 7);
        const double Cai = th[0], Cwe = th[1], Cwi = th[2], Re = th[3], Ri = th[4], Rw = th[5], Rg = th[6];
        const double iCai = 1.0 / Cai, iCwe = 1.0 / Cwe, iCwi = 1.0 / Cwi, iRe = 1.0 / Re, iRi = 1.0 / Ri, iRw = 1.0 / Rw,
                     iRg = 1.0 / Rg;
        const double A00 = -1.0 / Cai * (1.0 / Rg + 1.0 / Ri), A02 = 1.0 / (Cai * Ri), A11 = -1.0 / Cwe * (1.0 / Re + 1.0 / Rw),
                     A12 = 1.0 / (Cwe * Rw), A20 = 1.0 / (Cwi * Ri), A21 = 1.0 / (Cwi * Rw),
                     A22 = -1.0 / Cwi * (1.0 / Rw + 1.0 / Ri);
        const double B00 = 1.0 / (Cai * Rg), B01 = 1.0 / Cai, B02 = 1.0 / Cai, B10 = 1.0 / (Cwe * Re), B13 = 1.0 / Cwe,
                     B24 = 1.0 / Cwi;
        double x[3] = {x0[(size_t)i * 3], x0[(size_t)i * 3 + 1], x0[(size_t)i * 3 + 2]};
        double S[3][7], g[7], gb[7], sq = 0.0, lbud = 0.0;
        memset(S, 0, sizeof S); memset(g, 0, sizeof g); memset(gb, 0, sizeof gb);
        for (long k = 0; k < T; ++k) {
            const float *uu = shared_u ? (u + (size_t)k * 5) : (u + ((size_t)k * N + i) * 5);
            const double tg = shared_target ? target[k] : target[(size_t)k * N + i];
            const double y = x[0], r = y - tg, w = traj_rtol * fabs(y);
            sq += r * r;
            lbud += fabs(r) * w;
            for (int j = 0; j < 7; ++j) { g[j] += r * S[0][j]; gb[j] += fabs(S[0][j]) * w; }
            const double r0 = (A00 * x[0] + A02 * x[2]) + (B00 * uu[0] + B01 * uu[1] + B02 * uu[2]);
            const double r1 = (A11 * x[1] + A12 * x[2]) + (B10 * uu[0] + B13 * uu[3]);
            const double r2 = (A20 * x[0] + A21 * x[1] + A22 * x[2]) + B24 * uu[4];
            double F[3][7];
            memset(F, 0, sizeof F);
            F[0][0] = -r0 * iCai; F[1][1] = -r1 * iCwe; F[2][2] = -r2 * iCwi;
            F[1][3] = -iCwe * iRe * iRe * (uu[0] - x[1]);
            F[0][4] = -iCai * iRi * iRi * (x[2] - x[0]); F[2][4] = -iCwi * iRi * iRi * (x[0] - x[2]);
            F[1][5] = -iCwe * iRw * iRw * (x[2] - x[1]); F[2][5] = -iCwi * iRw * iRw * (x[1] - x[2]);
            F[0][6] = -iCai * iRg * iRg * (uu[0] - x[0]);
            for (int j = 0; j < 7; ++j) {
                const double a0 = S[0][j], a1 = S[1][j], a2 = S[2][j];
                S[0][j] = a0 + dt * (A00 * a0 + A02 * a2 + F[0][j]);
                S[1][j] = a1 + dt * (A11 * a1 + A12 * a2 + F[1][j]);
                S[2][j] = a2 + dt * (A20 * a0 + A21 * a1 + A22 * a2 + F[2][j]);
            }
            x[0] += dt * r0; x[1] += dt * r1; x[2] += dt * r2;   /* simulator.py:40 */
        }
        double L = sq / (double)T;
        for (int j = 0; j < 7; ++j) g[j] *= 2.0 / (double)T;
        if (lb && ub)
            for (int j = 0; j < 7; ++j) { /* 3-inverse.py:103-107, normaliser = ub */
                const double over = (double)th[j] - ub[j], under = lb[j] - (double)th[j];
                if (over > 0) { L += over / ub[j]; g[j] += 1.0 / ub[j]; }
                if (under > 0) { L += under / ub[j]; g[j] -= 1.0 / ub[j]; }
            }
        loss[i] = L;
        for (int j = 0; j < 7; ++j) grad[(size_t)i * 7 + j] = g[j];
        if (loss_budget) loss_budget[i] = 2.0 / (double)T * lbud;
        if (grad_budget) for (int j = 0; j < 7; ++j) grad_budget[(size_t)i * 7 + j] = 2.0 / (double)T * gb[j];
    }
    return 0;
}
