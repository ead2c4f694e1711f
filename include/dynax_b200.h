/* dynax_b200 -- C ABI of the B200-native batched rollout engine (libdynax_b200.so).
 *
 * This is the drop-in boundary for dynax's hot path.  The reference has no FFI of its own;
 * the boundary it exposes is the Flax call convention
 *     xsol, ysol = DifferentiableSimulator(model, dt).apply(params, x0, inputs)
 *         (/root/reference/dynax/simulators/simulator.py:12-55)
 * over a BaseBlockSSM model (dynax/core/base_block_state_space.py:51-76), differentiated by
 * jax.value_and_grad (examples/3-inverse.py:96-112).  Each entry point below names the
 * reference code it replaces; INTEGRATION.md shows the jax.ffi / ctypes stubs that bind it.
 *
 * Conventions
 *  - All arithmetic is fp32 (JAX default dtype); gradient sums are folded in fp64 internally.
 *  - Unless a name ends in _host, every pointer is a DEVICE pointer owned by the caller, every
 *    call only enqueues work on `stream` (a cudaStream_t passed as void*) and returns at once.
 *    No hidden allocation: scratch comes from the caller (`ws`, size from dynax_workspace_bytes).
 *  - *_host entry points take HOST pointers, allocate device scratch themselves, tile the
 *    system axis, overlap the H2D/D2H copies with the kernels, and return when results are in
 *    the host buffers.
 *  - Layouts are row-major with the system axis in the middle, as north_star prescribes:
 *      theta[N,7]  x0[N,n]  u[T,N,m]  xsol[T,N,n]  ysol[T,N,p]  target[T,N]
 *    xsol[k] = x_{k+1} (post-update), ysol[k] = y(x_k,u_k) (pre-update): simulator.py:38-43.
 *  - Return value: DYNAX_OK or a negative DYNAX_ERR_*; dynax_last_error() gives a thread-local
 *    message.  Nothing throws.  There is no CPU fallback: without an sm_100 device every compute
 *    entry point fails with DYNAX_ERR_NO_DEVICE.
 */
#ifndef DYNAX_B200_H_
#define DYNAX_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DYNAX_ABI_VERSION 1

#define DYNAX_OK 0
#define DYNAX_ERR_INVALID_ARGUMENT (-1)
#define DYNAX_ERR_UNSUPPORTED_DIMS (-2)
#define DYNAX_ERR_WORKSPACE (-3)
#define DYNAX_ERR_CUDA (-4)
#define DYNAX_ERR_NO_DEVICE (-5)

/* flags (bit-or) */
#define DYNAX_SHARED_THETA 1u   /* model parameters are one set [1,...] broadcast to all N systems;
                                   parameter gradients come back SUMMED over the N systems          */
#define DYNAX_SHARED_U 2u       /* u is [T,1,m]: one input series for all systems                  */
#define DYNAX_SHARED_TARGET 4u  /* target is [T]: one measured series for all systems              */
#define DYNAX_DISCRETE 8u       /* x_{k+1} = rhs (stepping a Discrete* model directly,
                                   tests/test_forward_simulation.py:9-12,33-36) instead of the
                                   simulator's x += dt*rhs (simulator.py:40)                        */

#define DYNAX_SOLVER_EULER 16u  /* node_rk4_forward: explicit Euler instead of classical RK4            */

/* ops for dynax_workspace_bytes */
#define DYNAX_OP_RC_FORWARD 0
#define DYNAX_OP_RC_LOSS_GRAD 1
#define DYNAX_OP_RC_VJP 2
#define DYNAX_OP_LTI_FORWARD 3
#define DYNAX_OP_LTI_VJP 4

int dynax_version(void);
const char *dynax_last_error(void);
/* 0 if the current CUDA device is compute capability 10.x, else DYNAX_ERR_NO_DEVICE. */
int dynax_device_check(void);
/* Debug hook for the variant-coverage tests (tests/test_gpu_variants.py): the library registers every rollout-kernel
 * launch site it instantiates (template arguments included) when it is loaded and counts the launches of each, so a
 * test can prove that the case it runs selected the kernel variant it means to check.  No reference counterpart. */
int dynax_debug_kernel_count(void);
const char *dynax_debug_kernel_tag(int i);        /* "" if i is out of range */
long long dynax_debug_kernel_launches(int i);     /* launches of site i so far in this process; -1 if out of range */
const char *dynax_debug_last_kernel(void);        /* tag of the calling thread's last launch ("" if none) */
/* Bytes of device scratch `ws` the op needs (0 for forward ops). */
size_t dynax_workspace_bytes(int op, int64_t N, int64_t T, int n, int m, int p, uint32_t flags);
/* 1 if dense-LTI kernels are instantiated for (n,m,p), else 0. */
int dynax_lti_supported(int n, int m, int p);

/* ---- 4R3C RC zone model: Continuous4R3C / Discrete4R3C (dynax/models/RC.py:21-249) rolled out
 *      by DifferentiableSimulator.__call__ (simulators/simulator.py:22-55).  n=3, m=5, p=1.
 *      theta = [Cai,Cwe,Cwi,Re,Ri,Rw,Rg] (RC.py:176-182); A(theta),B(theta) are built in-kernel
 *      (RC.py:201-210,222-230), C=[1,0,0], D=0 (RC.py:236-246).
 *      xsol, ysol, x_final may each be NULL (not emitted). ---------------------------------- */
int dynax_rc_forward_f32(const float *theta, const float *x0, const float *u,
                         float *xsol, float *ysol, float *x_final,
                         int64_t N, int64_t T, float dt, uint32_t flags, void *stream);

/* Time-parallel variant of dynax_rc_forward_f32 for launches with too few systems to fill the GPU
 * (long-horizon LTI, north_star "chunked associative-scan variant"): the horizon is cut into chunks
 * of `chunk_len` steps (0 = choose automatically) that are rolled in parallel from zero state, stitched
 * with Psi = (I+dt A)^L - I, and rolled again from their true boundary states.
 * Same outputs as dynax_rc_forward_f32 within the trajectory tolerance (not bit-identical).  Euler
 * residual form only (no DYNAX_DISCRETE).  ws: dynax_rc_forward_chunked_workspace(N,T,chunk_len) bytes. */
size_t dynax_rc_forward_chunked_workspace(int64_t N, int64_t T, int64_t chunk_len);
int dynax_rc_forward_chunked_f32(const float *theta, const float *x0, const float *u,
                                 float *xsol, float *ysol, float *x_final,
                                 int64_t N, int64_t T, float dt, uint32_t flags, int64_t chunk_len,
                                 void *ws, size_t ws_bytes, void *stream);

/* Time-parallel variant of dynax_rc_loss_grad_f32 (same outputs within the gradient tolerance) for
 * launches with few systems and a long horizon (one building over a year: N=1, T=8760).  Forward chunk
 * boundaries as above; then every chunk runs the forward-mode (tangent) kernel from its true boundary state with
 * zero sensitivities, and a stitch over the chunks assembles loss, g_theta and g_x0 (u is read twice in total).
 * With DYNAX_GRAD_ALGO=0 in the environment: the adjoint-based scheme (zero-terminal adjoint pass, backward
 * stitch with Psi^T, true-terminal pass; u read 5 times).  Use it only when N cannot fill the GPU. */
size_t dynax_rc_loss_grad_chunked_workspace(int64_t N, int64_t T, int64_t chunk_len);
int dynax_rc_loss_grad_chunked_f32(const float *theta, const float *x0, const float *u, const float *target,
                                   const float *lb, const float *ub,
                                   float *loss, float *g_theta, float *g_x0,
                                   int64_t N, int64_t T, float dt, uint32_t flags, int64_t chunk_len,
                                   void *ws, size_t ws_bytes, void *stream);

/* Fused loss + reverse-mode gradient of examples/3-inverse.py:96-112 for N candidates:
 *   loss_i = mean_k (ysol_i[k]-target_i[k])^2 + sum_j relu(theta_ij-ub_j)/ub_j + relu(lb_j-theta_ij)/ub_j
 *   g_theta = d loss / d theta   (== jax.value_and_grad through the scan).  The gradient is evaluated in forward
 *   mode (3x7 sensitivities beside the state, +3 columns for g_x0; one sweep, 24 B per system-step:
 *   csrc/rc_fwdsens.cuh) for every combination of shared / per-system theta, u and target; DYNAX_DISCRETE -- and
 *   every call when the environment holds DYNAX_GRAD_ALGO=0 -- runs the checkpointed discrete adjoint
 *   (csrc/rollout_adj.cuh).  The two agree within the gradient tolerance, not bit for bit.  The workspace is
 *   required either way.
 * lb/ub: [7] device arrays or NULL (no penalty).  loss:[N], g_theta:[N,7]; with DYNAX_SHARED_THETA
 * loss:[1] (mean over systems is NOT taken: it is the SUM over systems of the MSE terms plus one
 * penalty) and g_theta:[7].  g_theta may be NULL (loss only), g_x0:[N,3] or NULL. */
int dynax_rc_loss_grad_f32(const float *theta, const float *x0, const float *u, const float *target,
                           const float *lb, const float *ub,
                           float *loss, float *g_theta, float *g_x0,
                           int64_t N, int64_t T, float dt, uint32_t flags,
                           void *ws, size_t ws_bytes, void *stream);

/* Shared input series + one per-system channel: u_k(system) = u_shared[k] with column `ctrl_col`
 * replaced by ctrl[k, system].  This is how the RC env assembles its inputs from the shared
 * disturbance table and the per-env control (dynax/wrapper/envs/rc.py:128-131); the streamed input
 * drops from 20 to 4 bytes per system-step.  u_shared:[T,5], ctrl:[T,N].  DYNAX_SHARED_U is implied. */
int dynax_rc_forward_ctrl_f32(const float *theta, const float *x0, const float *u_shared, const float *ctrl,
                              int ctrl_col, float *xsol, float *ysol, float *x_final,
                              int64_t N, int64_t T, float dt, uint32_t flags, void *stream);
int dynax_rc_loss_grad_ctrl_f32(const float *theta, const float *x0, const float *u_shared, const float *ctrl,
                                int ctrl_col, const float *target, const float *lb, const float *ub,
                                float *loss, float *g_theta, float *g_x0,
                                int64_t N, int64_t T, float dt, uint32_t flags,
                                void *ws, size_t ws_bytes, void *stream);

/* Tabular time-table lookup (dynax/agents/tabular.py:33-44 over dynax/utils/interpolate.py:26-68):
 * out[i,:] = interpolation of the rows xs[n_rows,n_cols] sampled at times ts[n_rows] evaluated at at[i].
 * mode 1 = 'linear' (jnp.interp to a fractional row index, then order-1 map_coordinates),
 * mode 0 = 'constant' (order-0 map_coordinates = NEAREST row, halves round away from zero; it is not a
 * zero-order hold).  Queries outside [ts[0], ts[-1]] clamp to the end rows.  out:[n_at,n_cols]. */
int dynax_tabular_interp_f32(const float *ts, const float *xs, int64_t n_rows, int n_cols,
                             const float *at, int64_t n_at, int mode, float *out, void *stream);

/* General-cotangent VJP of dynax_rc_forward_f32 (what jax.vjp/jax.grad of simulator.apply needs
 * for an arbitrary downstream loss).  g_xsol:[T,N,3] or NULL, g_ysol:[T,N,1] or NULL (at least
 * one non-NULL).  Outputs (each may be NULL): g_theta:[N,7] ([7] if SHARED_THETA), g_x0:[N,3],
 * g_u:[T,N,5]; with DYNAX_SHARED_U g_u:[T,5], the cotangent of the ONE shared series = the sum over systems, reduced
 * inside the kernel (per-CTA sums in the workspace, then a fixed-order reduction: bit-reproducible). */
int dynax_rc_vjp_f32(const float *theta, const float *x0, const float *u,
                     const float *g_xsol, const float *g_ysol,
                     float *g_theta, float *g_x0, float *g_u,
                     int64_t N, int64_t T, float dt, uint32_t flags,
                     void *ws, size_t ws_bytes, void *stream);

/* Time-parallel variant of dynax_rc_vjp_f32 (same outputs within the gradient tolerance) for launches with few
 * systems and a long horizon: forward chunk boundaries, every chunk's adjoint for a zero terminal adjoint, a backward
 * stitch with Psi^T, every chunk's adjoint again from its true terminal adjoint.  Per-system u, Euler residual form;
 * flags: DYNAX_SHARED_THETA only.  ws: dynax_rc_vjp_chunked_workspace(N,T,chunk_len) bytes; chunk_len 0 = automatic. */
size_t dynax_rc_vjp_chunked_workspace(int64_t N, int64_t T, int64_t chunk_len);
int dynax_rc_vjp_chunked_f32(const float *theta, const float *x0, const float *u,
                             const float *g_xsol, const float *g_ysol,
                             float *g_theta, float *g_x0, float *g_u,
                             int64_t N, int64_t T, float dt, uint32_t flags, int64_t chunk_len,
                             void *ws, size_t ws_bytes, void *stream);

/* ---- Dense linear block SSM: DiscreteLinearSSM / ContinuousLinearSSM
 *      (dynax/core/discrete_block_state_space.py:5-54, continuous_block_state_space.py:5-54).
 *      A:[N|1,n,n] B:[N|1,n,m] C:[N|1,p,n] D:[N|1,p,m] are the MATRICES (row-major), i.e. the
 *      transposes of the Flax nn.Dense kernels.  (n,m,p) must satisfy dynax_lti_supported. ---- */
int dynax_lti_forward_f32(const float *A, const float *B, const float *C, const float *D,
                          const float *x0, const float *u,
                          float *xsol, float *ysol, float *x_final,
                          int64_t N, int64_t T, int n, int m, int p, float dt, uint32_t flags, void *stream);

int dynax_lti_vjp_f32(const float *A, const float *B, const float *C, const float *D,
                      const float *x0, const float *u,
                      const float *g_xsol, const float *g_ysol,
                      float *g_A, float *g_B, float *g_C, float *g_D, float *g_x0, float *g_u,
                      int64_t N, int64_t T, int n, int m, int p, float dt, uint32_t flags,
                      void *ws, size_t ws_bytes, void *stream);

/* Time-parallel variant of dynax_lti_forward_f32 (same scheme as dynax_rc_forward_chunked_f32: Psi = (I+dt A)^L - I,
 * zero-state pass, stitch, second pass) for launches with few systems and a long horizon.  Euler residual form only
 * (no DYNAX_DISCRETE).  ws: dynax_lti_forward_chunked_workspace(N,T,n,chunk_len) bytes; chunk_len 0 = automatic. */
size_t dynax_lti_forward_chunked_workspace(int64_t N, int64_t T, int n, int64_t chunk_len);
int dynax_lti_forward_chunked_f32(const float *A, const float *B, const float *C, const float *D,
                                  const float *x0, const float *u,
                                  float *xsol, float *ysol, float *x_final,
                                  int64_t N, int64_t T, int n, int m, int p, float dt, uint32_t flags,
                                  int64_t chunk_len, void *ws, size_t ws_bytes, void *stream);

/* Time-parallel variant of dynax_lti_vjp_f32 (same scheme as dynax_rc_vjp_chunked_f32) for launches with few systems
 * and a long horizon -- fitting A,B,C,D to one measured trajectory.  Per-system u, Euler residual form; flags:
 * DYNAX_SHARED_THETA only.  ws: dynax_lti_vjp_chunked_workspace(N,T,n,m,p,chunk_len) bytes. */
size_t dynax_lti_vjp_chunked_workspace(int64_t N, int64_t T, int n, int m, int p, int64_t chunk_len);
int dynax_lti_vjp_chunked_f32(const float *A, const float *B, const float *C, const float *D,
                              const float *x0, const float *u, const float *g_xsol, const float *g_ysol,
                              float *g_A, float *g_B, float *g_C, float *g_D, float *g_x0, float *g_u,
                              int64_t N, int64_t T, int n, int m, int p, float dt, uint32_t flags,
                              int64_t chunk_len, void *ws, size_t ws_bytes, void *stream);

/* ---- General fx/fy block SSM (dynax/core/base_block_state_space.py:62-63,72-73) with MLP dynamics,
 *      BASELINE.json config 5.  The reference has only the dispatch for this form (its example
 *      examples/4-neural-ode.py does not run), so the model is the restated
 *          rhs = W3 tanh(W2 tanh(W1 [x;u] + b1) + b2) + b3,   y = x[0:p]
 *      stepped with classical RK4 (u held constant over a step) or Euler (DYNAX_SOLVER_EULER).
 *      Weights are row-major [out,in] matrices shared by all systems: W1:[hidden,n+m] b1:[hidden]
 *      W2:[hidden,hidden] b2:[hidden] W3:[n,hidden] b3:[n].  Instantiated for (n,m,p)=(3,5,1) and hidden = 32, 64
 *      or 128 (forward and gradient); anything else: DYNAX_ERR_UNSUPPORTED_DIMS.
 *      u: [T,N,5], or ONE series [T,5] for all trajectories with DYNAX_SHARED_U.
 *      Output convention as dynax_rc_forward_f32.  PARITY UNPINNED BY THE REFERENCE. ---- */
int dynax_node_rk4_forward_f32(const float *W1, const float *b1, const float *W2, const float *b2,
                               const float *W3, const float *b3, const float *x0, const float *u,
                               float *xsol, float *ysol, float *x_final,
                               int64_t N, int64_t T, int n, int m, int p, int hidden, float dt,
                               uint32_t flags, void *stream);
/* Same rollout, additionally emitting stage_x:[T,N,9] = the state parts of the RK4 stage points z2, z3, z4 of
 * every step (36 B per trajectory-step).  Passing them to dynax_node_rk4_vjp_stages_f32 saves the backward pass the
 * three forward evaluations per step it would spend recomputing them.  tcgen05 kernel only (the product path;
 * DYNAX_NODE_IMPL=0 selects the CUDA-core cross-check kernel, hidden = 64, which does not emit them). */
int dynax_node_rk4_forward_stages_f32(const float *W1, const float *b1, const float *W2, const float *b2,
                                      const float *W3, const float *b3, const float *x0, const float *u,
                                      float *xsol, float *ysol, float *x_final, float *stage_x,
                                      int64_t N, int64_t T, int n, int m, int p, int hidden, float dt,
                                      uint32_t flags, void *stream);

/* Reverse-mode gradient of dynax_node_rk4_forward_f32 (exact discrete adjoint of RK4 / Euler): cotangents
 * g_xsol:[T,N,3] and/or g_ysol:[T,N,1] in; weight gradients (summed over systems, each may be NULL),
 * g_x0:[N,3], g_u:[T,N,5] out.  xsol is the forward trajectory (x_{k+1} at row k).  hidden = 64: tcgen05 kernel
 * (node_vjp_tc.cuh), CUDA-core kernel with DYNAX_NODE_VJP_IMPL=0; hidden = 32: CUDA-core kernel; hidden = 128: the wide
 * CUDA-core kernel (node_vjp_wide.cuh: 32-trajectory CTAs, eight threads per trajectory).  With DYNAX_SHARED_U
 * u is ONE series [T,5] and g_u:[T,5] its gradient summed over the trajectories (fp32 atomics: one per warp, step and
 * channel, so the last bits depend on the order).
 * ws: dynax_node_rk4_vjp_workspace() bytes.  PARITY UNPINNED BY THE REFERENCE. */
size_t dynax_node_rk4_vjp_workspace(void);
int dynax_node_rk4_vjp_f32(const float *W1, const float *b1, const float *W2, const float *b2,
                           const float *W3, const float *b3, const float *x0, const float *u, const float *xsol,
                           const float *g_xsol, const float *g_ysol,
                           float *gW1, float *gb1, float *gW2, float *gb2, float *gW3, float *gb3,
                           float *g_x0, float *g_u,
                           int64_t N, int64_t T, int n, int m, int p, int hidden, float dt, uint32_t flags,
                           void *ws, size_t ws_bytes, void *stream);
/* Same gradient with the stage points saved by dynax_node_rk4_forward_stages_f32 (stage_x:[T,N,9], or NULL to
 * recompute them as dynax_node_rk4_vjp_f32 does). */
int dynax_node_rk4_vjp_stages_f32(const float *W1, const float *b1, const float *W2, const float *b2,
                                  const float *W3, const float *b3, const float *x0, const float *u, const float *xsol,
                                  const float *stage_x, const float *g_xsol, const float *g_ysol,
                                  float *gW1, float *gb1, float *gW2, float *gb2, float *gW3, float *gb3,
                                  float *g_x0, float *g_u,
                                  int64_t N, int64_t T, int n, int m, int p, int hidden, float dt, uint32_t flags,
                                  void *ws, size_t ws_bytes, void *stream);

/* ---- Sampling-based MPC (BASELINE.json config 4): cost of N candidate control sequences u[H,N] through
 *      ONE shared 4R3C model.  Inputs per step are assembled as the RC env does,
 *      [d_k0, d_k1, u_k, d_k2, d_k3] (dynax/wrapper/envs/rc.py:131), rolled out by simulator.py:36-43, and
 *      cost_i = sum_k  w_energy*price[h]*(-u_k/COP)*dt/3600 + w_comfort*(relu(y_k-t_high[h]) + relu(t_low[h]-y_k))
 *      i.e. the negated env reward (rc.py:227-252): y_k from the pre-update state, hour h from the
 *      post-update time start_time+(k+1)*dt.  theta:[7] x0:[3] d:[H,4] u:[H,N] price/t_low/t_high:[24]
 *      cost:[N]; argmin (int64) / min_cost (float) may be NULL, else ws must hold >= 16 bytes. ---- */
int dynax_rc_mpc_cost_f32(const float *theta, const float *x0, const float *d, const float *u,
                          const float *price, const float *t_low, const float *t_high,
                          float cop, float w_energy, float w_comfort, float start_time,
                          float *cost, long long *argmin, float *min_cost,
                          int64_t N, int64_t H, float dt, void *ws, size_t ws_bytes, void *stream);

/* ---- Batched RC-env step: K consecutive RC.step calls
 *      (dynax/wrapper/envs/rc.py:103-158) for N independent environments in one launch: action ->
 *      control (rc.py:166), linear Tabular lookup of the 4 disturbances at each env's own time (rc.py:128),
 *      one Euler step (rc.py:131-132), observation (rc.py:200-222), reward (rc.py:227-252) and
 *      termination flag (rc.py:255-264).  theta:[7]|[N,7] x:[N,3] t:[N] action:[K,N] int32
 *      ts:[R] dist:[R,4] price/t_low/t_high:[24];  outputs x_next:[N,3] t_next:[N] obs:[K,N,5]
 *      reward:[K,N] terminated:[K,N] (1.0/0.0), cost/dTz:[K,N] or NULL. ---- */
int dynax_rc_env_step_f32(const float *theta, const float *x, const float *t, const int32_t *action,
                          const float *ts, const float *dist, int64_t R,
                          const float *price, const float *t_low, const float *t_high,
                          float cop, float w_energy, float w_comfort, float u_low, float u_high, int num_actions,
                          float *x_next, float *t_next, float *obs, float *reward, float *terminated,
                          float *cost, float *dTz,
                          int64_t N, int64_t K, float dt, uint32_t flags, void *stream);

/* ---- On-device calibration loop pieces.
 *      dynax_lamb_update_f32: the optax.lamb(lr) update that examples/3-inverse.py:113-122 applies after
 *      value_and_grad, element-wise over n = N*7 scalar parameters (every Flax leaf there is a scalar, so
 *      the LAMB trust ratio is per element).  m, v: Adam moments (zero-initialised by the caller),
 *      step: 1-based epoch.  PARITY UNPINNED BY THE REFERENCE (optax is an absent, unpinned dependency).
 *      dynax_resample_mean_f32: data.groupby(index // dt).mean() of examples/1-forward.py:26-27:
 *      out[r,:] = mean(in[r*factor : (r+1)*factor, :]), out rows = ceil(rows/factor). ---- */
int dynax_lamb_update_f32(float *theta, const float *grad, float *m, float *v, int64_t n, int64_t step,
                          float lr, float b1, float b2, float eps, void *stream);
/* Same update with the epoch number in DEVICE memory (*step_counter = epochs done so far; incremented by
 * the call), so a CUDA graph capturing {dynax_rc_loss_grad_f32, dynax_lamb_update_dev_f32} replays as is. */
int dynax_lamb_update_dev_f32(float *theta, const float *grad, float *m, float *v, int64_t n, int64_t *step_counter,
                              float lr, float b1, float b2, float eps, void *stream);
int dynax_resample_mean_f32(const float *in, float *out, int64_t rows, int cols, int64_t factor, void *stream);

/* Non-finite guard: explicit Euler is only conditionally stable, and the parameter box of examples/3-inverse.py:66-85
 * contains corners where x + dt*rhs blows up (dt*|lambda| > 2).  For any per-system output a[rows,width] (loss[N,1],
 * x_final[N,3], g_theta[N,7] ...): flags[i] = 1 if row i holds a NaN/Inf (int32[rows] or NULL), *count = number of such
 * rows (device int or NULL).  No reference counterpart (there the caller sees an inf loss). */
int dynax_nonfinite_rows_f32(const float *a, int64_t rows, int width, int32_t *flags, int *count, void *stream);

/* ---- Host-buffer entry points (the call a reference user makes with NumPy/JAX-CPU arrays):
 *      pageable or pinned HOST pointers in, results in HOST buffers on return.  The system axis is
 *      processed in tiles of `tile_systems` (0 = library default) with copies overlapped. ------ */
int dynax_rc_forward_host_f32(const float *theta, const float *x0, const float *u,
                              float *xsol, float *ysol, float *x_final,
                              int64_t N, int64_t T, float dt, uint32_t flags, int64_t tile_systems);

int dynax_rc_loss_grad_host_f32(const float *theta, const float *x0, const float *u, const float *target,
                                const float *lb, const float *ub,
                                float *loss, float *g_theta,
                                int64_t N, int64_t T, float dt, uint32_t flags, int64_t tile_systems);

/* Host-buffer flavour of dynax_rc_loss_grad_ctrl_f32: only ctrl[T,N] (+ target) cross PCIe per system. */
int dynax_rc_loss_grad_ctrl_host_f32(const float *theta, const float *x0, const float *u_shared, const float *ctrl,
                                     int ctrl_col, const float *target, const float *lb, const float *ub,
                                     float *loss, float *g_theta,
                                     int64_t N, int64_t T, float dt, uint32_t flags, int64_t tile_systems);

/* Resident data set: upload u[T,N,5] (or [T,5] with DYNAX_SHARED_U) and target[T,N] (or [T] with
 * DYNAX_SHARED_TARGET) ONCE, then evaluate loss and gradient for changing parameters -- the loop of
 * examples/3-inverse.py:133-138 (5000 epochs over the same measured series).  Per call only theta[N,7] ([7] with
 * DYNAX_SHARED_THETA) and x0[N,3] go in and loss[N|1], g_theta[N|1,7] come out (HOST pointers); returns when the
 * results are in the host buffers.  The handle belongs to the device that was current at creation. */
int dynax_rc_dataset_create_f32(const float *u, const float *target, int64_t N, int64_t T, uint32_t flags, void **handle);
int dynax_rc_dataset_loss_grad_f32(void *handle, const float *theta, const float *x0, const float *lb, const float *ub,
                                   float dt, uint32_t flags, float *loss, float *g_theta);
int dynax_rc_dataset_destroy(void *handle);

/* Page-lock / unlock a long-lived HOST buffer (cudaHostRegister): pageable memory -- what NumPy allocates -- is staged
 * by the driver at a fraction of the link rate; a registered buffer is copied at the pinned rate by the *_host calls. */
int dynax_host_register(const void *ptr, size_t bytes);
int dynax_host_unregister(const void *ptr);

/* Frees the device scratch the *_host entry points cache between calls. */
int dynax_host_release(void);

#ifdef __cplusplus
}
#endif
#endif /* DYNAX_B200_H_ */
