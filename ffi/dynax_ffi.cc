// XLA FFI handlers over the C-ABI of libdynax_b200.so: the reference-side binding north_star asks for
// (jax.ffi custom calls + jax.custom_vjp).  NOT built in this repository's image --
// JAX (and with it xla/ffi/api/ffi.h) is not installable there -- so this file is shipped as source; on a
// machine with JAX >= 0.4.31 and a B200:
//
//   g++ -std=c++17 -shared -fPIC -I"$(python -c 'import jax.ffi; print(jax.ffi.include_dir())')" \
//       -I include -I /usr/local/cuda/include ffi/dynax_ffi.cc -L dynax_b200 -ldynax_b200 -o ffi/dynax_ffi.so
//
// Every handler forwards device pointers and the XLA stream to one C-ABI entry point and turns a negative
// return code into an ffi::Error carrying dynax_last_error().  Scratch memory is an extra XLA-allocated
// result buffer (the C-ABI never allocates).  tests/test_ffi_shim.py keeps the calls below in step with
// include/dynax_b200.h (names and argument counts) without needing JAX.
#include <cstdint>

#include <cuda_runtime_api.h>

#include "dynax_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;
using F32 = ffi::Buffer<ffi::F32>;
using RF32 = ffi::ResultBuffer<ffi::F32>;
using RU8 = ffi::ResultBuffer<ffi::U8>;

static ffi::Error status(int rc) {
  return rc == DYNAX_OK ? ffi::Error::Success() : ffi::Error::Internal(dynax_last_error());
}

// u[T,N,5] or u[T,5] (shared series): the flag is derived from the rank, as the Python facade does
static uint32_t shared_u_flag(const F32 &u) { return u.dimensions().size() == 2 ? DYNAX_SHARED_U : 0u; }

// xsol[T,N,3], ysol[T,N,1] = DifferentiableSimulator(Continuous4R3C, dt).apply(theta, x0, u)
// replaces /root/reference/dynax/simulators/simulator.py:22-55 over dynax/models/RC.py:137-249
static ffi::Error RcForward(cudaStream_t stream, F32 theta, F32 x0, F32 u, float dt, int32_t flags, RF32 xsol,
                            RF32 ysol) {
  const int64_t T = u.dimensions()[0], N = x0.dimensions()[0];
  return status(dynax_rc_forward_f32(theta.typed_data(), x0.typed_data(), u.typed_data(), xsol->typed_data(),
                                     ysol->typed_data(), nullptr, N, T, dt, (uint32_t)flags | shared_u_flag(u),
                                     stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(dynax_rc_forward, RcForward,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<F32>()   // theta [N,7]
                                  .Arg<F32>()   // x0    [N,3]
                                  .Arg<F32>()   // u     [T,N,5] | [T,5]
                                  .Attr<float>("dt")
                                  .Attr<int32_t>("flags")
                                  .Ret<F32>()   // xsol  [T,N,3]
                                  .Ret<F32>()); // ysol  [T,N,1]

// (g_theta, g_x0, g_u) = vjp of the call above for cotangents (g_xsol, g_ysol): what jax.grad needs
// (the transpose XLA would derive through nn.scan, examples/3-inverse.py:112)
static ffi::Error RcVjp(cudaStream_t stream, F32 theta, F32 x0, F32 u, F32 g_xsol, F32 g_ysol, float dt, int32_t flags,
                        RF32 g_theta, RF32 g_x0, RF32 g_u, RU8 ws) {
  const int64_t T = u.dimensions()[0], N = x0.dimensions()[0];
  return status(dynax_rc_vjp_f32(theta.typed_data(), x0.typed_data(), u.typed_data(), g_xsol.typed_data(),
                                 g_ysol.typed_data(), g_theta->typed_data(), g_x0->typed_data(), g_u->typed_data(), N,
                                 T, dt, (uint32_t)flags, ws->typed_data(), ws->size_bytes(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(dynax_rc_vjp, RcVjp,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<F32>()   // theta
                                  .Arg<F32>()   // x0
                                  .Arg<F32>()   // u      [T,N,5]
                                  .Arg<F32>()   // g_xsol [T,N,3]
                                  .Arg<F32>()   // g_ysol [T,N,1]
                                  .Attr<float>("dt")
                                  .Attr<int32_t>("flags")
                                  .Ret<F32>()   // g_theta [N,7]
                                  .Ret<F32>()   // g_x0    [N,3]
                                  .Ret<F32>()   // g_u     [T,N,5]
                                  .Ret<ffi::Buffer<ffi::U8>>());  // scratch: dynax_workspace_bytes(DYNAX_OP_RC_VJP, ...)

// loss[N], g_theta[N,7] = jax.value_and_grad(mse_loss) of examples/3-inverse.py:96-112, fused
// lb/ub: [7] parameter box (always passed; use +-inf to disable a bound)
static ffi::Error RcLossGrad(cudaStream_t stream, F32 theta, F32 x0, F32 u, F32 target, F32 lb, F32 ub, float dt,
                             int32_t flags, RF32 loss, RF32 g_theta, RU8 ws) {
  const int64_t T = u.dimensions()[0], N = x0.dimensions()[0];
  const uint32_t f = (uint32_t)flags | shared_u_flag(u) | (target.dimensions().size() == 1 ? DYNAX_SHARED_TARGET : 0u) |
                     (theta.dimensions().size() == 1 ? DYNAX_SHARED_THETA : 0u);
  return status(dynax_rc_loss_grad_f32(theta.typed_data(), x0.typed_data(), u.typed_data(), target.typed_data(),
                                       lb.typed_data(), ub.typed_data(), loss->typed_data(), g_theta->typed_data(),
                                       nullptr, N, T, dt, f, ws->typed_data(), ws->size_bytes(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(dynax_rc_loss_grad, RcLossGrad,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<F32>()   // theta  [N,7] | [7]
                                  .Arg<F32>()   // x0     [N,3]
                                  .Arg<F32>()   // u      [T,N,5] | [T,5]
                                  .Arg<F32>()   // target [T,N] | [T]
                                  .Arg<F32>()   // lb [7]
                                  .Arg<F32>()   // ub [7]
                                  .Attr<float>("dt")
                                  .Attr<int32_t>("flags")
                                  .Ret<F32>()   // loss    [N] | [1]
                                  .Ret<F32>()   // g_theta [N,7] | [7]
                                  .Ret<ffi::Buffer<ffi::U8>>());  // scratch: dynax_workspace_bytes(DYNAX_OP_RC_LOSS_GRAD, ...)

// ---- shared disturbance series + one per-system channel (dynax/wrapper/envs/rc.py:128-131): fused loss + gradient
static ffi::Error RcLossGradCtrl(cudaStream_t stream, F32 theta, F32 x0, F32 u_shared, F32 ctrl, F32 target, F32 lb, F32 ub,
                                 float dt, int32_t ctrl_col, int32_t flags, RF32 loss, RF32 g_theta, RU8 ws) {
  const int64_t T = u_shared.dimensions()[0], N = x0.dimensions()[0];
  const uint32_t f = (uint32_t)flags | (target.dimensions().size() == 1 ? DYNAX_SHARED_TARGET : 0u) |
                     (theta.dimensions().size() == 1 ? DYNAX_SHARED_THETA : 0u);
  return status(dynax_rc_loss_grad_ctrl_f32(theta.typed_data(), x0.typed_data(), u_shared.typed_data(), ctrl.typed_data(),
                                            ctrl_col, target.typed_data(), lb.typed_data(), ub.typed_data(),
                                            loss->typed_data(), g_theta->typed_data(), nullptr, N, T, dt, f,
                                            ws->typed_data(), ws->size_bytes(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(dynax_rc_loss_grad_ctrl, RcLossGradCtrl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<F32>()   // theta    [N,7] | [7]
                                  .Arg<F32>()   // x0       [N,3]
                                  .Arg<F32>()   // u_shared [T,5]
                                  .Arg<F32>()   // ctrl     [T,N]
                                  .Arg<F32>()   // target   [T,N] | [T]
                                  .Arg<F32>()   // lb [7]
                                  .Arg<F32>()   // ub [7]
                                  .Attr<float>("dt")
                                  .Attr<int32_t>("ctrl_col")
                                  .Attr<int32_t>("flags")
                                  .Ret<F32>()   // loss
                                  .Ret<F32>()   // g_theta
                                  .Ret<ffi::Buffer<ffi::U8>>());

// ---- dense linear block SSMs (dynax/core/{discrete,continuous}_block_state_space.py:5-54).  A,B,C,D are the
//      matrices [N|1, rows, cols] (the transposes of the Flax nn.Dense kernels; the Python front end transposes).
static ffi::Error LtiForward(cudaStream_t stream, F32 A, F32 B, F32 C, F32 D, F32 x0, F32 u, float dt, int32_t flags,
                             RF32 xsol, RF32 ysol) {
  const int64_t T = u.dimensions()[0], N = x0.dimensions()[0];
  const int n = (int)x0.dimensions()[1], m = (int)B.dimensions()[2], p = (int)C.dimensions()[1];
  const uint32_t f = (uint32_t)flags | (u.dimensions().size() == 2 ? DYNAX_SHARED_U : 0u) |
                     ((A.dimensions()[0] == 1 && N > 1) ? DYNAX_SHARED_THETA : 0u);
  return status(dynax_lti_forward_f32(A.typed_data(), B.typed_data(), C.typed_data(), D.typed_data(), x0.typed_data(),
                                      u.typed_data(), xsol->typed_data(), ysol->typed_data(), nullptr, N, T, n, m, p, dt, f,
                                      stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(dynax_lti_forward, LtiForward,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()   // A [L,n,n] B [L,n,m] C [L,p,n] D [L,p,m]
                                  .Arg<F32>()   // x0 [N,n]
                                  .Arg<F32>()   // u  [T,N,m] | [T,m]
                                  .Attr<float>("dt")
                                  .Attr<int32_t>("flags")
                                  .Ret<F32>()   // xsol [T,N,n]
                                  .Ret<F32>()); // ysol [T,N,p]

static ffi::Error LtiVjp(cudaStream_t stream, F32 A, F32 B, F32 C, F32 D, F32 x0, F32 u, F32 g_xsol, F32 g_ysol, float dt,
                         int32_t flags, RF32 gA, RF32 gB, RF32 gC, RF32 gD, RF32 g_x0, RF32 g_u, RU8 ws) {
  const int64_t T = u.dimensions()[0], N = x0.dimensions()[0];
  const int n = (int)x0.dimensions()[1], m = (int)B.dimensions()[2], p = (int)C.dimensions()[1];
  const uint32_t f = (uint32_t)flags | ((A.dimensions()[0] == 1 && N > 1) ? DYNAX_SHARED_THETA : 0u);
  return status(dynax_lti_vjp_f32(A.typed_data(), B.typed_data(), C.typed_data(), D.typed_data(), x0.typed_data(),
                                  u.typed_data(), g_xsol.typed_data(), g_ysol.typed_data(), gA->typed_data(),
                                  gB->typed_data(), gC->typed_data(), gD->typed_data(), g_x0->typed_data(),
                                  g_u->typed_data(), N, T, n, m, p, dt, f, ws->typed_data(), ws->size_bytes(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(dynax_lti_vjp, LtiVjp,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()   // A B C D
                                  .Arg<F32>()   // x0
                                  .Arg<F32>()   // u      [T,N,m]
                                  .Arg<F32>()   // g_xsol [T,N,n]
                                  .Arg<F32>()   // g_ysol [T,N,p]
                                  .Attr<float>("dt")
                                  .Attr<int32_t>("flags")
                                  .Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>()   // gA gB gC gD
                                  .Ret<F32>()   // g_x0
                                  .Ret<F32>()   // g_u
                                  .Ret<ffi::Buffer<ffi::U8>>());  // scratch: dynax_workspace_bytes(DYNAX_OP_LTI_VJP, ...)

// ---- general fx/fy model with MLP dynamics (dynax/core/base_block_state_space.py:62-63,72-73), RK4 or Euler.
//      stage_x [T,N,9] is a residual the backward pass reuses (it saves three MLP evaluations per step).
static ffi::Error NodeForward(cudaStream_t stream, F32 W1, F32 b1, F32 W2, F32 b2, F32 W3, F32 b3, F32 x0, F32 u, float dt,
                              int32_t flags, RF32 xsol, RF32 ysol, RF32 stage_x) {
  const int64_t T = u.dimensions()[0], N = x0.dimensions()[0];
  const int n = (int)x0.dimensions()[1], hidden = (int)b1.dimensions()[0], m = (int)W1.dimensions()[1] - n;
  return status(dynax_node_rk4_forward_stages_f32(W1.typed_data(), b1.typed_data(), W2.typed_data(), b2.typed_data(),
                                                  W3.typed_data(), b3.typed_data(), x0.typed_data(), u.typed_data(),
                                                  xsol->typed_data(), ysol->typed_data(), nullptr, stage_x->typed_data(), N, T,
                                                  n, m, 1, hidden, dt, (uint32_t)flags, stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(dynax_node_forward, NodeForward,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()   // W1 b1 W2 b2 W3 b3 ([out,in])
                                  .Arg<F32>()   // x0 [N,3]
                                  .Arg<F32>()   // u  [T,N,5]
                                  .Attr<float>("dt")
                                  .Attr<int32_t>("flags")   // DYNAX_SOLVER_EULER or 0
                                  .Ret<F32>()   // xsol    [T,N,3]
                                  .Ret<F32>()   // ysol    [T,N,1]
                                  .Ret<F32>()); // stage_x [T,N,9]

static ffi::Error NodeVjp(cudaStream_t stream, F32 W1, F32 b1, F32 W2, F32 b2, F32 W3, F32 b3, F32 x0, F32 u, F32 xsol,
                          F32 stage_x, F32 g_xsol, F32 g_ysol, float dt, int32_t flags, RF32 gW1, RF32 gb1, RF32 gW2,
                          RF32 gb2, RF32 gW3, RF32 gb3, RF32 g_x0, RF32 g_u, RU8 ws) {
  const int64_t T = u.dimensions()[0], N = x0.dimensions()[0];
  const int n = (int)x0.dimensions()[1], hidden = (int)b1.dimensions()[0], m = (int)W1.dimensions()[1] - n;
  const bool euler = ((uint32_t)flags & DYNAX_SOLVER_EULER) != 0;
  return status(dynax_node_rk4_vjp_stages_f32(W1.typed_data(), b1.typed_data(), W2.typed_data(), b2.typed_data(),
                                              W3.typed_data(), b3.typed_data(), x0.typed_data(), u.typed_data(),
                                              xsol.typed_data(), euler ? nullptr : stage_x.typed_data(), g_xsol.typed_data(),
                                              g_ysol.typed_data(), gW1->typed_data(), gb1->typed_data(), gW2->typed_data(),
                                              gb2->typed_data(), gW3->typed_data(), gb3->typed_data(), g_x0->typed_data(),
                                              g_u->typed_data(), N, T, n, m, 1, hidden, dt, (uint32_t)flags, ws->typed_data(),
                                              ws->size_bytes(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(dynax_node_vjp, NodeVjp,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()   // W1 b1 W2 b2 W3 b3
                                  .Arg<F32>()   // x0
                                  .Arg<F32>()   // u
                                  .Arg<F32>()   // xsol    (forward trajectory)
                                  .Arg<F32>()   // stage_x (forward stage points)
                                  .Arg<F32>()   // g_xsol
                                  .Arg<F32>()   // g_ysol
                                  .Attr<float>("dt")
                                  .Attr<int32_t>("flags")
                                  .Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>()   // gW1 gb1 gW2 gb2 gW3 gb3
                                  .Ret<F32>()   // g_x0
                                  .Ret<F32>()   // g_u
                                  .Ret<ffi::Buffer<ffi::U8>>());  // scratch: dynax_node_rk4_vjp_workspace()

// ---- sampling-based MPC (BASELINE.json config 4): cost of N candidate control sequences + argmin
//      (negated env reward, dynax/wrapper/envs/rc.py:227-252, summed over the horizon)
static ffi::Error RcMpcCost(cudaStream_t stream, F32 theta, F32 x0, F32 d, F32 u, F32 price, F32 t_low, F32 t_high, float dt,
                            float cop, float w_energy, float w_comfort, float start_time, RF32 cost,
                            ffi::ResultBuffer<ffi::S64> argmin, RF32 min_cost, RU8 ws) {
  const int64_t H = u.dimensions()[0], N = u.dimensions()[1];
  return status(dynax_rc_mpc_cost_f32(theta.typed_data(), x0.typed_data(), d.typed_data(), u.typed_data(), price.typed_data(),
                                      t_low.typed_data(), t_high.typed_data(), cop, w_energy, w_comfort, start_time,
                                      cost->typed_data(), reinterpret_cast<long long *>(argmin->typed_data()),
                                      min_cost->typed_data(), N, H, dt, ws->typed_data(), ws->size_bytes(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(dynax_rc_mpc_cost, RcMpcCost,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<F32>()   // theta [7]
                                  .Arg<F32>()   // x0    [3]
                                  .Arg<F32>()   // d     [H,4]
                                  .Arg<F32>()   // u     [H,N]
                                  .Arg<F32>().Arg<F32>().Arg<F32>()   // price, t_low, t_high [24]
                                  .Attr<float>("dt")
                                  .Attr<float>("cop")
                                  .Attr<float>("w_energy")
                                  .Attr<float>("w_comfort")
                                  .Attr<float>("start_time")
                                  .Ret<F32>()                          // cost [N]
                                  .Ret<ffi::Buffer<ffi::S64>>()        // argmin []
                                  .Ret<F32>()                          // min_cost []
                                  .Ret<ffi::Buffer<ffi::U8>>());       // scratch: 16 bytes
