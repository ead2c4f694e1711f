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
