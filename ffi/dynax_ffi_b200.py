"""JAX front end over ffi/dynax_ffi.so: jax.ffi custom calls + jax.custom_vjp.

Needs JAX >= 0.4.31 and the shim built as described at the top of ffi/dynax_ffi.cc.  NOT importable in this
repository's image (no JAX); the tested binding there is dynax_b200.ops (ctypes + torch).  The functions
below are what a maintainer of the reference would call from DifferentiableSimulator.__call__
(/root/reference/dynax/simulators/simulator.py:22-55) and from examples/3-inverse.py:96-112.
"""
import ctypes
import functools
import os

import jax
import jax.numpy as jnp

_HERE = os.path.dirname(os.path.abspath(__file__))
_core = ctypes.CDLL(os.path.join(_HERE, "..", "dynax_b200", "libdynax_b200.so"), mode=ctypes.RTLD_GLOBAL)
_shim = ctypes.CDLL(os.path.join(_HERE, "dynax_ffi.so"))
_core.dynax_workspace_bytes.restype = ctypes.c_size_t
_core.dynax_workspace_bytes.argtypes = [ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_int, ctypes.c_uint32]
OP_RC_LOSS_GRAD, OP_RC_VJP = 1, 2          # DYNAX_OP_* in include/dynax_b200.h

for _name in ("dynax_rc_forward", "dynax_rc_vjp", "dynax_rc_loss_grad"):
    jax.ffi.register_ffi_target(_name, jax.ffi.pycapsule(getattr(_shim, _name)), platform="CUDA")


def _f32(shape):
    return jax.ShapeDtypeStruct(tuple(shape), jnp.float32)


@functools.partial(jax.custom_vjp, nondiff_argnums=(3,))
def rc_rollout(theta, x0, u, dt):
    """xsol[T,N,3], ysol[T,N,1] for theta[N,7], x0[N,3], u[T,N,5]: one kernel launch for the whole batch."""
    T, N = u.shape[0], x0.shape[0]
    return jax.ffi.ffi_call("dynax_rc_forward", (_f32((T, N, 3)), _f32((T, N, 1))))(
        theta, x0, u, dt=jnp.float32(dt), flags=jnp.int32(0))


def _rc_rollout_fwd(theta, x0, u, dt):
    return rc_rollout(theta, x0, u, dt), (theta, x0, u)      # residuals: the inputs (checkpoints are internal)


def _rc_rollout_bwd(dt, res, cot):
    theta, x0, u = res
    g_xsol, g_ysol = cot
    T, N = u.shape[0], x0.shape[0]
    ws = _core.dynax_workspace_bytes(OP_RC_VJP, N, T, 3, 5, 1, 0)
    g_theta, g_x0, g_u, _ = jax.ffi.ffi_call(
        "dynax_rc_vjp", (_f32(theta.shape), _f32(x0.shape), _f32(u.shape), jax.ShapeDtypeStruct((ws,), jnp.uint8)))(
        theta, x0, u, g_xsol, g_ysol, dt=jnp.float32(dt), flags=jnp.int32(0))
    return g_theta, g_x0, g_u


rc_rollout.defvjp(_rc_rollout_fwd, _rc_rollout_bwd)


def rc_loss_and_grad(theta, x0, u, target, dt, lb=None, ub=None):
    """Fused replacement of jax.value_and_grad(mse_loss) (examples/3-inverse.py:96-112) for N candidates."""
    T, N = u.shape[0], x0.shape[0]
    lb = jnp.full((7,), -jnp.inf, jnp.float32) if lb is None else jnp.asarray(lb, jnp.float32)
    ub = jnp.full((7,), jnp.inf, jnp.float32) if ub is None else jnp.asarray(ub, jnp.float32)
    nout = 1 if theta.ndim == 1 else N
    ws = _core.dynax_workspace_bytes(OP_RC_LOSS_GRAD, N, T, 3, 5, 1, 0)
    loss, grad, _ = jax.ffi.ffi_call(
        "dynax_rc_loss_grad", (_f32((nout,)), _f32(theta.shape), jax.ShapeDtypeStruct((ws,), jnp.uint8)))(
        theta, x0, u, target, lb, ub, dt=jnp.float32(dt), flags=jnp.int32(0))
    return loss, grad
