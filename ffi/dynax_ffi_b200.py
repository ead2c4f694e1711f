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
_core.dynax_node_rk4_vjp_workspace.restype = ctypes.c_size_t
OP_RC_LOSS_GRAD, OP_RC_VJP, OP_LTI_VJP = 1, 2, 4          # DYNAX_OP_* in include/dynax_b200.h
SHARED_THETA, SOLVER_EULER = 1, 16                        # flags in include/dynax_b200.h

for _name in ("dynax_rc_forward", "dynax_rc_vjp", "dynax_rc_loss_grad", "dynax_rc_loss_grad_ctrl", "dynax_lti_forward",
              "dynax_lti_vjp", "dynax_node_forward", "dynax_node_vjp", "dynax_rc_mpc_cost"):
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
    ws = max(16, _core.dynax_workspace_bytes(OP_RC_LOSS_GRAD, N, T, 3, 5, 1, SHARED_THETA if theta.ndim == 1 else 0))
    loss, grad, _ = jax.ffi.ffi_call(
        "dynax_rc_loss_grad", (_f32((nout,)), _f32(theta.shape), jax.ShapeDtypeStruct((ws,), jnp.uint8)))(
        theta, x0, u, target, lb, ub, dt=jnp.float32(dt), flags=jnp.int32(0))
    return loss, grad


def rc_loss_and_grad_ctrl(theta, x0, u_shared, ctrl, target, dt, ctrl_col=2, lb=None, ub=None):
    """Fused loss + gradient when the disturbances are ONE shared series u_shared[T,5] and only column ``ctrl_col``
    is per system (ctrl[T,N]): how dynax/wrapper/envs/rc.py:128-131 assembles the env inputs."""
    T, N = u_shared.shape[0], x0.shape[0]
    lb = jnp.full((7,), -jnp.inf, jnp.float32) if lb is None else jnp.asarray(lb, jnp.float32)
    ub = jnp.full((7,), jnp.inf, jnp.float32) if ub is None else jnp.asarray(ub, jnp.float32)
    nout = 1 if theta.ndim == 1 else N
    ws = max(16, _core.dynax_workspace_bytes(OP_RC_LOSS_GRAD, N, T, 3, 5, 1, SHARED_THETA if theta.ndim == 1 else 0))
    loss, grad, _ = jax.ffi.ffi_call(
        "dynax_rc_loss_grad_ctrl", (_f32((nout,)), _f32(theta.shape), jax.ShapeDtypeStruct((ws,), jnp.uint8)))(
        theta, x0, u_shared, ctrl, target, lb, ub, dt=jnp.float32(dt), ctrl_col=jnp.int32(ctrl_col), flags=jnp.int32(0))
    return loss, grad


# ---- dense linear block SSMs: Discrete/ContinuousLinearSSM (dynax/core/*_block_state_space.py:5-54) -------------
def _mats(params):
    """Flax ``{'_fxx': {'dense': {'kernel': [n,n]}}, '_fxu': ..., '_fyx': ..., '_fyu': ...}`` -> A,B,C,D as [1,rows,cols]
    matrices (nn.Dense computes x @ kernel, so the matrix is the kernel transposed)."""
    return tuple(jnp.swapaxes(params[k]["dense"]["kernel"], -1, -2)[None] for k in ("_fxx", "_fxu", "_fyx", "_fyu"))


@functools.partial(jax.custom_vjp, nondiff_argnums=(6,))
def lti_rollout(A, B, C, D, x0, u, dt):
    """xsol[T,N,n], ysol[T,N,p] for matrices A[L,n,n] B[L,n,m] C[L,p,n] D[L,p,m] (L = N or 1), x0[N,n], u[T,N,m]."""
    T, N, n, p = u.shape[0], x0.shape[0], x0.shape[1], C.shape[1]
    return jax.ffi.ffi_call("dynax_lti_forward", (_f32((T, N, n)), _f32((T, N, p))))(
        A, B, C, D, x0, u, dt=jnp.float32(dt), flags=jnp.int32(0))


def _lti_fwd(A, B, C, D, x0, u, dt):
    return lti_rollout(A, B, C, D, x0, u, dt), (A, B, C, D, x0, u)


def _lti_bwd(dt, res, cot):
    A, B, C, D, x0, u = res
    T, N, n, m, p = u.shape[0], x0.shape[0], x0.shape[1], B.shape[2], C.shape[1]
    ws = _core.dynax_workspace_bytes(OP_LTI_VJP, N, T, n, m, p, 0)
    out = jax.ffi.ffi_call(
        "dynax_lti_vjp", (_f32(A.shape), _f32(B.shape), _f32(C.shape), _f32(D.shape), _f32(x0.shape), _f32(u.shape),
                          jax.ShapeDtypeStruct((ws,), jnp.uint8)))(A, B, C, D, x0, u, cot[0], cot[1], dt=jnp.float32(dt),
                                                                    flags=jnp.int32(0))
    return tuple(out[:6])


lti_rollout.defvjp(_lti_fwd, _lti_bwd)


# ---- general fx/fy model with MLP dynamics (base_block_state_space.py:62-63,72-73), RK4 (or Euler) ---------------
@functools.partial(jax.custom_vjp, nondiff_argnums=(8, 9))
def node_rollout(W1, b1, W2, b2, W3, b3, x0, u, dt, euler=False):
    """xsol[T,N,3], ysol[T,N,1] of rhs = W3 tanh(W2 tanh(W1 [x;u] + b1) + b2) + b3 (weights [out,in])."""
    return _node_fwd(W1, b1, W2, b2, W3, b3, x0, u, dt, euler)[0]


def _node_call(W1, b1, W2, b2, W3, b3, x0, u, dt, euler):
    T, N = u.shape[0], x0.shape[0]
    return jax.ffi.ffi_call("dynax_node_forward", (_f32((T, N, 3)), _f32((T, N, 1)), _f32((T, N, 9))))(
        W1, b1, W2, b2, W3, b3, x0, u, dt=jnp.float32(dt), flags=jnp.int32(SOLVER_EULER if euler else 0))


def _node_fwd(W1, b1, W2, b2, W3, b3, x0, u, dt, euler):
    xsol, ysol, stage_x = _node_call(W1, b1, W2, b2, W3, b3, x0, u, dt, euler)
    return (xsol, ysol), (W1, b1, W2, b2, W3, b3, x0, u, xsol, stage_x)


def _node_bwd(dt, euler, res, cot):
    W1, b1, W2, b2, W3, b3, x0, u, xsol, stage_x = res
    ws = _core.dynax_node_rk4_vjp_workspace()
    shapes = tuple(_f32(t.shape) for t in (W1, b1, W2, b2, W3, b3, x0, u)) + (jax.ShapeDtypeStruct((ws,), jnp.uint8),)
    out = jax.ffi.ffi_call("dynax_node_vjp", shapes)(W1, b1, W2, b2, W3, b3, x0, u, xsol, stage_x, cot[0], cot[1],
                                                     dt=jnp.float32(dt), flags=jnp.int32(SOLVER_EULER if euler else 0))
    return tuple(out[:8])


node_rollout.defvjp(_node_fwd, _node_bwd)


# ---- sampling-based MPC (BASELINE.json config 4) -----------------------------------------------------------------
def rc_mpc_cost(theta, x0, d, u, price, t_low, t_high, dt, cop=2.5, w_energy=100.0, w_comfort=1.0, start_time=0.0):
    """cost[N] of N candidate control sequences u[H,N] through one shared 4R3C model + (argmin, min_cost)."""
    N = u.shape[1]
    cost, argmin, cmin, _ = jax.ffi.ffi_call(
        "dynax_rc_mpc_cost", (_f32((N,)), jax.ShapeDtypeStruct((), jnp.int64), _f32(()), jax.ShapeDtypeStruct((16,), jnp.uint8)))(
        theta, x0, d, u, price, t_low, t_high, dt=jnp.float32(dt), cop=jnp.float32(cop), w_energy=jnp.float32(w_energy),
        w_comfort=jnp.float32(w_comfort), start_time=jnp.float32(start_time))
    return cost, argmin, cmin
