"""NumPy-buffer front end of the *_host C-ABI entry points (no torch involved).

This is the call a reference user makes with host arrays (the reference's own arrays live on the
host when it runs on XLA:CPU): the library tiles the system axis, overlaps H2D / kernel / D2H on two
streams and returns NumPy results.  Pass pinned memory (e.g. ``torch.empty(..., pin_memory=True).numpy()``)
for full PCIe rate; pageable memory works but copies are staged by the driver.
"""
from __future__ import annotations

import numpy as np

from ._lib import DISCRETE, SHARED_TARGET, SHARED_THETA, SHARED_U, check, lib

__all__ = ["rc_forward", "rc_loss_grad", "rc_loss_grad_ctrl", "release", "RcDataset", "pinned"]


def _f32(a):
    a = np.asarray(a)
    if a.dtype != np.float32 or not a.flags["C_CONTIGUOUS"]:
        a = np.ascontiguousarray(a, dtype=np.float32)
    return a


def _p(a):
    return None if a is None else a.ctypes.data


def _canon(theta, x0, u):
    theta, x0, u = _f32(theta), _f32(x0), _f32(u)
    if x0.ndim != 2 or x0.shape[1] != 3:
        raise ValueError(f"x0 must be [N,3], got {x0.shape}")
    N = x0.shape[0]
    if theta.ndim == 1:
        theta = theta[None]
    if theta.shape[1:] != (7,) or theta.shape[0] not in (1, N):
        raise ValueError(f"theta must be [N,7] or [7], got {theta.shape}")
    if u.ndim == 2:
        u = u[:, None, :]
    if u.ndim != 3 or u.shape[2] != 5 or u.shape[1] not in (1, N):
        raise ValueError(f"u must be [T,N,5] or [T,5], got {u.shape}")
    flags = (SHARED_THETA if (theta.shape[0] == 1 and N != 1) else 0) | (SHARED_U if (u.shape[1] == 1 and N != 1) else 0)
    return theta, x0, u, N, u.shape[0], flags


def rc_forward(theta, x0, u, dt, emit_x=True, emit_y=True, emit_final=False, discrete=False, tile_systems=0,
               out=None):
    """theta[N,7]|[7], x0[N,3], u[T,N,5]|[T,5] (host) -> xsol[T,N,3], ysol[T,N,1], x_final[N,3] (host)."""
    theta, x0, u, N, T, flags = _canon(theta, x0, u)
    out = out or {}
    xs = out.get("xsol", np.empty((T, N, 3), np.float32)) if emit_x else None
    ys = out.get("ysol", np.empty((T, N, 1), np.float32)) if emit_y else None
    xf = out.get("x_final", np.empty((N, 3), np.float32)) if emit_final else None
    check(lib().dynax_rc_forward_host_f32(_p(theta), _p(x0), _p(u), _p(xs), _p(ys), _p(xf), N, T, float(dt),
                                          flags | (DISCRETE if discrete else 0), int(tile_systems)))
    return xs, ys, xf


def rc_loss_grad(theta, x0, u, target, dt, lb=None, ub=None, discrete=False, tile_systems=0):
    """Fused loss + gradient of examples/3-inverse.py:96-112 on host buffers -> loss[N|1], grad[N,7]|[7]."""
    theta, x0, u, N, T, flags = _canon(theta, x0, u)
    target = _f32(target)
    if target.shape == (T, N) and N != 1:
        pass
    elif target.size == T:
        target = target.reshape(T)
        flags |= SHARED_TARGET if N != 1 else 0
    else:
        raise ValueError(f"target must be [T,N] or [T], got {target.shape}")
    if (lb is None) != (ub is None):
        raise ValueError("lb and ub must be given together")
    lb = None if lb is None else _f32(lb).reshape(7)
    ub = None if ub is None else _f32(ub).reshape(7)
    shared = bool(flags & SHARED_THETA)
    loss = np.empty((1 if shared else N,), np.float32)
    grad = np.empty((7,) if shared else (N, 7), np.float32)
    check(lib().dynax_rc_loss_grad_host_f32(_p(theta), _p(x0), _p(u), _p(target), _p(lb), _p(ub), _p(loss), _p(grad),
                                            N, T, float(dt), flags | (DISCRETE if discrete else 0), int(tile_systems)))
    return loss, grad


def release():
    check(lib().dynax_host_release())


def rc_loss_grad_ctrl(theta, x0, u_shared, ctrl, target, dt, ctrl_col=2, lb=None, ub=None, tile_systems=0):
    """Host-buffer loss + gradient with a shared input series u_shared[T,5] and a per-system channel
    ctrl[T,N] (rc.py:128-131): only 4 B (+ target) per system-step cross PCIe instead of 24."""
    theta, x0 = _f32(theta), _f32(x0)
    u_shared, ctrl, target = _f32(u_shared).reshape(-1, 5), _f32(ctrl), _f32(target)
    N, T = x0.shape[0], u_shared.shape[0]
    if theta.ndim == 1:
        theta = theta[None]
    if ctrl.shape != (T, N):
        raise ValueError(f"ctrl must be [T,N]=({T},{N}), got {ctrl.shape}")
    flags = SHARED_THETA if (theta.shape[0] == 1 and N != 1) else 0
    if target.shape == (T, N) and N != 1:
        pass
    elif target.size == T:
        target = target.reshape(T)
        flags |= SHARED_TARGET if N != 1 else 0
    else:
        raise ValueError(f"target must be [T,N] or [T], got {target.shape}")
    lb = None if lb is None else _f32(lb).reshape(7)
    ub = None if ub is None else _f32(ub).reshape(7)
    shared = bool(flags & SHARED_THETA)
    loss = np.empty((1 if shared else N,), np.float32)
    grad = np.empty((7,) if shared else (N, 7), np.float32)
    check(lib().dynax_rc_loss_grad_ctrl_host_f32(_p(theta), _p(x0), _p(u_shared), _p(ctrl), int(ctrl_col), _p(target),
                                                 _p(lb), _p(ub), _p(loss), _p(grad), N, T, float(dt), flags,
                                                 int(tile_systems)))
    return loss, grad


class RcDataset:
    """u and target resident in device memory across calls (``dynax_rc_dataset_*``): the calibration loop of
    examples/3-inverse.py:133-138 re-evaluates ONE measured series for changing parameters, so per call only
    theta and x0 cross the link.  ``ds = RcDataset(u, target); loss, grad = ds.loss_grad(theta, x0, dt)``."""

    def __init__(self, u, target, n_systems=None):
        import ctypes
        u, target = _f32(u), _f32(target)
        if u.ndim == 2:
            u = u[:, None, :]
        if u.ndim != 3 or u.shape[2] != 5:
            raise ValueError(f"u must be [T,N,5] or [T,5], got {u.shape}")
        T = u.shape[0]
        if target.ndim == 2 and target.shape[0] == T:
            N = target.shape[1]
        elif target.size == T:
            N, target = u.shape[1], target.reshape(T)
        else:
            raise ValueError(f"target must be [T,N] or [T], got {target.shape}")
        N = max(N, u.shape[1], int(n_systems or 1))   # n_systems: needed when BOTH series are shared
        if u.shape[1] not in (1, N):
            raise ValueError("u and target disagree on the number of systems")
        self.N, self.T = N, T
        self._flags = (SHARED_U if (u.shape[1] == 1 and N != 1) else 0) | (SHARED_TARGET if (target.ndim == 1 and N != 1) else 0)
        h = ctypes.c_void_p()
        check(lib().dynax_rc_dataset_create_f32(_p(u), _p(target), N, T, self._flags, ctypes.byref(h)))
        self._h = h

    def loss_grad(self, theta, x0, dt, lb=None, ub=None, discrete=False):
        theta, x0 = _f32(theta), _f32(x0)
        if theta.ndim == 1:
            theta = theta[None]
        if x0.shape != (self.N, 3) or theta.shape[1:] != (7,) or theta.shape[0] not in (1, self.N):
            raise ValueError(f"theta must be [N,7] or [7] and x0 [N,3] with N={self.N}")
        if (lb is None) != (ub is None):
            raise ValueError("lb and ub must be given together")
        lb = None if lb is None else _f32(lb).reshape(7)
        ub = None if ub is None else _f32(ub).reshape(7)
        shared = theta.shape[0] == 1 and self.N != 1
        loss = np.empty((1 if shared else self.N,), np.float32)
        grad = np.empty((7,) if shared else (self.N, 7), np.float32)
        check(lib().dynax_rc_dataset_loss_grad_f32(self._h, _p(theta), _p(x0), _p(lb), _p(ub), float(dt),
                                                   (SHARED_THETA if shared else 0) | (DISCRETE if discrete else 0),
                                                   _p(loss), _p(grad)))
        return loss, grad

    def close(self):
        if getattr(self, "_h", None):
            check(lib().dynax_rc_dataset_destroy(self._h))
            self._h = None

    __del__ = close


class pinned:
    """Context manager: page-locks NumPy arrays for the duration of the block (``dynax_host_register``), so that the
    *_host calls copy them at the pinned PCIe rate instead of through the driver's staging buffer."""

    def __init__(self, *arrays):
        self._arrays = [a for a in arrays if a is not None]
        self._done = []

    def __enter__(self):
        for a in self._arrays:
            if a.dtype != np.float32 or not a.flags["C_CONTIGUOUS"]:
                raise ValueError("pin contiguous float32 arrays (the *_host calls would copy anything else first)")
            check(lib().dynax_host_register(a.ctypes.data, a.nbytes))
            self._done.append(a)
        return self

    def __exit__(self, *exc):
        for a in self._done:
            lib().dynax_host_unregister(a.ctypes.data)
        self._done = []
        return False
