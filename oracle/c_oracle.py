"""ctypes loader for oracle/dynax_oracle.c (TEST INFRASTRUCTURE ONLY -- see that file's header)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(native=False):
    subprocess.check_call(["make", "-s", "-C", _DIR, "native" if native else "all"])


def lib(native=False):
    """Load (building if needed) the C oracle.  ``native=True`` prefers a -march=native build."""
    global _LIB
    if _LIB is not None:
        return _LIB
    names = (["libdynax_oracle_native.so"] if native else []) + ["libdynax_oracle.so"]
    # make rebuilds whatever is older than dynax_oracle.c (a stale .so travels to the GPU box with the snapshot)
    for tgt in (["native"] if native else []) + ["all"]:
        try:
            build(native=(tgt == "native"))
        except Exception:
            if tgt == "all" and not os.path.exists(os.path.join(_DIR, "libdynax_oracle.so")):
                raise
    for nm in names:
        path = os.path.join(_DIR, nm)
        if os.path.exists(path):
            L = ctypes.CDLL(path)
            break
    else:
        raise RuntimeError("C oracle not built")
    fp = ctypes.POINTER(ctypes.c_float)
    L.oracle_rc_forward_f32.argtypes = [fp] * 6 + [ctypes.c_long, ctypes.c_long, ctypes.c_float, ctypes.c_int, ctypes.c_int]
    L.oracle_rc_loss_grad_f32.argtypes = [fp] * 8 + [ctypes.c_long, ctypes.c_long, ctypes.c_float] + [ctypes.c_int] * 3
    dp = ctypes.POINTER(ctypes.c_double)
    L.oracle_rc_loss_grad_budget_f64.argtypes = [fp] * 4 + [dp] * 6 + [ctypes.c_long, ctypes.c_long, ctypes.c_double] + \
        [ctypes.c_int] * 3 + [ctypes.c_double]
    L.oracle_num_threads.restype = ctypes.c_int
    L.oracle_set_num_threads.argtypes = [ctypes.c_int]
    _LIB = L
    return L


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def num_threads():
    return lib().oracle_num_threads()


def use_all_cores():
    """Override OMP_NUM_THREADS (torchrun exports OMP_NUM_THREADS=1) with every host core."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().oracle_set_num_threads(n)
    return n


def rc_forward(theta, x0, u, dt, emit_x=True, emit_y=True, native=False):
    """theta[N|1,7], x0[N,3], u[T,N|1,5] -> xsol[T,N,3], ysol[T,N,1], x_final[N,3]."""
    theta, x0, u = _f32(np.atleast_2d(theta)), _f32(x0), _f32(u)
    T, N = u.shape[0], x0.shape[0]
    xs = np.empty((T, N, 3), np.float32) if emit_x else None
    ys = np.empty((T, N, 1), np.float32) if emit_y else None
    xf = np.empty((N, 3), np.float32)
    rc = lib(native).oracle_rc_forward_f32(_p(theta), _p(x0), _p(u), _p(xs), _p(ys), _p(xf), N, T, dt,
                                           int(theta.shape[0] == 1 and N > 1), int(u.shape[1] == 1 and N > 1))
    assert rc == 0
    return xs, ys, xf


def rc_loss_grad(theta, x0, u, target, dt, lb=None, ub=None, native=False):
    """All-fp32 BPTT like the reference's default dtype: loss[N], grad[N,7]."""
    theta, x0, u, target = _f32(np.atleast_2d(theta)), _f32(x0), _f32(u), _f32(target)
    lb, ub = _f32(lb), _f32(ub)
    T, N = u.shape[0], x0.shape[0]
    loss = np.empty((N,), np.float32)
    grad = np.empty((N, 7), np.float32)
    rc = lib(native).oracle_rc_loss_grad_f32(_p(theta), _p(x0), _p(u), _p(target), _p(lb), _p(ub), _p(loss), _p(grad),
                                             N, T, dt, int(theta.shape[0] == 1 and N > 1),
                                             int(u.shape[1] == 1 and N > 1), int(target.ndim == 1 or target.shape[1] == 1 and N > 1))
    assert rc == 0
    return loss, grad


def rc_loss_grad_budget_f64(theta, x0, u, target, dt, lb=None, ub=None, traj_rtol=1e-5, native=False):
    """fp64 forward-mode loss[N], grad[N,7] (box penalty included when lb/ub are given) and the conditioning budgets
    loss_budget[N], grad_budget[N,7] of dynax_oracle.rc_output_sensitivity_budget, in one OpenMP sweep."""
    theta, x0, u, target = _f32(np.atleast_2d(theta)), _f32(x0), _f32(u), _f32(target)
    T, N = u.shape[0], x0.shape[0]
    dp = ctypes.POINTER(ctypes.c_double)
    d = lambda a: None if a is None else a.ctypes.data_as(dp)
    lb64 = None if lb is None else np.ascontiguousarray(lb, dtype=np.float64)
    ub64 = None if ub is None else np.ascontiguousarray(ub, dtype=np.float64)
    loss, grad = np.empty((N,), np.float64), np.empty((N, 7), np.float64)
    lbud, gbud = np.empty((N,), np.float64), np.empty((N, 7), np.float64)
    rc = lib(native).oracle_rc_loss_grad_budget_f64(
        _p(theta), _p(x0), _p(u), _p(target), d(lb64), d(ub64), d(loss), d(grad), d(lbud), d(gbud), N, T, float(dt),
        int(theta.shape[0] == 1 and N > 1), int(u.shape[1] == 1 and N > 1),
        int(target.ndim == 1 or target.shape[1] == 1 and N > 1), float(traj_rtol))
    assert rc == 0
    return loss, grad, lbud, gbud
