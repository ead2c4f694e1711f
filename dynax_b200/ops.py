"""Torch-tensor glue over the C-ABI (device pointers + the current CUDA stream).

PyTorch is plumbing here (device memory, streams, autograd bookkeeping); every number is
produced by libdynax_b200.so.  All tensors must be CUDA float32; nothing falls back to torch ops.

The autograd Functions are the ``custom_vjp`` analogue of the north-star's
``jax.ffi`` + ``jax.custom_vjp`` binding (see INTEGRATION.md for that stub).
"""
from __future__ import annotations

import os

import torch

from . import _lib
from ._lib import DISCRETE, SHARED_TARGET, SHARED_THETA, SHARED_U, check, lib

__all__ = ["rc_forward", "rc_loss_grad", "rc_vjp", "lti_forward", "lti_vjp", "rc_rollout", "lti_rollout",
           "lti_supported"]


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _f32c(t, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise ValueError(f"{name}: must be a CUDA tensor (dynax_b200 has no CPU path); "
                         "use the *_host functions for NumPy buffers")
    if t.dtype != torch.float32:
        raise TypeError(f"{name}: must be float32, got {t.dtype}")
    return t.contiguous()


def _canon_batch(theta_like_dims, x0, u, n, m):
    """Normalise shapes.  Returns (x0[N,n], u[T,Nu,m], N, T, shared_u)."""
    if x0.dim() != 2 or x0.shape[1] != n:
        raise ValueError(f"x0 must be [N,{n}], got {tuple(x0.shape)}")
    N = x0.shape[0]
    if u.dim() == 2:
        u = u[:, None, :]
    if u.dim() != 3 or u.shape[2] != m:
        raise ValueError(f"u must be [T,N,{m}] (or [T,1,{m}] / [T,{m}] shared), got {tuple(u.shape)}")
    T = u.shape[0]
    if u.shape[1] == N and N != 1:
        shared_u = False
    elif u.shape[1] == 1:
        shared_u = N != 1
    else:
        raise ValueError(f"u has {u.shape[1]} systems but x0 has {N}")
    return x0, u, N, T, shared_u


def _ws(nbytes, device):
    return torch.empty((max(int(nbytes), 16),), dtype=torch.uint8, device=device)


# ----------------------------------------------------------------------------------------------
# RC 4R3C
# ----------------------------------------------------------------------------------------------
def _rc_theta(theta, N):
    theta = _f32c(theta, "theta")
    if theta.dim() == 1:
        theta = theta[None]
    if theta.dim() != 2 or theta.shape[1] != 7:
        raise ValueError(f"theta must be [N,7] or [7], got {tuple(theta.shape)}")
    if theta.shape[0] == N and N != 1:
        return theta, False
    if theta.shape[0] == 1:
        return theta, N != 1
    raise ValueError(f"theta has {theta.shape[0]} systems but x0 has {N}")


def rc_forward(theta, x0, u, dt, emit_x=True, emit_y=True, emit_final=False, discrete=False, time_chunks=None):
    """Batched ``DifferentiableSimulator(Continuous4R3C, dt).apply`` (simulator.py:22-55, RC.py:137-249).

    theta[N,7]|[7], x0[N,3], u[T,N,5]|[T,1,5]|[T,5]  ->  xsol[T,N,3], ysol[T,N,1], x_final[N,3]
    (entries not requested are None).

    ``time_chunks``: None = serial-in-time kernel unless the launch cannot fill the GPU with systems alone
    (N <= 8192 and T >= 256: measured break-evens; a single building over four weeks, examples/1-forward.py,
    is 0.045 ms instead of 0.24 ms), where the time-parallel chunked kernel is used with an automatic chunk
    length; 0 = always serial; k > 0 = k chunks.
    """
    x0, u = _f32c(x0, "x0"), _f32c(u, "u")
    x0, u, N, T, shared_u = _canon_batch(None, x0, u, 3, 5)
    theta, shared_th = _rc_theta(theta, N)
    dev = x0.device
    xsol = torch.empty((T, N, 3), dtype=torch.float32, device=dev) if emit_x else None
    ysol = torch.empty((T, N, 1), dtype=torch.float32, device=dev) if emit_y else None
    xfin = torch.empty((N, 3), dtype=torch.float32, device=dev) if emit_final else None
    flags = (SHARED_THETA if shared_th else 0) | (SHARED_U if shared_u else 0) | (DISCRETE if discrete else 0)
    if time_chunks is None:
        time_chunks = -1 if (not discrete and 0 < N <= 8192 and T >= 256) else 0
    with torch.cuda.device(dev):
        if time_chunks != 0 and not discrete and N > 0 and T > 0:
            chunk_len = 0 if time_chunks < 0 else -(-T // int(time_chunks))
            ws = _ws(lib().dynax_rc_forward_chunked_workspace(N, T, chunk_len), dev)
            check(lib().dynax_rc_forward_chunked_f32(_ptr(theta), _ptr(x0), _ptr(u), _ptr(xsol), _ptr(ysol), _ptr(xfin),
                                                     N, T, float(dt), flags, chunk_len, _ptr(ws), ws.numel(), _stream()))
        else:
            check(lib().dynax_rc_forward_f32(_ptr(theta), _ptr(x0), _ptr(u), _ptr(xsol), _ptr(ysol), _ptr(xfin),
                                             N, T, float(dt), flags, _stream()))
    return xsol, ysol, xfin


def rc_loss_grad(theta, x0, u, target, dt, lb=None, ub=None, want_gx0=False, discrete=False, time_chunks=None,
                 out=None):
    """Fused MSE(+box penalty) loss and its gradient (examples/3-inverse.py:96-112) for N candidates.

    theta[N,7] -> loss[N], grad[N,7];  theta[7] (shared) -> loss[1] (sum over systems), grad[7].
    target[T,N] or [T] (shared).  Returns (loss, grad, g_x0 or None).
    ``out``: optional (loss, grad) float32 CUDA tensors to write into -- e.g. two views of ONE [8] buffer for a
    shared parameter vector, so that the all-reduce over ranks (distributed.py) needs no packing kernel.
    ``time_chunks``: None = time-parallel chunked kernels only for small long launches (N <= 4096 and
    T >= 512 -- measured break-evens against the serial tangent kernel), 0 = never, k > 0 = k chunks.
    """
    x0, u, target = _f32c(x0, "x0"), _f32c(u, "u"), _f32c(target, "target")
    x0, u, N, T, shared_u = _canon_batch(None, x0, u, 3, 5)
    theta, shared_th = _rc_theta(theta, N)
    if target.dim() == 2 and target.shape == (T, N) and N != 1:
        shared_tg = False
    elif target.numel() == T:
        shared_tg = N != 1
        target = target.reshape(T)
    else:
        raise ValueError(f"target must be [T,N] or [T], got {tuple(target.shape)}")
    dev = x0.device
    if (lb is None) != (ub is None):
        raise ValueError("lb and ub must be given together")
    if lb is not None:
        lb = _f32c(torch.as_tensor(lb, dtype=torch.float32, device=dev), "lb").reshape(7)
        ub = _f32c(torch.as_tensor(ub, dtype=torch.float32, device=dev), "ub").reshape(7)
    nout = 1 if shared_th else N
    if out is not None:
        loss, grad = out
        for t, nm, cnt in ((loss, "out[0]", nout), (grad, "out[1]", nout * 7)):
            if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
                    and t.numel() == cnt):
                raise ValueError(f"{nm} must be a contiguous float32 CUDA tensor of {cnt} elements")
        grad = grad.view(nout, 7)
    else:
        loss = torch.empty((nout,), dtype=torch.float32, device=dev)
        grad = torch.empty((nout, 7), dtype=torch.float32, device=dev)
    gx0 = torch.empty((N, 3), dtype=torch.float32, device=dev) if want_gx0 else None
    flags = ((SHARED_THETA if shared_th else 0) | (SHARED_U if shared_u else 0) |
             (SHARED_TARGET if shared_tg else 0) | (DISCRETE if discrete else 0))
    if time_chunks is None:
        small_long = N <= 4096 and T >= 512
        time_chunks = -1 if (not discrete and N > 0 and small_long) else 0
    with torch.cuda.device(dev):
        if time_chunks != 0 and not discrete and N > 0:
            chunk_len = 0 if time_chunks < 0 else -(-T // int(time_chunks))
            ws = _ws(lib().dynax_rc_loss_grad_chunked_workspace(N, T, chunk_len), dev)
            check(lib().dynax_rc_loss_grad_chunked_f32(_ptr(theta), _ptr(x0), _ptr(u), _ptr(target), _ptr(lb), _ptr(ub),
                                                       _ptr(loss), _ptr(grad), _ptr(gx0), N, T, float(dt), flags,
                                                       chunk_len, _ptr(ws), ws.numel(), _stream()))
        else:
            ws = _ws(lib().dynax_workspace_bytes(_lib.OP_RC_LOSS_GRAD, N, T, 3, 5, 1, flags), dev)
            check(lib().dynax_rc_loss_grad_f32(_ptr(theta), _ptr(x0), _ptr(u), _ptr(target), _ptr(lb), _ptr(ub),
                                               _ptr(loss), _ptr(grad), _ptr(gx0), N, T, float(dt), flags,
                                               _ptr(ws), ws.numel(), _stream()))
    if shared_th:
        grad = grad.reshape(7)
    return loss, grad, gx0


def rc_vjp(theta, x0, u, dt, g_xsol=None, g_ysol=None, need_theta=True, need_x0=True, need_u=True, discrete=False,
           time_chunks=None):
    """VJP of rc_forward for arbitrary cotangents g_xsol[T,N,3], g_ysol[T,N,1] -> (g_theta, g_x0, g_u).

    ``time_chunks``: None = the time-parallel chunked adjoint for launches that cannot fill the GPU with systems alone
    (N <= 8192 and T >= 512: measured break-evens; continuous stepping), 0 = serial, k > 0 = k chunks."""
    x0, u = _f32c(x0, "x0"), _f32c(u, "u")
    x0, u, N, T, shared_u = _canon_batch(None, x0, u, 3, 5)
    theta, shared_th = _rc_theta(theta, N)
    dev = x0.device
    if g_xsol is not None:
        g_xsol = _f32c(g_xsol, "g_xsol").reshape(T, N, 3)
    if g_ysol is not None:
        g_ysol = _f32c(g_ysol, "g_ysol").reshape(T, N, 1)
    gth = torch.empty((1 if shared_th else N, 7), dtype=torch.float32, device=dev) if need_theta else None
    gx0 = torch.empty((N, 3), dtype=torch.float32, device=dev) if need_x0 else None
    # a shared input series has ONE cotangent row per step: the kernel sums g_u over the systems (no [T,N,5] array)
    u_k = u
    gu = torch.empty((T, 1 if shared_u else N, 5), dtype=torch.float32, device=dev) if need_u else None
    flags = ((SHARED_THETA if shared_th else 0) | (SHARED_U if shared_u else 0) | (DISCRETE if discrete else 0))
    if time_chunks is None:
        time_chunks = -1 if (not discrete and 0 < N <= 8192 and T >= 512) else 0
    chunked = time_chunks != 0 and not discrete and not (flags & SHARED_U) and N > 0
    with torch.cuda.device(dev):
        if chunked:
            chunk_len = 0 if time_chunks < 0 else -(-T // int(time_chunks))
            ws = _ws(lib().dynax_rc_vjp_chunked_workspace(N, T, chunk_len), dev)
            check(lib().dynax_rc_vjp_chunked_f32(_ptr(theta), _ptr(x0), _ptr(u_k), _ptr(g_xsol), _ptr(g_ysol), _ptr(gth),
                                                 _ptr(gx0), _ptr(gu), N, T, float(dt), flags & SHARED_THETA, chunk_len,
                                                 _ptr(ws), ws.numel(), _stream()))
        else:
            ws = _ws(lib().dynax_workspace_bytes(_lib.OP_RC_VJP, N, T, 3, 5, 1, flags), dev)
            check(lib().dynax_rc_vjp_f32(_ptr(theta), _ptr(x0), _ptr(u_k), _ptr(g_xsol), _ptr(g_ysol), _ptr(gth),
                                         _ptr(gx0), _ptr(gu), N, T, float(dt), flags, _ptr(ws), ws.numel(), _stream()))
    return gth, gx0, gu


class _RCRollout(torch.autograd.Function):
    """xsol, ysol = rollout(theta, x0, u); backward = one dynax_rc_vjp_f32 launch."""

    @staticmethod
    def forward(ctx, theta, x0, u, dt, discrete):
        xsol, ysol, _ = rc_forward(theta, x0, u, dt, discrete=discrete)
        ctx.save_for_backward(theta, x0, u)
        ctx.dt, ctx.discrete = dt, discrete
        return xsol, ysol

    @staticmethod
    def backward(ctx, g_xsol, g_ysol):
        theta, x0, u = ctx.saved_tensors
        nt, nx, nu = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        if g_xsol is None and g_ysol is None:
            return None, None, None, None, None
        gth, gx0, gu = rc_vjp(theta, x0, u, ctx.dt, g_xsol, g_ysol, nt, nx, nu, ctx.discrete)
        if gth is not None:
            gth = gth.reshape(theta.shape)
        if gu is not None:
            gu = gu.reshape(u.shape)
        return gth, gx0, gu, None, None


def rc_rollout(theta, x0, u, dt, discrete=False):
    """Differentiable (torch autograd) batched RC rollout: returns xsol[T,N,3], ysol[T,N,1]."""
    return _RCRollout.apply(theta, x0, u, float(dt), bool(discrete))


# ----------------------------------------------------------------------------------------------
# dense LTI
# ----------------------------------------------------------------------------------------------
def lti_supported(n, m, p):
    return bool(lib().dynax_lti_supported(int(n), int(m), int(p)))


def _lti_mats(A, B, C, D, N):
    A, B, C, D = (_f32c(M, nm) for M, nm in ((A, "A"), (B, "B"), (C, "C"), (D, "D")))
    mats = [M[None] if M.dim() == 2 else M for M in (A, B, C, D)]
    n, m, p = mats[0].shape[-1], mats[1].shape[-1], mats[2].shape[-2]
    want = [(n, n), (n, m), (p, n), (p, m)]
    lead = {M.shape[0] for M in mats}
    for M, w, nm in zip(mats, want, "ABCD"):
        if M.dim() != 3 or tuple(M.shape[1:]) != w:
            raise ValueError(f"{nm} must be [N|1,{w[0]},{w[1]}], got {tuple(M.shape)}")
    if len(lead) != 1:
        raise ValueError("A,B,C,D must share their leading (system) dimension")
    L = lead.pop()
    if L == N and N != 1:
        shared = False
    elif L == 1:
        shared = N != 1
    else:
        raise ValueError(f"matrices have {L} systems but x0 has {N}")
    if not lti_supported(n, m, p):
        raise _lib.DynaxError(-2, f"no dense-LTI kernel instantiated for (n,m,p)=({n},{m},{p})")
    return [M.contiguous() for M in mats], n, m, p, shared


def lti_forward(A, B, C, D, x0, u, dt, emit_x=True, emit_y=True, emit_final=False, discrete=False, time_chunks=None):
    """Batched rollout of a dense linear block SSM (A,B,C,D are matrices = Flax kernels transposed).

    ``time_chunks``: as in ``rc_forward`` -- None = the time-parallel chunked kernels for launches that cannot fill
    the GPU with systems alone (N <= 4096 and T >= 256, continuous stepping), 0 = serial, k > 0 = k chunks."""
    x0, u = _f32c(x0, "x0"), _f32c(u, "u")
    if x0.dim() != 2:
        raise ValueError("x0 must be [N,n]")
    N = x0.shape[0]
    (A, B, C, D), n, m, p, shared_th = _lti_mats(A, B, C, D, N)
    x0, u, N, T, shared_u = _canon_batch(None, x0, u, n, m)
    dev = x0.device
    xsol = torch.empty((T, N, n), dtype=torch.float32, device=dev) if emit_x else None
    ysol = torch.empty((T, N, p), dtype=torch.float32, device=dev) if emit_y else None
    xfin = torch.empty((N, n), dtype=torch.float32, device=dev) if emit_final else None
    flags = (SHARED_THETA if shared_th else 0) | (SHARED_U if shared_u else 0) | (DISCRETE if discrete else 0)
    if time_chunks is None:
        time_chunks = -1 if (not discrete and 0 < N <= 4096 and T >= 256) else 0
    with torch.cuda.device(dev):
        if time_chunks != 0 and not discrete and N > 0 and T > 0:
            chunk_len = 0 if time_chunks < 0 else -(-T // int(time_chunks))
            ws = _ws(lib().dynax_lti_forward_chunked_workspace(N, T, n, chunk_len), dev)
            check(lib().dynax_lti_forward_chunked_f32(_ptr(A), _ptr(B), _ptr(C), _ptr(D), _ptr(x0), _ptr(u), _ptr(xsol),
                                                      _ptr(ysol), _ptr(xfin), N, T, n, m, p, float(dt), flags, chunk_len,
                                                      _ptr(ws), ws.numel(), _stream()))
        else:
            check(lib().dynax_lti_forward_f32(_ptr(A), _ptr(B), _ptr(C), _ptr(D), _ptr(x0), _ptr(u), _ptr(xsol),
                                              _ptr(ysol), _ptr(xfin), N, T, n, m, p, float(dt), flags, _stream()))
    return xsol, ysol, xfin


def lti_vjp(A, B, C, D, x0, u, dt, g_xsol=None, g_ysol=None, need_mats=True, need_x0=True, need_u=True,
            discrete=False, time_chunks=None):
    """VJP of lti_forward -> (gA, gB, gC, gD, g_x0, g_u); matrix grads are summed over systems if shared.

    ``time_chunks``: None = the time-parallel chunked adjoint for launches that cannot fill the GPU with systems alone
    (N <= 4096 and T >= 512: measured break-evens; continuous stepping), 0 = serial, k > 0 = k chunks."""
    x0, u = _f32c(x0, "x0"), _f32c(u, "u")
    N = x0.shape[0]
    (A, B, C, D), n, m, p, shared_th = _lti_mats(A, B, C, D, N)
    x0, u, N, T, shared_u = _canon_batch(None, x0, u, n, m)
    dev = x0.device
    if g_xsol is not None:
        g_xsol = _f32c(g_xsol, "g_xsol").reshape(T, N, n)
    if g_ysol is not None:
        g_ysol = _f32c(g_ysol, "g_ysol").reshape(T, N, p)
    L = 1 if shared_th else N
    gm = [torch.empty((L,) + s, dtype=torch.float32, device=dev) if need_mats else None
          for s in ((n, n), (n, m), (p, n), (p, m))]
    gx0 = torch.empty((N, n), dtype=torch.float32, device=dev) if need_x0 else None
    u_k = u      # a shared input series: g_u[T,1,m] is summed over the systems inside the kernel
    gu = torch.empty((T, 1 if shared_u else N, m), dtype=torch.float32, device=dev) if need_u else None
    flags = ((SHARED_THETA if shared_th else 0) | (SHARED_U if shared_u else 0) | (DISCRETE if discrete else 0))
    if time_chunks is None:
        time_chunks = -1 if (not discrete and 0 < N <= 4096 and T >= 512) else 0
    chunked = time_chunks != 0 and not discrete and not (flags & SHARED_U) and N > 0
    with torch.cuda.device(dev):
        if chunked:
            chunk_len = 0 if time_chunks < 0 else -(-T // int(time_chunks))
            ws = _ws(lib().dynax_lti_vjp_chunked_workspace(N, T, n, m, p, chunk_len), dev)
            check(lib().dynax_lti_vjp_chunked_f32(_ptr(A), _ptr(B), _ptr(C), _ptr(D), _ptr(x0), _ptr(u_k), _ptr(g_xsol),
                                                  _ptr(g_ysol), _ptr(gm[0]), _ptr(gm[1]), _ptr(gm[2]), _ptr(gm[3]),
                                                  _ptr(gx0), _ptr(gu), N, T, n, m, p, float(dt), flags & SHARED_THETA,
                                                  chunk_len, _ptr(ws), ws.numel(), _stream()))
        else:
            ws = _ws(lib().dynax_workspace_bytes(_lib.OP_LTI_VJP, N, T, n, m, p, flags), dev)
            check(lib().dynax_lti_vjp_f32(_ptr(A), _ptr(B), _ptr(C), _ptr(D), _ptr(x0), _ptr(u_k), _ptr(g_xsol),
                                          _ptr(g_ysol), _ptr(gm[0]), _ptr(gm[1]), _ptr(gm[2]), _ptr(gm[3]), _ptr(gx0),
                                          _ptr(gu), N, T, n, m, p, float(dt), flags, _ptr(ws), ws.numel(), _stream()))
    return gm[0], gm[1], gm[2], gm[3], gx0, gu


class _LTIRollout(torch.autograd.Function):
    @staticmethod
    def forward(ctx, A, B, C, D, x0, u, dt, discrete):
        xsol, ysol, _ = lti_forward(A, B, C, D, x0, u, dt, discrete=discrete)
        ctx.save_for_backward(A, B, C, D, x0, u)
        ctx.dt, ctx.discrete = dt, discrete
        return xsol, ysol

    @staticmethod
    def backward(ctx, g_xsol, g_ysol):
        A, B, C, D, x0, u = ctx.saved_tensors
        need = ctx.needs_input_grad
        if g_xsol is None and g_ysol is None:
            return (None,) * 8
        gA, gB, gC, gD, gx0, gu = lti_vjp(A, B, C, D, x0, u, ctx.dt, g_xsol, g_ysol, any(need[:4]), need[4],
                                          need[5], ctx.discrete)
        outs = [g.reshape(M.shape) if (g is not None and nd) else None
                for g, M, nd in zip((gA, gB, gC, gD), (A, B, C, D), need[:4])]
        if gu is not None:
            gu = gu.reshape(u.shape)
        return outs[0], outs[1], outs[2], outs[3], gx0, gu, None, None


def lti_rollout(A, B, C, D, x0, u, dt, discrete=False):
    """Differentiable batched dense-LTI rollout: returns xsol[T,N,n], ysol[T,N,p]."""
    return _LTIRollout.apply(A, B, C, D, x0, u, float(dt), bool(discrete))


# ----------------------------------------------------------------------------------------------
# sampling-based MPC (config 4)
# ----------------------------------------------------------------------------------------------
def rc_mpc_cost(theta, x0, d, u, price, t_low, t_high, dt, cop=2.5, w_energy=100.0, w_comfort=1.0,
                start_time=0.0, want_argmin=True):
    """Cost of N candidate control sequences through one shared 4R3C model (rc.py:131,227-252).

    theta[7], x0[3], d[H,4] (Tao,qCon_i,qRad_e,qRad_i), u[H,N], price/t_low/t_high[24].
    Returns (cost[N], argmin (0-d int64 tensor) or None, min_cost (0-d) or None).
    """
    theta, x0, d, u = (_f32c(t, nm) for t, nm in ((theta, "theta"), (x0, "x0"), (d, "d"), (u, "u")))
    price, t_low, t_high = (_f32c(t, nm) for t, nm in ((price, "price"), (t_low, "t_low"), (t_high, "t_high")))
    if theta.numel() != 7 or x0.numel() != 3:
        raise ValueError("theta must be [7] and x0 [3] (one shared model)")
    if u.dim() != 2 or d.dim() != 2 or d.shape != (u.shape[0], 4):
        raise ValueError(f"u must be [H,N] and d [H,4], got {tuple(u.shape)} and {tuple(d.shape)}")
    if price.numel() != 24 or t_low.numel() != 24 or t_high.numel() != 24:
        raise ValueError("price, t_low, t_high must be hourly tables of 24 entries")
    H, N = u.shape
    dev = u.device
    cost = torch.empty((N,), dtype=torch.float32, device=dev)
    amin = torch.empty((), dtype=torch.int64, device=dev) if want_argmin else None
    cmin = torch.empty((), dtype=torch.float32, device=dev) if want_argmin else None
    ws = _ws(16, dev)
    with torch.cuda.device(dev):
        check(lib().dynax_rc_mpc_cost_f32(_ptr(theta), _ptr(x0), _ptr(d), _ptr(u), _ptr(price), _ptr(t_low),
                                          _ptr(t_high), float(cop), float(w_energy), float(w_comfort),
                                          float(start_time), _ptr(cost), _ptr(amin), _ptr(cmin), N, H, float(dt),
                                          _ptr(ws), ws.numel(), _stream()))
    return cost, amin, cmin


# ----------------------------------------------------------------------------------------------
# general fx/fy: MLP dynamics + RK4 (config 5)
# ----------------------------------------------------------------------------------------------
def node_rk4_forward(W1, b1, W2, b2, W3, b3, x0, u, dt, emit_x=True, emit_y=True, emit_final=False, euler=False,
                     emit_stages=False):
    """Rollout of rhs = W3 tanh(W2 tanh(W1 [x;u] + b1) + b2) + b3, y = x[0] with RK4 (or Euler).

    Weights are row-major [out,in] matrices shared by all systems; x0[N,3], u[T,N,5]|[T,5].
    ``emit_stages``: also return stage_x[T,N,9], the RK4 stage points of every step, which ``node_rk4_vjp`` can
    reuse instead of recomputing them (tcgen05 kernel, RK4 only; None otherwise).
    Hidden widths 32, 64 and 128.
    """
    ws = [_f32c(t, nm) for t, nm in ((W1, "W1"), (b1, "b1"), (W2, "W2"), (b2, "b2"), (W3, "W3"), (b3, "b3"))]
    x0, u = _f32c(x0, "x0"), _f32c(u, "u")
    hidden = ws[1].numel()
    n = ws[5].numel()
    m = ws[0].shape[1] - n
    if ws[0].shape != (hidden, n + m) or ws[2].shape != (hidden, hidden) or ws[3].numel() != hidden or ws[4].shape != (n, hidden):
        raise ValueError("weight shapes must be W1[h,n+m] b1[h] W2[h,h] b2[h] W3[n,h] b3[n]")
    x0, u, N, T, shared_u = _canon_batch(None, x0, u, n, m)
    dev = x0.device
    p = 1
    xsol = torch.empty((T, N, n), dtype=torch.float32, device=dev) if emit_x else None
    ysol = torch.empty((T, N, p), dtype=torch.float32, device=dev) if emit_y else None
    xfin = torch.empty((N, n), dtype=torch.float32, device=dev) if emit_final else None
    flags = (SHARED_U if shared_u else 0) | (_lib.SOLVER_EULER if euler else 0)
    stages = None
    if emit_stages and not euler and os.environ.get("DYNAX_NODE_IMPL", "1")[:1] != "0":
        stages = torch.empty((T, N, 3 * n), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib().dynax_node_rk4_forward_stages_f32(*[_ptr(t) for t in ws], _ptr(x0), _ptr(u), _ptr(xsol), _ptr(ysol),
                                                      _ptr(xfin), _ptr(stages), N, T, n, m, p, hidden, float(dt), flags,
                                                      _stream()))
    if emit_stages:
        return xsol, ysol, xfin, stages
    return xsol, ysol, xfin


# ----------------------------------------------------------------------------------------------
# shared input series + per-system control channel; tabular interpolation (row f-1)
# ----------------------------------------------------------------------------------------------
def _ctrl_args(theta, x0, u_shared, ctrl, ctrl_col):
    x0, u_shared, ctrl = _f32c(x0, "x0"), _f32c(u_shared, "u_shared"), _f32c(ctrl, "ctrl")
    if x0.dim() != 2 or x0.shape[1] != 3:
        raise ValueError(f"x0 must be [N,3], got {tuple(x0.shape)}")
    N = x0.shape[0]
    u_shared = u_shared.reshape(-1, 5)
    T = u_shared.shape[0]
    if ctrl.shape != (T, N):
        raise ValueError(f"ctrl must be [T,N]=({T},{N}), got {tuple(ctrl.shape)}")
    if not 0 <= int(ctrl_col) < 5:
        raise ValueError("ctrl_col must be in [0,5)")
    theta, shared_th = _rc_theta(theta, N)
    return theta, shared_th, x0, u_shared, ctrl, N, T


def rc_forward_ctrl(theta, x0, u_shared, ctrl, dt, ctrl_col=2, emit_x=True, emit_y=True, emit_final=False):
    """RC rollout with inputs u_k = u_shared[k] except column ``ctrl_col`` = ctrl[k, system]
    (dynax/wrapper/envs/rc.py:128-131: shared disturbances + per-system HVAC control)."""
    theta, shared_th, x0, u_shared, ctrl, N, T = _ctrl_args(theta, x0, u_shared, ctrl, ctrl_col)
    dev = x0.device
    xsol = torch.empty((T, N, 3), dtype=torch.float32, device=dev) if emit_x else None
    ysol = torch.empty((T, N, 1), dtype=torch.float32, device=dev) if emit_y else None
    xfin = torch.empty((N, 3), dtype=torch.float32, device=dev) if emit_final else None
    flags = SHARED_THETA if shared_th else 0
    with torch.cuda.device(dev):
        check(lib().dynax_rc_forward_ctrl_f32(_ptr(theta), _ptr(x0), _ptr(u_shared), _ptr(ctrl), int(ctrl_col),
                                              _ptr(xsol), _ptr(ysol), _ptr(xfin), N, T, float(dt), flags, _stream()))
    return xsol, ysol, xfin


def rc_loss_grad_ctrl(theta, x0, u_shared, ctrl, target, dt, ctrl_col=2, lb=None, ub=None, want_gx0=False):
    """Fused MSE loss + gradient with shared disturbances and a per-system control channel."""
    theta, shared_th, x0, u_shared, ctrl, N, T = _ctrl_args(theta, x0, u_shared, ctrl, ctrl_col)
    target = _f32c(target, "target")
    if target.dim() == 2 and target.shape == (T, N) and N != 1:
        shared_tg = False
    elif target.numel() == T:
        shared_tg, target = N != 1, target.reshape(T)
    else:
        raise ValueError(f"target must be [T,N] or [T], got {tuple(target.shape)}")
    dev = x0.device
    if (lb is None) != (ub is None):
        raise ValueError("lb and ub must be given together")
    if lb is not None:
        lb = torch.as_tensor(lb, dtype=torch.float32, device=dev).reshape(7).contiguous()
        ub = torch.as_tensor(ub, dtype=torch.float32, device=dev).reshape(7).contiguous()
    nout = 1 if shared_th else N
    loss = torch.empty((nout,), dtype=torch.float32, device=dev)
    grad = torch.empty((nout, 7), dtype=torch.float32, device=dev)
    gx0 = torch.empty((N, 3), dtype=torch.float32, device=dev) if want_gx0 else None
    flags = (SHARED_THETA if shared_th else 0) | (SHARED_TARGET if shared_tg else 0)
    ws = _ws(lib().dynax_workspace_bytes(_lib.OP_RC_LOSS_GRAD, N, T, 3, 5, 1, flags), dev)
    with torch.cuda.device(dev):
        check(lib().dynax_rc_loss_grad_ctrl_f32(_ptr(theta), _ptr(x0), _ptr(u_shared), _ptr(ctrl), int(ctrl_col),
                                                _ptr(target), _ptr(lb), _ptr(ub), _ptr(loss), _ptr(grad), _ptr(gx0),
                                                N, T, float(dt), flags, _ptr(ws), ws.numel(), _stream()))
    return loss, (grad.reshape(7) if shared_th else grad), gx0


def tabular_interp(ts, xs, at, mode="linear"):
    """``Tabular(ts, xs, mode)(at)`` (dynax/agents/tabular.py:33-44): rows xs[n_rows,n_cols] at times ts,
    evaluated at at[...] -> out[..., n_cols].  mode 'linear' or 'constant' (= nearest row)."""
    ts, xs, at = _f32c(ts, "ts"), _f32c(xs, "xs"), _f32c(at, "at")
    if ts.dim() != 1:
        raise AssertionError("rank of ts has to be 1 !!!")          # interpolate.py:15
    if xs.dim() != 2:
        raise AssertionError("rank of xs has to be 2 !!!")          # interpolate.py:16
    if xs.shape[0] != ts.shape[0]:
        raise ValueError("ts and xs must have the same number of rows")
    if mode not in ("linear", "constant"):
        raise ValueError("Interpolation mode in Tabular has to be either 'linear' or 'constant' !!!")   # tabular.py:39
    flat = at.reshape(-1)
    out = torch.empty((flat.numel(), xs.shape[1]), dtype=torch.float32, device=xs.device)
    with torch.cuda.device(xs.device):
        check(lib().dynax_tabular_interp_f32(_ptr(ts), _ptr(xs), xs.shape[0], xs.shape[1], _ptr(flat), flat.numel(),
                                             1 if mode == "linear" else 0, _ptr(out), _stream()))
    return out.reshape(tuple(at.shape) + (xs.shape[1],)) if at.dim() != 2 or at.shape[-1] != 1 else out


# ----------------------------------------------------------------------------------------------
# batched RC-env step (row f-2)
# ----------------------------------------------------------------------------------------------
def rc_env_step(theta, x, t, action, ts, dist, price, t_low, t_high, dt, cop=2.5, w_energy=100.0, w_comfort=1.0,
                u_low=-10.0, u_high=0.0, num_actions=101):
    """K consecutive ``RC.step`` calls (dynax/wrapper/envs/rc.py:103-158) for N environments.

    theta[7]|[N,7], x[N,3], t[N], action[K,N]|[N] (int), ts[R], dist[R,4], price/t_low/t_high[24].
    Returns dict(x_next[N,3], t_next[N], obs[K,N,5], reward[K,N], terminated[K,N], cost[K,N], dTz[K,N]).
    """
    x, t, ts, dist = _f32c(x, "x"), _f32c(t, "t"), _f32c(ts, "ts"), _f32c(dist, "dist")
    price, t_low, t_high = _f32c(price, "price"), _f32c(t_low, "t_low"), _f32c(t_high, "t_high")
    N = x.shape[0]
    theta, shared_th = _rc_theta(theta, N)
    action = action.to(torch.int32)
    if action.dim() == 1:
        action = action[None]
    action = action.contiguous()
    if not action.is_cuda or action.shape[1] != N or t.numel() != N or dist.shape != (ts.numel(), 4):
        raise ValueError("action must be CUDA [K,N], t [N], dist [R,4] with R = len(ts)")
    K = action.shape[0]
    dev = x.device
    mk = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
    out = dict(x_next=mk(N, 3), t_next=mk(N), obs=mk(K, N, 5), reward=mk(K, N), terminated=mk(K, N), cost=mk(K, N), dTz=mk(K, N))
    with torch.cuda.device(dev):
        check(lib().dynax_rc_env_step_f32(_ptr(theta), _ptr(x), _ptr(t), _ptr(action), _ptr(ts), _ptr(dist), ts.numel(),
                                          _ptr(price), _ptr(t_low), _ptr(t_high), float(cop), float(w_energy),
                                          float(w_comfort), float(u_low), float(u_high), int(num_actions),
                                          _ptr(out["x_next"]), _ptr(out["t_next"]), _ptr(out["obs"]), _ptr(out["reward"]),
                                          _ptr(out["terminated"]), _ptr(out["cost"]), _ptr(out["dTz"]), N, K, float(dt),
                                          SHARED_THETA if shared_th else 0, _stream()))
    return out


# ----------------------------------------------------------------------------------------------
# on-device calibration loop pieces (rows f-3 / f-4)
# ----------------------------------------------------------------------------------------------
def lamb_update_(theta, grad, m, v, step, lr=1e-3, b1=0.9, b2=0.999, eps=1e-6):
    """In-place optax.lamb(lr) update (examples/3-inverse.py:113-122) of every scalar in ``theta``."""
    for t, nm in ((theta, "theta"), (grad, "grad"), (m, "m"), (v, "v")):
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise ValueError(f"{nm} must be a contiguous float32 CUDA tensor")
        if t.numel() != theta.numel():
            raise ValueError("theta, grad, m, v must have the same number of elements")
    with torch.cuda.device(theta.device):
        check(lib().dynax_lamb_update_f32(_ptr(theta), _ptr(grad), _ptr(m), _ptr(v), theta.numel(), int(step), float(lr),
                                          float(b1), float(b2), float(eps), _stream()))
    return theta


def lamb_update_dev_(theta, grad, m, v, step_counter, lr=1e-3, b1=0.9, b2=0.999, eps=1e-6):
    """LAMB update whose epoch number lives in ``step_counter`` (int64 CUDA tensor, incremented by the
    call) -- the graph-capturable form of :func:`lamb_update_`."""
    if not (step_counter.is_cuda and step_counter.dtype == torch.int64 and step_counter.numel() == 1):
        raise ValueError("step_counter must be a 1-element int64 CUDA tensor")
    for t, nm in ((theta, "theta"), (grad, "grad"), (m, "m"), (v, "v")):
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise ValueError(f"{nm} must be a contiguous float32 CUDA tensor")
    with torch.cuda.device(theta.device):
        check(lib().dynax_lamb_update_dev_f32(_ptr(theta), _ptr(grad), _ptr(m), _ptr(v), theta.numel(),
                                              _ptr(step_counter), float(lr), float(b1), float(b2), float(eps), _stream()))
    return theta


def nonfinite_rows(a):
    """Per-system non-finite guard for any per-system output a[N, ...] (loss, x_final, gradients): returns
    (flags int32[N], count 0-d int32) -- flags[i] = 1 if system i holds a NaN/Inf.  Euler can blow up inside the
    parameter box of examples/3-inverse.py:66-85; check after a rollout instead of discovering an inf loss later."""
    a = _f32c(a, "a")
    rows = a.shape[0] if a.dim() > 0 else 1
    width = max(1, a.numel() // max(rows, 1))
    flags = torch.empty((rows,), dtype=torch.int32, device=a.device)
    count = torch.empty((), dtype=torch.int32, device=a.device)
    with torch.cuda.device(a.device):
        check(lib().dynax_nonfinite_rows_f32(_ptr(a), rows, width, _ptr(flags), _ptr(count), _stream()))
    return flags, count


def resample_mean(x, factor):
    """``data.groupby(index // factor).mean()`` (examples/1-forward.py:26-27) for x[rows, cols] on the device."""
    x = _f32c(x, "x")
    if x.dim() != 2:
        raise ValueError("x must be [rows, cols]")
    rows, cols = x.shape
    out = torch.empty((-(-rows // int(factor)), cols), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(lib().dynax_resample_mean_f32(_ptr(x), _ptr(out), rows, cols, int(factor), _stream()))
    return out


def node_rk4_vjp(W1, b1, W2, b2, W3, b3, x0, u, xsol, dt, g_xsol=None, g_ysol=None, need_u=True, euler=False,
                 stage_x=None):
    """VJP of node_rk4_forward: returns (gW1, gb1, gW2, gb2, gW3, gb3, g_x0, g_u).
    ``stage_x``: the stage points ``node_rk4_forward(..., emit_stages=True)`` returned (saves three forward
    evaluations per step), or None."""
    ws_ = [_f32c(t, nm) for t, nm in ((W1, "W1"), (b1, "b1"), (W2, "W2"), (b2, "b2"), (W3, "W3"), (b3, "b3"))]
    x0, u, xsol = _f32c(x0, "x0"), _f32c(u, "u"), _f32c(xsol, "xsol")
    hidden, n = ws_[1].numel(), ws_[5].numel()
    m = ws_[0].shape[1] - n
    # same shape rules as the forward (a shared series is u[T,m] or u[T,1,m]: DYNAX_SHARED_U, g_u is then [T,1,m], summed
    # over the systems inside the kernel)
    x0, u, N, T, shared_u = _canon_batch(None, x0, u, n, m)
    dev = x0.device
    if g_xsol is not None:
        g_xsol = _f32c(g_xsol, "g_xsol").reshape(T, N, n)
    if g_ysol is not None:
        g_ysol = _f32c(g_ysol, "g_ysol").reshape(T, N, 1)
    gw = [torch.empty_like(t) for t in ws_]
    gx0 = torch.empty((N, n), dtype=torch.float32, device=dev)
    gu = torch.empty((T, 1 if shared_u else N, m), dtype=torch.float32, device=dev) if need_u else None
    wsb = _ws(lib().dynax_node_rk4_vjp_workspace(), dev)
    with torch.cuda.device(dev):
        if stage_x is not None:
            stage_x = _f32c(stage_x, "stage_x").reshape(T, N, 3 * n)
        check(lib().dynax_node_rk4_vjp_stages_f32(*[_ptr(t) for t in ws_], _ptr(x0), _ptr(u), _ptr(xsol), _ptr(stage_x),
                                                  _ptr(g_xsol), _ptr(g_ysol), *[_ptr(t) for t in gw], _ptr(gx0),
                                                  _ptr(gu), N, T, n, m, 1, hidden, float(dt),
                                                  (_lib.SOLVER_EULER if euler else 0) | (SHARED_U if shared_u else 0),
                                                  _ptr(wsb), wsb.numel(), _stream()))
    return (*gw, gx0, gu)


class _NodeRollout(torch.autograd.Function):
    @staticmethod
    def forward(ctx, W1, b1, W2, b2, W3, b3, x0, u, dt, euler):
        # the stage points cost 36 B per trajectory-step; they are kept only when a gradient can be asked for
        want = any(t.requires_grad for t in (W1, b1, W2, b2, W3, b3, x0, u))
        out = node_rk4_forward(W1, b1, W2, b2, W3, b3, x0, u, dt, euler=euler, emit_stages=want)
        xsol, ysol, stages = out[0], out[1], (out[3] if want else None)
        ctx.save_for_backward(W1, b1, W2, b2, W3, b3, x0, u, xsol, *([stages] if stages is not None else []))
        ctx.dt, ctx.euler = dt, euler
        return xsol, ysol

    @staticmethod
    def backward(ctx, g_xsol, g_ysol):
        saved = ctx.saved_tensors
        W, (x0, u, xsol), stages = saved[:6], saved[6:9], (saved[9] if len(saved) > 9 else None)
        if g_xsol is None and g_ysol is None:
            return (None,) * 10
        out = node_rk4_vjp(*W, x0, u, xsol, ctx.dt, g_xsol, g_ysol, need_u=ctx.needs_input_grad[7], euler=ctx.euler,
                           stage_x=stages)
        gu = out[7]
        if gu is not None:
            gu = gu.reshape(u.shape)        # [T,1,m] (summed over systems by node_rk4_vjp) -> the caller's [T,m] / [T,1,m]
        return (*out[:6], out[6], gu, None, None)


def node_rollout(W1, b1, W2, b2, W3, b3, x0, u, dt, euler=False):
    """Differentiable MLP-dynamics rollout (RK4 or Euler): xsol[T,N,3], ysol[T,N,1]."""
    return _NodeRollout.apply(W1, b1, W2, b2, W3, b3, x0, u, float(dt), bool(euler))
