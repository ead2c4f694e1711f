"""numba cross-check of the CPU baseline (TEST INFRASTRUCTURE ONLY -- like everything under oracle/).

BASELINE.md section 5 asks for a second, independently written CPU implementation of the reference path next to the
C/OpenMP port (oracle/dynax_oracle.c), so that the CPU figure bench.py reports is not an artefact of one port:
``@njit(parallel=True)`` over the systems, a scalar fp32 time loop per system -- what a NumPy/numba user would write.

Same algorithm as the C port and as XLA's transpose of the scan: forward sweep storing the per-step states
(/root/reference/dynax/simulators/simulator.py:36-43 with the 4R3C matrices of dynax/models/RC.py:203-209,224-229),
reverse sweep accumulating dA, dB, chain rule to theta, MSE loss of
examples/3-inverse.py:96-101.  All arithmetic in fp32 (the reference's default dtype); no box penalty (bench.py's
baseline sample passes lb = ub = None).
"""
from __future__ import annotations

import numpy as np
from numba import get_num_threads, njit, prange, set_num_threads

f32 = np.float32


@njit(parallel=True, cache=True, fastmath=False)
def _rc_loss_grad_f32(theta, x0, u, target, dt, loss, grad):
    T, N = target.shape
    one = f32(1.0)
    for i in prange(N):
        Cai, Cwe, Cwi = theta[i, 0], theta[i, 1], theta[i, 2]
        Re, Ri, Rw, Rg = theta[i, 3], theta[i, 4], theta[i, 5], theta[i, 6]
        A00 = -one / Cai * (one / Rg + one / Ri)     # RC.py:203
        A02 = one / (Cai * Ri)                        # RC.py:204
        A11 = -one / Cwe * (one / Re + one / Rw)     # RC.py:205
        A12 = one / (Cwe * Rw)                        # RC.py:206
        A20 = one / (Cwi * Ri)                        # RC.py:207
        A21 = one / (Cwi * Rw)                        # RC.py:208
        A22 = -one / Cwi * (one / Rw + one / Ri)     # RC.py:209
        B00 = one / (Cai * Rg)                        # RC.py:224
        B01 = one / Cai                               # RC.py:225-226
        B10 = one / (Cwe * Re)                        # RC.py:227
        B13 = one / Cwe                               # RC.py:228
        B24 = one / Cwi                               # RC.py:229
        xs = np.empty((T, 3), dtype=np.float32)       # the residuals XLA's scan transpose stores
        a, b, c = x0[i, 0], x0[i, 1], x0[i, 2]
        sq = f32(0.0)
        for k in range(T):
            u0, u1, u2, u3, u4 = u[k, i, 0], u[k, i, 1], u[k, i, 2], u[k, i, 3], u[k, i, 4]
            r0 = (A00 * a + A02 * c) + (B00 * u0 + B01 * u1 + B01 * u2)   # fxx(x) + fxu(u)
            r1 = (A11 * b + A12 * c) + (B10 * u0 + B13 * u3)
            r2 = (A20 * a + A21 * b + A22 * c) + B24 * u4
            xs[k, 0] = a; xs[k, 1] = b; xs[k, 2] = c
            res = a - target[k, i]                                         # y_k = x_k[0], pre-update
            sq += res * res
            a = a + dt * r0; b = b + dt * r1; c = c + dt * r2
        l0 = f32(0.0); l1 = f32(0.0); l2 = f32(0.0)
        gA00 = f32(0.0); gA02 = f32(0.0); gA11 = f32(0.0); gA12 = f32(0.0); gA20 = f32(0.0); gA21 = f32(0.0); gA22 = f32(0.0)
        gB00 = f32(0.0); gB01 = f32(0.0); gB02 = f32(0.0); gB10 = f32(0.0); gB13 = f32(0.0); gB24 = f32(0.0)
        c2T = f32(2.0) / f32(T)
        for kk in range(T):
            k = T - 1 - kk
            X0, X1, X2 = xs[k, 0], xs[k, 1], xs[k, 2]
            gy = c2T * (X0 - target[k, i])
            s0 = dt * l0; s1 = dt * l1; s2 = dt * l2                       # cotangent of rhs
            gA00 += s0 * X0; gA02 += s0 * X2
            gA11 += s1 * X1; gA12 += s1 * X2
            gA20 += s2 * X0; gA21 += s2 * X1; gA22 += s2 * X2
            gB00 += s0 * u[k, i, 0]; gB01 += s0 * u[k, i, 1]; gB02 += s0 * u[k, i, 2]
            gB10 += s1 * u[k, i, 0]; gB13 += s1 * u[k, i, 3]
            gB24 += s2 * u[k, i, 4]
            n0 = l0 + (A00 * s0 + A20 * s2) + gy                           # a + A^T s + C^T gy
            n1 = l1 + (A11 * s1 + A21 * s2)
            n2 = l2 + (A02 * s0 + A12 * s1 + A22 * s2)
            l0 = n0; l1 = n1; l2 = n2
        iCai = one / Cai; iCwe = one / Cwe; iCwi = one / Cwi
        iRe = one / Re; iRi = one / Ri; iRw = one / Rw; iRg = one / Rg
        grad[i, 0] = -iCai * (gA00 * A00 + gA02 * A02 + gB00 * B00 + (gB01 + gB02) * B01)
        grad[i, 1] = -iCwe * (gA11 * A11 + gA12 * A12 + gB10 * B10 + gB13 * B13)
        grad[i, 2] = -iCwi * (gA20 * A20 + gA21 * A21 + gA22 * A22 + gB24 * B24)
        grad[i, 3] = iCwe * iRe * iRe * (gA11 - gB10)
        grad[i, 4] = iRi * iRi * (iCai * (gA00 - gA02) + iCwi * (gA22 - gA20))
        grad[i, 5] = iRw * iRw * (iCwe * (gA11 - gA12) + iCwi * (gA22 - gA21))
        grad[i, 6] = iCai * iRg * iRg * (gA00 - gB00)
        loss[i] = sq / f32(T)


def rc_loss_grad(theta, x0, u, target, dt):
    """theta[N,7], x0[N,3], u[T,N,5], target[T,N] (fp32) -> loss[N], grad[N,7] (fp32, all-fp32 BPTT)."""
    theta, x0, u, target = (np.ascontiguousarray(v, dtype=np.float32) for v in (theta, x0, u, target))
    N = theta.shape[0]
    loss = np.empty((N,), np.float32)
    grad = np.empty((N, 7), np.float32)
    _rc_loss_grad_f32(theta, x0, u, target, np.float32(dt), loss, grad)
    return loss, grad


def use_all_cores():
    import os
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        set_num_threads(n)
    except ValueError:      # more than NUMBA_NUM_THREADS allows: keep numba's own default
        pass
    return get_num_threads()
