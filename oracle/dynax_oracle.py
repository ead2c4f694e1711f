"""CPU restatement (NumPy) of dynax's rollout hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  Nothing under ``dynax_b200/``
imports it: the product path is the CUDA library and fails loudly without it.

Parity status
-------------
* Forward rollout (RC and dense LTI): **pinned** by the reference's own golden
  vectors -- ``tests/test_simulator.py:24-36`` (golden #1) and
  ``tests/test_gymwrapper_basics.py:7-12`` (golden #2) -- see
  ``tests/golden/`` and ``tests/test_oracle_golden.py``.
* Gradient / loss / MPC cost / RK4-MLP: **parity unpinned by the reference** (it
  holds no gradient, RK4 or cost test).  The hand adjoint here is pinned instead
  against torch autograd (fp64) over this same restated loop
  (``tests/test_oracle_adjoint.py``).
* The reference is JAX/Flax; neither is installed in this image, so the
  reference itself cannot be run here (``import jax`` -> ModuleNotFoundError).

Each function cites the reference file:line it follows (paths relative to
``/root/reference``).
"""
from __future__ import annotations

import numpy as np

# Parameter order used everywhere in this repo (RC.py:176-182 declaration order).
RC_PARAM_NAMES = ("Cai", "Cwe", "Cwi", "Re", "Ri", "Rw", "Rg")

# examples/1-forward.py:41-43 == dynax/wrapper/envs/rc.py:25-27
RC_NOMINAL = np.array([10384.31640625, 499089.09375, 1321535.125,
                       1.5348844528198242, 0.5000327825546265, 1.000040054321289,
                       20.119935989379883], dtype=np.float64)

# examples/3-inverse.py:66-85
RC_LB = np.array([3.0e3, 5.0e4, 5.0e5, 1.0, 0.5, 1.0, 5.0], dtype=np.float64)
RC_UB = np.array([3.0e4, 5.0e5, 5.0e6, 1.0e1, 5.0, 1.0e1, 5.0e1], dtype=np.float64)


# --------------------------------------------------------------------------------------
# A3: theta -> A, B, C, D          dynax/models/RC.py:201-210, 222-230, 236-239, 245-246
# --------------------------------------------------------------------------------------
def rc_abcd(theta, dtype=np.float32):
    """theta[..., 7] -> A[...,3,3], B[...,3,5], C[...,1,3], D[...,1,5].

    Expressions are written exactly as in RC.py so fp32 rounding of the coefficients
    matches the reference (e.g. ``-1/Cai*(1/Rg+1/Ri)`` is ``((-1)/Cai)*(1/Rg+1/Ri)``).
    """
    th = np.asarray(theta, dtype=dtype)
    Cai, Cwe, Cwi, Re, Ri, Rw, Rg = (th[..., i] for i in range(7))
    one = dtype(1)
    lead = th.shape[:-1]
    A = np.zeros(lead + (3, 3), dtype=dtype)
    A[..., 0, 0] = -one / Cai * (one / Rg + one / Ri)      # RC.py:203
    A[..., 0, 2] = one / (Cai * Ri)                         # RC.py:204
    A[..., 1, 1] = -one / Cwe * (one / Re + one / Rw)      # RC.py:205
    A[..., 1, 2] = one / (Cwe * Rw)                         # RC.py:206
    A[..., 2, 0] = one / (Cwi * Ri)                         # RC.py:207
    A[..., 2, 1] = one / (Cwi * Rw)                         # RC.py:208
    A[..., 2, 2] = -one / Cwi * (one / Rw + one / Ri)      # RC.py:209
    B = np.zeros(lead + (3, 5), dtype=dtype)
    B[..., 0, 0] = one / (Cai * Rg)                         # RC.py:224
    B[..., 0, 1] = one / Cai                                # RC.py:225
    B[..., 0, 2] = one / Cai                                # RC.py:226
    B[..., 1, 0] = one / (Cwe * Re)                         # RC.py:227
    B[..., 1, 3] = one / Cwe                                # RC.py:228
    B[..., 2, 4] = one / Cwi                                # RC.py:229
    C = np.zeros(lead + (1, 3), dtype=dtype)
    C[..., 0, 0] = one                                      # RC.py:237-238
    D = np.zeros(lead + (1, 5), dtype=dtype)                # RC.py:246
    return A, B, C, D


# --------------------------------------------------------------------------------------
# A1 + A2: the scan body             dynax/simulators/simulator.py:36-43
#                                    dynax/core/base_block_state_space.py:59-61, 68-71
# --------------------------------------------------------------------------------------
def _bmv(M, v):
    """Batched mat-vec: M[N,r,c] (or [1,r,c]) times v[N,c] -> [N,r]."""
    return np.einsum("nrc,nc->nr", M, v)


def lti_rollout(A, B, C, D, x0, u, dt, dtype=np.float32, discrete=False):
    """Batched restatement of ``DifferentiableSimulator`` over a linear block SSM.

    A[N|1,n,n], B[N|1,n,m], C[N|1,p,n], D[N|1,p,m], x0[N,n], u[T,N|1,m].
    Returns xsol[T,N,n] (x_{k+1}, post-update) and ysol[T,N,p] (y(x_k,u_k), pre-update):
    simulator.py:38-43.  ``discrete=True`` replaces the Euler residual update by
    ``x = rhs`` (tests/test_forward_simulation.py:30-36 steps a discrete model directly).
    """
    A, B, C, D = (np.asarray(M, dtype=dtype) for M in (A, B, C, D))
    x = np.array(x0, dtype=dtype, copy=True)
    u = np.asarray(u, dtype=dtype)
    T = u.shape[0]
    N, n = x.shape
    p = C.shape[-2]
    dt = dtype(dt)
    xsol = np.empty((T, N, n), dtype=dtype)
    ysol = np.empty((T, N, p), dtype=dtype)
    for k in range(T):
        uk = np.broadcast_to(u[k], (N, u.shape[-1]))
        rhs = _bmv(A, x) + _bmv(B, uk)          # base_block_state_space.py:61  fxx(x)+fxu(u)
        y = _bmv(C, x) + _bmv(D, uk)            # base_block_state_space.py:69  fyx(x)+fyu(u)
        if discrete:
            x = rhs
        else:
            x = x + dt * rhs                    # simulator.py:40   states += dt*rhs
        xsol[k] = x                             # simulator.py:43   emit post-update state
        ysol[k] = y                             #                   and pre-update output
    return xsol, ysol


def rc_rollout(theta, x0, u, dt, dtype=np.float32):
    """theta[N|1,7], x0[N,3], u[T,N|1,5] -> xsol[T,N,3], ysol[T,N,1]."""
    theta = np.atleast_2d(np.asarray(theta))
    A, B, C, D = rc_abcd(theta, dtype)
    return lti_rollout(A, B, C, D, x0, u, dt, dtype)


def rc_rollout_single(theta, x0, u, dt, dtype=np.float32):
    """Unbatched signature of the reference: theta[7], x0[3], u[T,5] -> xsol[T,3], ysol[T,1]."""
    xs, ys = rc_rollout(np.asarray(theta)[None], np.asarray(x0)[None], np.asarray(u)[:, None, :], dt, dtype)
    return xs[:, 0], ys[:, 0]


# --------------------------------------------------------------------------------------
# A6: loss                           examples/3-inverse.py:96-110
# --------------------------------------------------------------------------------------
def bound_penalty(theta, lb, ub, dtype=np.float64):
    """sum_j relu(theta_j-ub_j)/ub_j + relu(lb_j-theta_j)/ub_j  (normaliser is ub for both)."""
    th = np.asarray(theta, dtype=dtype)
    lb = np.asarray(lb, dtype=dtype)
    ub = np.asarray(ub, dtype=dtype)
    over = np.maximum(th - ub, 0) / ub          # 3-inverse.py:104
    under = np.maximum(lb - th, 0) / ub         # 3-inverse.py:105
    return (over + under).sum(-1)


def bound_penalty_grad(theta, lb, ub, dtype=np.float64):
    th = np.asarray(theta, dtype=dtype)
    lb = np.asarray(lb, dtype=dtype)
    ub = np.asarray(ub, dtype=dtype)
    return (th > ub).astype(dtype) / ub - (th < lb).astype(dtype) / ub


def rc_loss(theta, x0, u, target, dt, lb=None, ub=None, dtype=np.float32):
    """loss[N] = mean_k (y_k - target_k)^2 (+ bound penalty).  target[T,N|1]."""
    _, ys = rc_rollout(theta, x0, u, dt, dtype)
    r = ys[..., 0] - np.asarray(target, dtype=dtype)
    loss = np.mean(r * r, axis=0, dtype=dtype)              # 3-inverse.py:100
    if lb is not None:
        loss = loss + bound_penalty(np.atleast_2d(theta), lb, ub, dtype)
    return loss


# --------------------------------------------------------------------------------------
# A5: reverse-mode gradient (discrete adjoint == what jax.value_and_grad derives through
#     lax.scan for this body; examples/3-inverse.py:112).  Derived in SURVEY.md section 8(a) A5.
# --------------------------------------------------------------------------------------
def lti_vjp(A, B, C, D, x0, u, dt, g_xsol, g_ysol, dtype=np.float64, discrete=False):
    """General-cotangent VJP of ``lti_rollout``.

    g_xsol[T,N,n] / g_ysol[T,N,p] are dL/dxsol and dL/dysol (either may be None == 0).
    Returns dict of per-system gradients gA[N,n,n], gB[N,n,m], gC[N,p,n], gD[N,p,m],
    gx0[N,n], gu[T,N,m].  (If A.. or u were shared the caller sums over N.)
    """
    A, B, C, D = (np.asarray(M, dtype=dtype) for M in (A, B, C, D))
    u = np.asarray(u, dtype=dtype)
    T = u.shape[0]
    x0 = np.asarray(x0, dtype=dtype)
    N, n = x0.shape
    m = u.shape[-1]
    p = C.shape[-2]
    dt = dtype(dt)
    # forward sweep storing x_k (the residuals XLA would store)
    xs = np.empty((T + 1, N, n), dtype=dtype)
    xs[0] = x0
    for k in range(T):
        uk = np.broadcast_to(u[k], (N, m))
        rhs = _bmv(A, xs[k]) + _bmv(B, uk)
        xs[k + 1] = rhs if discrete else xs[k] + dt * rhs
    gA = np.zeros((N, n, n), dtype=dtype)
    gB = np.zeros((N, n, m), dtype=dtype)
    gC = np.zeros((N, p, n), dtype=dtype)
    gD = np.zeros((N, p, m), dtype=dtype)
    gu = np.zeros((T, N, m), dtype=dtype)
    a = np.zeros((N, n), dtype=dtype)            # adjoint of x_{k+1}
    AT = np.swapaxes(A, -1, -2)
    BT = np.swapaxes(B, -1, -2)
    CT = np.swapaxes(C, -1, -2)
    DT = np.swapaxes(D, -1, -2)
    for k in range(T - 1, -1, -1):
        if g_xsol is not None:
            a = a + np.asarray(g_xsol[k], dtype=dtype)      # xsol[k] = x_{k+1}
        gy = np.zeros((N, p), dtype=dtype) if g_ysol is None else np.asarray(g_ysol[k], dtype=dtype)
        uk = np.broadcast_to(u[k], (N, m))
        s = a if discrete else dt * a                       # cotangent of rhs
        gA += s[:, :, None] * xs[k][:, None, :]
        gB += s[:, :, None] * uk[:, None, :]
        gC += gy[:, :, None] * xs[k][:, None, :]
        gD += gy[:, :, None] * uk[:, None, :]
        gu[k] = _bmv(BT, s) + _bmv(DT, gy)
        a = (0 if discrete else a) + _bmv(AT, s) + _bmv(CT, gy)
    return dict(gA=gA, gB=gB, gC=gC, gD=gD, gx0=a, gu=gu)


def rc_chain(theta, gA, gB, dtype=np.float64):
    """dL/dtheta[N,7] from dL/dA[N,3,3], dL/dB[N,3,5] for the 4R3C parameterisation
    (derivative of the expressions at RC.py:203-209, 224-229)."""
    th = np.atleast_2d(np.asarray(theta, dtype=dtype))
    Cai, Cwe, Cwi, Re, Ri, Rw, Rg = (th[:, i] for i in range(7))
    iCai, iCwe, iCwi, iRe, iRi, iRw, iRg = (1 / v for v in (Cai, Cwe, Cwi, Re, Ri, Rw, Rg))
    A00 = -iCai * (iRg + iRi); A02 = iCai * iRi
    A11 = -iCwe * (iRe + iRw); A12 = iCwe * iRw
    A20 = iCwi * iRi; A21 = iCwi * iRw; A22 = -iCwi * (iRw + iRi)
    B00 = iCai * iRg; B01 = iCai; B10 = iCwe * iRe; B13 = iCwe; B24 = iCwi
    g = np.zeros_like(th)
    g[:, 0] = -iCai * (gA[:, 0, 0] * A00 + gA[:, 0, 2] * A02 + gB[:, 0, 0] * B00 + (gB[:, 0, 1] + gB[:, 0, 2]) * B01)
    g[:, 1] = -iCwe * (gA[:, 1, 1] * A11 + gA[:, 1, 2] * A12 + gB[:, 1, 0] * B10 + gB[:, 1, 3] * B13)
    g[:, 2] = -iCwi * (gA[:, 2, 0] * A20 + gA[:, 2, 1] * A21 + gA[:, 2, 2] * A22 + gB[:, 2, 4] * B24)
    g[:, 3] = iCwe * iRe * iRe * (gA[:, 1, 1] - gB[:, 1, 0])
    g[:, 4] = iRi * iRi * (iCai * (gA[:, 0, 0] - gA[:, 0, 2]) + iCwi * (gA[:, 2, 2] - gA[:, 2, 0]))
    g[:, 5] = iRw * iRw * (iCwe * (gA[:, 1, 1] - gA[:, 1, 2]) + iCwi * (gA[:, 2, 2] - gA[:, 2, 1]))
    g[:, 6] = iCai * iRg * iRg * (gA[:, 0, 0] - gB[:, 0, 0])
    return g


def rc_vjp(theta, x0, u, dt, g_xsol, g_ysol, dtype=np.float64):
    """VJP of rc_rollout: returns gtheta[N,7], gx0[N,3], gu[T,N,5]."""
    theta = np.atleast_2d(np.asarray(theta, dtype=dtype))
    A, B, C, D = rc_abcd(theta, dtype)
    r = lti_vjp(A, B, C, D, x0, u, dt, g_xsol, g_ysol, dtype)
    return rc_chain(theta, r["gA"], r["gB"], dtype), r["gx0"], r["gu"]


def rc_loss_grad(theta, x0, u, target, dt, lb=None, ub=None, dtype=np.float64):
    """loss[N], grad[N,7] of examples/3-inverse.py:96-112 for a batch of candidates.

    theta[N,7], x0[N,3], u[T,N|1,5], target[T,N|1].
    """
    theta = np.atleast_2d(np.asarray(theta, dtype=dtype))
    x0 = np.asarray(x0, dtype=dtype)
    N = x0.shape[0]
    _, ys = rc_rollout(theta, x0, u, dt, dtype)
    T = ys.shape[0]
    r = ys[..., 0] - np.asarray(target, dtype=dtype)
    loss = np.mean(r * r, axis=0)
    gy = (2.0 / T) * r                                      # d mean(r^2) / dy_k
    g, _, _ = rc_vjp(theta, x0, u, dt, None, gy[..., None].astype(dtype), dtype)
    if lb is not None:
        loss = loss + bound_penalty(theta, lb, ub, dtype)
        g = g + bound_penalty_grad(theta, lb, ub, dtype)
    return loss.astype(dtype), g.astype(dtype)


# --------------------------------------------------------------------------------------
# Synthetic workloads; shared by tests and bench so that any system
# of a large GPU run can be regenerated on the host.
# --------------------------------------------------------------------------------------
def spectral_radius_euler(theta, dt):
    A, _, _, _ = rc_abcd(np.atleast_2d(theta), np.float64)
    M = np.eye(3)[None] + dt * A
    return np.abs(np.linalg.eigvals(M)).max(-1)


def synth_theta(N, seed, spread=0.2, dt=3600.0):
    """theta_i = nominal * exp(U(-spread, spread)), Euler-stable at dt (asserted)."""
    rng = np.random.default_rng(seed)
    th = RC_NOMINAL[None] * np.exp(rng.uniform(-spread, spread, size=(N, 7)))
    assert spectral_radius_euler(th, dt).max() < 1.0, "unstable synthetic theta"
    return th.astype(np.float32)


def synth_x0(N, seed):
    rng = np.random.default_rng(seed + 1)
    return (np.array([20.0, 30.0, 26.0])[None] + rng.uniform(-1, 1, size=(N, 3))).astype(np.float32)


def synth_u(T, N, seed):
    """u[T,N,5] fp32: [Tao, qCon_i, qHVAC, qRad_e, qRad_i] in degC / kW."""
    rng = np.random.default_rng(seed + 2)
    k = np.arange(T, dtype=np.float64)[:, None]
    phi = rng.uniform(0, 2 * np.pi, size=(1, N))
    u = np.empty((T, N, 5), dtype=np.float32)
    u[..., 0] = 15 + 10 * np.sin(2 * np.pi * k / 8760) + 6 * np.sin(2 * np.pi * k / 24 + phi) + rng.normal(size=(T, N))
    u[..., 1] = rng.uniform(0, 1, size=(T, N))
    u[..., 2] = -rng.uniform(0, 5, size=(T, N))
    u[..., 3] = rng.uniform(0, 2, size=(T, N)) * (np.sin(2 * np.pi * k / 24) > 0)
    u[..., 4] = rng.uniform(0, 1, size=(T, N))
    return u


# --------------------------------------------------------------------------------------
# Conditioning-aware error budgets.  north_star states two tolerances: trajectories within
# rel 1e-5, gradients within rel 1e-4.  A loss/gradient is a functional of the trajectory, so a
# trajectory that is itself only specified to 1e-5 determines them no better than the first-order
# bounds below.  (SURVEY.md section 7 H2 / finding #9: per-element 1e-4 against an fp32 BPTT is
# below fp32 noise whenever the residual y-target is small against |y|.)
# --------------------------------------------------------------------------------------
def rc_output_sensitivity_budget(theta, x0, u, target, dt, traj_rtol=1e-5):
    """Returns (loss_budget[N], grad_budget[N,7]): how far loss_i and dloss_i/dtheta_j can move when
    every y_k is perturbed by up to traj_rtol*|y_k| (first order, fp64):
        loss_budget_i   = (2/T) sum_k |r_k| * traj_rtol*|y_k|
        grad_budget_ij  = (2/T) sum_k |dy_k/dtheta_j| * traj_rtol*|y_k|
    dy/dtheta by forward-mode sensitivities of the same Euler recurrence."""
    th = np.atleast_2d(np.asarray(theta, dtype=np.float64))
    x = np.array(x0, dtype=np.float64)
    u = np.asarray(u, dtype=np.float64)
    tg = np.asarray(target, dtype=np.float64)
    T = u.shape[0]
    N = x.shape[0]
    A, B, _, _ = rc_abcd(th, np.float64)
    dA = np.empty((N, 7, 3, 3)); dB = np.empty((N, 7, 3, 5))
    for j in range(7):
        h = 1e-6 * th[:, j]
        tp, tm = th.copy(), th.copy()
        tp[:, j] += h; tm[:, j] -= h
        Ap, Bp, _, _ = rc_abcd(tp, np.float64)
        Am, Bm, _, _ = rc_abcd(tm, np.float64)
        dA[:, j] = (Ap - Am) / (2 * h)[:, None, None]
        dB[:, j] = (Bp - Bm) / (2 * h)[:, None, None]
    S = np.zeros((N, 3, 7))
    lb = np.zeros(N); gb = np.zeros((N, 7))
    for k in range(T):
        uk = np.broadcast_to(u[k], (N, 5))
        y = x[:, 0]
        r = y - (tg[k] if tg.ndim == 2 else tg[k])
        w = traj_rtol * np.abs(y)
        lb += np.abs(r) * w
        gb += np.abs(S[:, 0, :]) * w[:, None]
        forcing = np.einsum("njab,nb->naj", dA, x) + np.einsum("njab,nb->naj", dB, uk)
        S = S + dt * (np.einsum("nab,nbj->naj", A, S) + forcing)
        x = x + dt * (_bmv(A, x) + _bmv(B, uk))
    return (2.0 / T) * lb, (2.0 / T) * gb


# --------------------------------------------------------------------------------------
# Config 4: sampling-based MPC cost = negated RC-env reward summed over the horizon.
#   dynax/wrapper/envs/rc.py:125-147 (step), :200-222 (obs), :227-252 (reward), :32-42 (schedules)
# Pinned for H=1 by golden #2 (REWARD=-0.05974, tests/test_gymwrapper_basics.py:8).
# --------------------------------------------------------------------------------------
RC_T_LOW = np.array([12.0] * 8 + [22.0] * 10 + [12.0] * 6)          # rc.py:32-33
RC_T_HIGH = np.array([30.0] * 8 + [26.0] * 10 + [30.0] * 6)         # rc.py:34-35
RC_PRICE = np.array([0.02987] * 6 + [0.04667] * 6 + [0.15877] * 7 + [0.04667] * 3 + [0.02987] * 2)  # rc.py:37-42


def mpc_cost(theta, x0, d, u, dt, price=RC_PRICE, t_low=RC_T_LOW, t_high=RC_T_HIGH, cop=2.5,
             w_energy=100.0, w_comfort=1.0, start_time=0.0, dtype=np.float32):
    """theta[7], x0[3], d[H,4], u[H,N] -> cost[N] (sum over the horizon of -reward)."""
    u = np.asarray(u, dtype=dtype)
    d = np.asarray(d, dtype=dtype)
    H, N = u.shape
    A, B, _, _ = rc_abcd(np.asarray(theta)[None], dtype)
    x = np.repeat(np.asarray(x0, dtype=dtype)[None], N, 0)
    price, t_low, t_high = (np.asarray(v, dtype=dtype) for v in (price, t_low, t_high))
    dtf = dtype(dt)
    cost = np.zeros(N, dtype=dtype)
    t = dtype(start_time)
    for k in range(H):
        inp = np.empty((N, 5), dtype=dtype)                      # rc.py:131
        inp[:, 0], inp[:, 1], inp[:, 2], inp[:, 3], inp[:, 4] = d[k, 0], d[k, 1], u[k], d[k, 2], d[k, 3]
        y = x[:, 0].copy()                                       # pre-update output
        x = x + dtf * (_bmv(A, x) + _bmv(B, inp))                # simulator.py:38-40
        t = t + dtf                                              # rc.py:137
        h = int((int(t) % 86400) / 3600)                         # rc.py:238
        energy = (-u[k] / dtype(cop)) * dtf / dtype(3600.0)      # rc.py:221,241
        c = price[h] * energy                                    # rc.py:244
        dTz = np.maximum(y - t_high[h], 0) + np.maximum(t_low[h] - y, 0)   # rc.py:247
        cost += dtype(w_energy) * c + dtype(w_comfort) * dTz     # -(reward) rc.py:250
    return cost


# --------------------------------------------------------------------------------------
# A7 / config 5: general fx/fy block SSM with MLP dynamics, classical RK4 (u held over the step).
# The reference has only the dispatch (base_block_state_space.py:62-63,72-73); its example
# (examples/4-neural-ode.py) does not run.  PARITY UNPINNED BY THE REFERENCE: this restatement is
# the oracle.  Weights are row-major [out,in].
# --------------------------------------------------------------------------------------
def node_mlp_rhs(W, x, u, dtype):
    W1, b1, W2, b2, W3, b3 = W
    z = np.concatenate([x, u], axis=-1)
    h1 = np.tanh(z @ W1.T + b1, dtype=dtype)
    h2 = np.tanh(h1 @ W2.T + b2, dtype=dtype)
    return (h2 @ W3.T + b3).astype(dtype)


def node_rk4_rollout(W1, b1, W2, b2, W3, b3, x0, u, dt, dtype=np.float32, euler=False, p=1):
    """x0[N,n], u[T,N|1,m] -> xsol[T,N,n] (x_{k+1}), ysol[T,N,p] (y_k = x_k[0:p], pre-update)."""
    W = tuple(np.asarray(w, dtype=dtype) for w in (W1, b1, W2, b2, W3, b3))
    x = np.array(x0, dtype=dtype, copy=True)
    u = np.asarray(u, dtype=dtype)
    T, N, n = u.shape[0], x.shape[0], x.shape[1]
    dt = dtype(dt)
    xsol = np.empty((T, N, n), dtype=dtype); ysol = np.empty((T, N, p), dtype=dtype)
    for k in range(T):
        uk = np.broadcast_to(u[k], (N, u.shape[-1]))
        ysol[k] = x[:, :p]
        k1 = node_mlp_rhs(W, x, uk, dtype)
        if euler:
            x = x + dt * k1
        else:
            k2 = node_mlp_rhs(W, x + dtype(0.5) * dt * k1, uk, dtype)
            k3 = node_mlp_rhs(W, x + dtype(0.5) * dt * k2, uk, dtype)
            k4 = node_mlp_rhs(W, x + dt * k3, uk, dtype)
            x = x + (dt / dtype(6)) * (k1 + dtype(2) * k2 + dtype(2) * k3 + k4)
        xsol[k] = x
    return xsol, ysol


def synth_mlp(seed, n=3, m=5, hidden=64, scale=0.1):
    """LeCun-normal x scale weights as row-major [out,in] matrices."""
    rng = np.random.default_rng(seed)
    mk = lambda o, i: (scale * rng.normal(size=(o, i)) / np.sqrt(i)).astype(np.float32)
    return (mk(hidden, n + m), (0.01 * rng.normal(size=hidden)).astype(np.float32), mk(hidden, hidden),
            (0.01 * rng.normal(size=hidden)).astype(np.float32), mk(n, hidden), np.zeros(n, np.float32))


# --------------------------------------------------------------------------------------
# f-1: Tabular time-table lookup      dynax/agents/tabular.py:33-44, dynax/utils/interpolate.py:26-68
# Pinned by tests/test_agents.py:13-25 (tests/golden/golden_tabular.json).
# --------------------------------------------------------------------------------------
def tabular_interp(ts, xs, at, mode="linear", dtype=np.float64):
    ts = np.asarray(ts, dtype=dtype); xs = np.asarray(xs, dtype=dtype); at = np.asarray(at, dtype=dtype).reshape(-1)
    idx = np.interp(at, ts, np.arange(len(ts), dtype=dtype))               # interpolate.py:42 / :63
    if mode == "linear":                                                    # map_coordinates(order=1)
        lo = np.minimum(np.floor(idx).astype(int), len(ts) - 1)
        hi = np.minimum(lo + 1, len(ts) - 1)
        w = (idx - np.floor(idx))[:, None]
        return xs[lo] * (1 - w) + xs[hi] * w
    r = np.floor(np.abs(idx) + 0.5) * np.sign(idx)                          # order=0: nearest, half away from zero
    return xs[np.clip(r.astype(int), 0, len(ts) - 1)]


# --------------------------------------------------------------------------------------
# f-2: one RC-env step for N environments          dynax/wrapper/envs/rc.py:103-158, 200-264
# Pinned by tests/test_gymwrapper_basics.py:7-12 (golden #2).
# --------------------------------------------------------------------------------------
def rc_env_step(theta, x, t, action, ts, dist, dt, price=RC_PRICE, t_low=RC_T_LOW, t_high=RC_T_HIGH, cop=2.5,
                weights=(100.0, 1.0), u_low=-10.0, u_high=0.0, num_actions=101, dtype=np.float32):
    """theta[N|1,7], x[N,3], t[N], action[N] -> dict(x_next, t_next, obs[N,5], reward, terminated, cost, dTz)."""
    x = np.asarray(x, dtype=dtype); t = np.asarray(t, dtype=dtype); N = x.shape[0]
    control = (np.asarray(action) * dtype((u_high - u_low) / (num_actions - 1)) + dtype(u_low)).astype(dtype)   # rc.py:166
    d = tabular_interp(ts, dist, t, "linear", dtype).astype(dtype)                                             # rc.py:128
    inputs = np.stack([d[:, 0], d[:, 1], control, d[:, 2], d[:, 3]], -1).astype(dtype)                         # rc.py:131
    xs, ys = rc_rollout(np.atleast_2d(theta), x, inputs[None], dt, dtype)                                      # rc.py:132
    x_next, y = xs[0], ys[0, :, 0]
    t_next = (t + dtype(dt)).astype(dtype)                                                                     # rc.py:137
    power = (-control / dtype(cop)).astype(dtype)
    obs = np.stack([t_next, y, d[:, 0], d[:, 3], power], -1).astype(dtype)                                     # rc.py:218-222
    h = ((t_next.astype(np.int64) % 86400) / 3600).astype(np.int64)                                            # rc.py:238
    energy = power * dtype(dt) / dtype(3600.0)                                                                 # rc.py:241
    cost = np.asarray(price, dtype)[h] * energy                                                                # rc.py:244
    dTz = np.maximum(y - np.asarray(t_high, dtype)[h], 0) + np.maximum(np.asarray(t_low, dtype)[h] - y, 0)     # rc.py:247
    reward = -dtype(weights[0]) * cost - dtype(weights[1]) * dTz                                               # rc.py:250
    viol = np.maximum(x_next - 40.0, 0).sum(-1) + np.maximum(12.0 - x_next, 0).sum(-1)                         # rc.py:258
    return dict(x_next=x_next, t_next=t_next, obs=obs, reward=reward, terminated=(viol > 0).astype(dtype), cost=cost, dTz=dTz)


# --------------------------------------------------------------------------------------
# f-3: optax.lamb(lr) update for scalar leaves (examples/3-inverse.py:113-122).  optax is an absent,
# unpinned third-party dependency: this restates its published chain scale_by_adam(b1=.9,b2=.999,
# eps=1e-6) -> add_decayed_weights(0) -> scale_by_trust_ratio() -> scale(-lr).  PARITY UNPINNED.
# f-4: data.groupby(index // factor).mean()  (examples/1-forward.py:26-27)
# --------------------------------------------------------------------------------------
def lamb_update(theta, grad, m, v, step, lr=1e-3, b1=0.9, b2=0.999, eps=1e-6):
    theta, grad, m, v = (np.asarray(a, dtype=np.float64) for a in (theta, grad, m, v))
    m = b1 * m + (1 - b1) * grad
    v = b2 * v + (1 - b2) * grad * grad
    upd = (m / (1 - b1 ** step)) / (np.sqrt(v / (1 - b2 ** step)) + eps)
    pn, un = np.abs(theta), np.abs(upd)
    trust = np.where((pn == 0) | (un == 0), 1.0, pn / np.where(un == 0, 1.0, un))
    return theta - lr * trust * upd, m, v


def resample_mean(x, factor):
    x = np.asarray(x, dtype=np.float64)
    rows = x.shape[0]
    return np.stack([x[r:r + factor].mean(0) for r in range(0, rows, factor)])
