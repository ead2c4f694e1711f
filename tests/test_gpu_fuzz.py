"""GPU: seeded sweep over ragged shapes and every input-sharing mode of the fused loss+gradient and of the forward
rollout, against the fp64 oracle.  Exercises tile tails (N around multiples of 32/128/256), chunk tails (T around
multiples of the ring depth), TMA vs cp.async staging (N % 4), narrow-row flattening (N < 7) and both gradient
kernels."""
import itertools

import numpy as np
import pytest

from oracle import dynax_oracle as orc
from tests.util import TRAJ_RTOL, assert_grad_parity, assert_loss_parity, rel_err

pytestmark = pytest.mark.gpu

NS = [1, 2, 3, 4, 5, 31, 33, 127, 129, 224, 255, 257, 300, 449, 512, 515]
TS = [1, 2, 7, 8, 9, 15, 16, 17, 31, 33, 100, 257]


def cu(T, a):
    return T.as_tensor(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def _cases(seed, count):
    rng = np.random.default_rng(seed)
    all_ = list(itertools.product(NS, TS))
    idx = rng.choice(len(all_), size=count, replace=False)
    return [all_[i] for i in idx]


@pytest.mark.parametrize("algo", ["1", "0"])
@pytest.mark.parametrize("N,Tn", _cases(7, 32))
def test_loss_grad_modes(monkeypatch, N, Tn, algo):
    import torch as T
    from dynax_b200 import ops
    monkeypatch.setenv("DYNAX_GRAD_ALGO", algo)
    seed = 1000 + 17 * N + Tn
    th, x0, u = orc.synth_theta(N, seed), orc.synth_x0(N, seed), orc.synth_u(Tn, N, seed)
    target = (22 + np.random.default_rng(seed).normal(size=(Tn, N))).astype(np.float32)
    th[0, 6] = 70.0                                   # one bound violated: penalty path
    for su, stg in ((False, False), (True, True), (True, False), (False, True)):
        if N == 1 and (su or stg):
            continue                                  # a batch of one has nothing to share
        uu = u[:, :1] if su else u
        tt = target[:, :1] if stg else target
        want_gx0 = (N + Tn) % 2 == 0                  # half of the cases also ask for d loss / d x0
        loss, grad, gx0 = ops.rc_loss_grad(cu(T, th), cu(T, x0), cu(T, uu), cu(T, tt[:, 0] if stg else tt), 3600.0,
                                           orc.RC_LB, orc.RC_UB, want_gx0=want_gx0, time_chunks=0)
        l64, g64 = orc.rc_loss_grad(th, x0, uu, tt, 3600.0, orc.RC_LB, orc.RC_UB, np.float64)
        lbud, gbud = orc.rc_output_sensitivity_budget(th, x0, uu, tt, 3600.0)
        assert_loss_parity(loss.cpu().numpy(), l64, lbud, what=(N, Tn, su, stg))
        assert_grad_parity(grad.cpu().numpy(), g64, gbud, what=(N, Tn, su, stg))
        if want_gx0:
            y64 = orc.rc_rollout(th, x0, uu, 3600.0, np.float64)[1][..., 0]
            gy = ((2.0 / Tn) * (y64 - tt))[..., None]
            _, gx0_64, _ = orc.rc_vjp(th, x0, uu, 3600.0, None, gy, np.float64)
            y32 = orc.rc_rollout(th, x0, uu, 3600.0, np.float32)[1][..., 0]
            _, gx0_32, _ = orc.rc_vjp(th, x0, uu, 3600.0, None, ((2.0 / Tn) * (y32 - tt))[..., None].astype(np.float32), np.float32)
            e, e32 = rel_err(gx0.cpu().numpy(), gx0_64, floor=1.0), rel_err(gx0_32, gx0_64, floor=1.0)
            assert e <= max(1e-4, 2.0 * e32), (N, Tn, su, stg, e, e32)
    # shared theta: the sum over systems
    if N > 1:
        th1 = th[1:2]
        loss, grad, _ = ops.rc_loss_grad(cu(T, th1[0]), cu(T, x0), cu(T, u), cu(T, target), 3600.0, time_chunks=0)
        l64, g64 = orc.rc_loss_grad(np.repeat(th1, N, 0), x0, u, target, 3600.0, None, None, np.float64)
        lbud, gbud = orc.rc_output_sensitivity_budget(np.repeat(th1, N, 0), x0, u, target, 3600.0)
        assert_loss_parity(loss.cpu().numpy()[0], l64.sum(), lbud.sum(), what=(N, Tn, "shared theta"))
        assert_grad_parity(grad.cpu().numpy(), g64.sum(0), gbud.sum(0), what=(N, Tn, "shared theta"))


@pytest.mark.parametrize("N,Tn", _cases(8, 32))
def test_forward_and_vjp_shapes(N, Tn):
    import torch as T
    from dynax_b200 import ops
    seed = 2000 + 13 * N + Tn
    th, x0, u = orc.synth_theta(N, seed), orc.synth_x0(N, seed), orc.synth_u(Tn, N, seed)
    xs, ys, xf = ops.rc_forward(cu(T, th), cu(T, x0), cu(T, u), 3600.0, emit_final=True, time_chunks=0)
    xr, yr = orc.rc_rollout(th, x0, u, 3600.0, np.float64)
    assert rel_err(xs.cpu().numpy(), xr) < TRAJ_RTOL and rel_err(ys.cpu().numpy(), yr) < TRAJ_RTOL
    assert T.equal(xf, xs[-1])
    if N > 1:                                          # one shared input series
        xs1, ys1, _ = ops.rc_forward(cu(T, th), cu(T, x0), cu(T, u[:, 0]), 3600.0, time_chunks=0)
        xr1, yr1 = orc.rc_rollout(th, x0, u[:, :1], 3600.0, np.float64)
        assert rel_err(xs1.cpu().numpy(), xr1) < TRAJ_RTOL and rel_err(ys1.cpu().numpy(), yr1) < TRAJ_RTOL
    rng = np.random.default_rng(seed)
    gx = rng.normal(size=(Tn, N, 3)).astype(np.float32)
    gy = rng.normal(size=(Tn, N, 1)).astype(np.float32)
    gth, gx0, gu = ops.rc_vjp(cu(T, th), cu(T, x0), cu(T, u), 3600.0, cu(T, gx), cu(T, gy))
    rth, rx0, ru = orc.rc_vjp(th, x0, u, 3600.0, gx, gy, np.float64)
    fth, fx0, fu = orc.rc_vjp(th, x0, u, 3600.0, gx, gy, np.float32)   # yardstick: the same adjoint in fp32
    nrm = lambda a, b: np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30)
    assert nrm(gth.cpu().numpy(), rth) <= max(1e-4, 2.0 * nrm(fth, rth))
    assert rel_err(gx0.cpu().numpy(), rx0, floor=1.0) <= max(1e-4, 2.0 * rel_err(fx0, rx0, floor=1.0))
    assert rel_err(gu.cpu().numpy(), ru, floor=1.0) <= max(1e-4, 2.0 * rel_err(fu, ru, floor=1.0))


@pytest.mark.parametrize("N,Tn", _cases(9, 24))
def test_chunked_and_ctrl_shapes(N, Tn):
    """Time-parallel kernels with automatic and forced chunk counts, and the shared-series + control-channel
    rollouts, on ragged shapes."""
    import torch as T
    from dynax_b200 import ops
    seed = 3000 + 11 * N + Tn
    th, x0, u = orc.synth_theta(N, seed), orc.synth_x0(N, seed), orc.synth_u(Tn, N, seed)
    target = (22 + np.random.default_rng(seed).normal(size=(Tn, N))).astype(np.float32)
    x64, y64 = orc.rc_rollout(th, x0, u, 3600.0, np.float64)
    l64, g64 = orc.rc_loss_grad(th, x0, u, target, 3600.0, None, None, np.float64)
    lbud, gbud = orc.rc_output_sensitivity_budget(th, x0, u, target, 3600.0)
    for chunks in (-1, 1, 3, Tn):
        xs, ys, _ = ops.rc_forward(cu(T, th), cu(T, x0), cu(T, u), 3600.0, time_chunks=chunks)
        assert rel_err(xs.cpu().numpy(), x64) < TRAJ_RTOL and rel_err(ys.cpu().numpy(), y64) < TRAJ_RTOL, chunks
        loss, grad, _ = ops.rc_loss_grad(cu(T, th), cu(T, x0), cu(T, u), cu(T, target), 3600.0, time_chunks=chunks)
        assert_loss_parity(loss.cpu().numpy(), l64, lbud, what=(N, Tn, chunks))
        assert_grad_parity(grad.cpu().numpy(), g64, gbud, what=(N, Tn, chunks))
    # general-cotangent VJP, time-parallel vs serial (both against the fp64 hand adjoint, fp32 adjoint as yardstick)
    rng = np.random.default_rng(seed)
    gx = rng.normal(size=(Tn, N, 3)).astype(np.float32)
    gy = rng.normal(size=(Tn, N, 1)).astype(np.float32)
    rth, rx0, ru = orc.rc_vjp(th, x0, u, 3600.0, gx, gy, np.float64)
    fth, fx0, fu = orc.rc_vjp(th, x0, u, 3600.0, gx, gy, np.float32)
    nrm = lambda a, b: np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30)
    for chunks in (-1, 2, Tn):
        gth, gx0, gu = ops.rc_vjp(cu(T, th), cu(T, x0), cu(T, u), 3600.0, cu(T, gx), cu(T, gy), time_chunks=chunks)
        assert nrm(gth.cpu().numpy(), rth) <= max(1e-4, 2.0 * nrm(fth, rth)), (N, Tn, chunks)
        assert rel_err(gx0.cpu().numpy(), rx0, floor=1.0) <= max(1e-4, 2.0 * rel_err(fx0, rx0, floor=1.0)), (N, Tn, chunks)
        assert rel_err(gu.cpu().numpy(), ru, floor=1.0) <= max(1e-4, 2.0 * rel_err(fu, ru, floor=1.0)), (N, Tn, chunks)
    # shared disturbances + one per-system channel in column `col`
    for col in (0, 2, 4):
        d = u[:, 0].copy()
        ctrl = u[:, :, col].copy()
        u_full = np.repeat(d[:, None, :], N, 1); u_full[:, :, col] = ctrl
        xr, yr = orc.rc_rollout(th, x0, u_full, 3600.0, np.float64)
        xs, ys, _ = ops.rc_forward_ctrl(cu(T, th), cu(T, x0), cu(T, d), cu(T, ctrl), 3600.0, ctrl_col=col)
        assert rel_err(xs.cpu().numpy(), xr) < TRAJ_RTOL and rel_err(ys.cpu().numpy(), yr) < TRAJ_RTOL
        lr, gr = orc.rc_loss_grad(th, x0, u_full, target, 3600.0, None, None, np.float64)
        lb_, gb_ = orc.rc_output_sensitivity_budget(th, x0, u_full, target, 3600.0)
        loss, grad, _ = ops.rc_loss_grad_ctrl(cu(T, th), cu(T, x0), cu(T, d), cu(T, ctrl), cu(T, target), 3600.0, ctrl_col=col)
        assert_loss_parity(loss.cpu().numpy(), lr, lb_, what=(N, Tn, "ctrl", col))
        assert_grad_parity(grad.cpu().numpy(), gr, gb_, what=(N, Tn, "ctrl", col))


LTI_DIMS = [(1, 1, 1), (2, 1, 1), (2, 2, 1), (2, 2, 2), (3, 1, 1), (3, 3, 1), (3, 3, 3), (3, 5, 1), (4, 1, 1), (4, 2, 2),
            (4, 3, 4), (4, 4, 4)]


@pytest.mark.parametrize("k", range(12))
def test_lti_shapes(k):
    """Dense LTI forward + VJP: every instantiated (n,m,p) on a ragged shape drawn from the same grid."""
    import torch as T
    from dynax_b200 import ops
    n, m, p = LTI_DIMS[k]
    N, Tn = _cases(10, 12)[k]
    rng = np.random.default_rng(4000 + k)
    shared = k % 3 == 0                                  # one set of matrices for all systems
    L = 1 if shared else N
    A = (-np.eye(n)[None] + 0.3 * rng.normal(size=(L, n, n))).astype(np.float32)
    B = rng.normal(size=(L, n, m)).astype(np.float32)
    C = rng.normal(size=(L, p, n)).astype(np.float32)
    D = rng.normal(size=(L, p, m)).astype(np.float32)
    x0 = rng.normal(size=(N, n)).astype(np.float32)
    u = rng.normal(size=(Tn, N, m)).astype(np.float32)
    dt = 0.05
    xs, ys, _ = ops.lti_forward(*(cu(T, v) for v in (A, B, C, D, x0, u)), dt)
    xr, yr = orc.lti_rollout(A, B, C, D, x0, u, dt, np.float64)
    assert rel_err(xs.cpu().numpy(), xr, floor=1.0) < 1e-5 and rel_err(ys.cpu().numpy(), yr, floor=1.0) < 1e-5
    gx = rng.normal(size=(Tn, N, n)).astype(np.float32)
    gy = rng.normal(size=(Tn, N, p)).astype(np.float32)
    out = ops.lti_vjp(*(cu(T, v) for v in (A, B, C, D, x0, u)), dt, cu(T, gx), cu(T, gy))
    r = orc.lti_vjp(A, B, C, D, x0, u, dt, gx, gy, np.float64)
    for got, key in zip(out, ("gA", "gB", "gC", "gD", "gx0", "gu")):
        ref = r[key].sum(0, keepdims=True) if (shared and key in ("gA", "gB", "gC", "gD")) else r[key]
        assert rel_err(got.cpu().numpy().reshape(ref.shape), ref, floor=1.0) < 1e-4, (key, n, m, p, N, Tn, shared)
    # the time-parallel variants with forced chunk counts
    for chunks in (2, Tn):
        xs, ys, _ = ops.lti_forward(*(cu(T, v) for v in (A, B, C, D, x0, u)), dt, time_chunks=chunks)
        assert rel_err(xs.cpu().numpy(), xr, floor=1.0) < 1e-5 and rel_err(ys.cpu().numpy(), yr, floor=1.0) < 1e-5
        out = ops.lti_vjp(*(cu(T, v) for v in (A, B, C, D, x0, u)), dt, cu(T, gx), cu(T, gy), time_chunks=chunks)
        for got, key in zip(out, ("gA", "gB", "gC", "gD", "gx0", "gu")):
            ref = r[key].sum(0, keepdims=True) if (shared and key in ("gA", "gB", "gC", "gD")) else r[key]
            assert rel_err(got.cpu().numpy().reshape(ref.shape), ref, floor=1.0) < 1e-4, (key, n, m, p, N, Tn, shared, chunks)


@pytest.mark.parametrize("N,Tn,euler", [(1, 1, False), (5, 9, False), (127, 2, True), (129, 9, False), (260, 5, False),
                                        (131, 17, True), (2, 33, False)])
def test_node_shapes(monkeypatch, N, Tn, euler):
    """MLP-dynamics rollout and its VJP on ragged shapes: tcgen05 and CUDA-core kernels, stage points saved and
    recomputed, against the fp64 oracle (forward) and each other (gradient)."""
    import torch as T
    from dynax_b200 import ops
    seed = 5000 + 7 * N + Tn
    W = orc.synth_mlp(seed, scale=0.5)
    rng = np.random.default_rng(seed)
    x0 = rng.normal(size=(N, 3)).astype(np.float32)
    u = rng.normal(size=(Tn, N, 5)).astype(np.float32)
    Wc = [cu(T, w) for w in W]
    xr, yr = orc.node_rk4_rollout(*W, x0, u, 0.25, np.float64, euler=euler)
    outs = {}
    for impl in ("1", "0"):
        monkeypatch.setenv("DYNAX_NODE_IMPL", impl)
        xs, ys, _, st = ops.node_rk4_forward(*Wc, cu(T, x0), cu(T, u), 0.25, euler=euler, emit_stages=True)
        assert rel_err(xs.cpu().numpy(), xr, floor=1.0) < 1e-5 and rel_err(ys.cpu().numpy(), yr, floor=1.0) < 1e-5
        outs[impl] = (xs, st)
    xs, st = outs["1"]
    assert (st is None) == euler
    gx = cu(T, rng.normal(size=(Tn, N, 3))); gy = cu(T, rng.normal(size=(Tn, N, 1)))
    grads = []
    for vimpl in ("1", "0"):
        monkeypatch.setenv("DYNAX_NODE_VJP_IMPL", vimpl)
        for stage in ((None, st) if st is not None else (None,)):
            grads.append(ops.node_rk4_vjp(*Wc, cu(T, x0), cu(T, u), xs, 0.25, gx, gy, euler=euler, stage_x=stage))
    for g in grads[1:]:
        for a, b, nm in zip(g, grads[0], ["W1", "b1", "W2", "b2", "W3", "b3", "x0", "u"]):
            assert rel_err(a.cpu().numpy(), b.cpu().numpy(), floor=1.0) < 3e-5, (nm, N, Tn, euler)


@pytest.mark.parametrize("hidden", [32, 128])
@pytest.mark.parametrize("N,Tn,euler", [(1, 1, False), (5, 9, False), (31, 3, True), (33, 7, False), (129, 4, False),
                                        (65, 18, True), (2, 33, False)])
def test_node_shapes_other_hidden_widths(N, Tn, euler, hidden):
    """The same ragged shapes at hidden widths 32 and 128 (one forward and one gradient kernel each): forward against the
    fp64 oracle, gradient with saved against recomputed stage points, shared input series against the materialised one."""
    import torch as T
    from dynax_b200 import ops
    seed = 6000 + 7 * N + Tn + hidden
    W = orc.synth_mlp(seed, hidden=hidden, scale=0.5)
    rng = np.random.default_rng(seed)
    x0 = rng.normal(size=(N, 3)).astype(np.float32)
    u = rng.normal(size=(Tn, N, 5)).astype(np.float32)
    Wc = [cu(T, w) for w in W]
    xr, yr = orc.node_rk4_rollout(*W, x0, u, 0.25, np.float64, euler=euler)
    xs, ys, _, st = ops.node_rk4_forward(*Wc, cu(T, x0), cu(T, u), 0.25, euler=euler, emit_stages=True)
    assert rel_err(xs.cpu().numpy(), xr, floor=1.0) < 1e-5 and rel_err(ys.cpu().numpy(), yr, floor=1.0) < 1e-5
    assert (st is None) == euler
    gx = cu(T, rng.normal(size=(Tn, N, 3))); gy = cu(T, rng.normal(size=(Tn, N, 1)))
    g0 = ops.node_rk4_vjp(*Wc, cu(T, x0), cu(T, u), xs, 0.25, gx, gy, euler=euler)
    if st is not None:
        g1 = ops.node_rk4_vjp(*Wc, cu(T, x0), cu(T, u), xs, 0.25, gx, gy, euler=euler, stage_x=st)
        for a, b, nm in zip(g1, g0, ["W1", "b1", "W2", "b2", "W3", "b3", "x0", "u"]):
            assert rel_err(a.cpu().numpy(), b.cpu().numpy(), floor=1.0) < 3e-5, (nm, N, Tn, hidden)
    # ONE series for all trajectories == the same series materialised per trajectory (g_u summed over trajectories)
    us = np.repeat(u[:, :1], N, 1)
    xs2, _, _ = ops.node_rk4_forward(*Wc, cu(T, x0), cu(T, us), 0.25, euler=euler)
    xs1, _, _ = ops.node_rk4_forward(*Wc, cu(T, x0), cu(T, u[:, 0]), 0.25, euler=euler)
    assert T.equal(xs1, xs2)
    ga = ops.node_rk4_vjp(*Wc, cu(T, x0), cu(T, u[:, 0]), xs1, 0.25, gx, gy, euler=euler)
    gb = ops.node_rk4_vjp(*Wc, cu(T, x0), cu(T, us), xs2, 0.25, gx, gy, euler=euler)
    for a, b, nm in zip(ga[:7], gb[:7], ["W1", "b1", "W2", "b2", "W3", "b3", "x0"]):
        assert rel_err(a.cpu().numpy(), b.cpu().numpy(), floor=1.0) < 3e-5, (nm, N, Tn, hidden)
    assert rel_err(ga[7].cpu().numpy(), gb[7].sum(dim=1, keepdim=True).cpu().numpy(), floor=1.0) < 3e-5


@pytest.mark.parametrize("N,H", [(1, 1), (3, 7), (31, 8), (257, 9), (1000, 96), (4099, 17)])
def test_mpc_shapes(N, H):
    import torch as T
    from dynax_b200 import ops
    rng = np.random.default_rng(6000 + N + H)
    d = np.stack([10 + 15 * rng.random(H), rng.random(H), 2 * rng.random(H), rng.random(H)], 1).astype(np.float32)
    u = (-10 * rng.random((H, N))).astype(np.float32)
    x0 = np.array([20.0, 35.8, 26.0], np.float32)
    sched = [cu(T, v) for v in (orc.RC_PRICE, orc.RC_T_LOW, orc.RC_T_HIGH)]
    start = float(rng.integers(0, 24)) * 3600.0
    cost, amin, cmin = ops.rc_mpc_cost(cu(T, orc.RC_NOMINAL), cu(T, x0), cu(T, d), cu(T, u), *sched, 900.0, start_time=start)
    ref = orc.mpc_cost(orc.RC_NOMINAL, x0, d, u, 900.0, start_time=start, dtype=np.float64)
    assert rel_err(cost.cpu().numpy(), ref, floor=1.0) < 1e-5
    c = cost.cpu().numpy()
    assert amin.item() == int(np.argmin(c)) and cmin.item() == c.min()


@pytest.mark.parametrize("N,Tn,tile", [(261, 40, 128), (5, 17, 0), (1, 9, 0), (515, 33, 128), (300, 8, 256)])
def test_host_api_shapes(N, Tn, tile):
    """Host-buffer entry points with ragged system counts: tiles of the system axis, strided 2-D copies, the last
    tile narrower than the rest (and not a multiple of 4: cp.async staging on the device)."""
    from dynax_b200 import host
    seed = 7000 + N + Tn
    th, x0, u = orc.synth_theta(N, seed), orc.synth_x0(N, seed), orc.synth_u(Tn, N, seed)
    target = (22 + np.random.default_rng(seed).normal(size=(Tn, N))).astype(np.float32)
    xs, ys, xf = host.rc_forward(th, x0, u, 3600.0, tile_systems=tile, emit_final=True)
    xr, yr = orc.rc_rollout(th, x0, u, 3600.0, np.float64)
    assert rel_err(xs, xr) < TRAJ_RTOL and rel_err(ys, yr) < TRAJ_RTOL and np.array_equal(xf, xs[-1])
    loss, grad = host.rc_loss_grad(th, x0, u, target, 3600.0, orc.RC_LB, orc.RC_UB, tile_systems=tile)
    l64, g64 = orc.rc_loss_grad(th, x0, u, target, 3600.0, orc.RC_LB, orc.RC_UB, np.float64)
    lbud, gbud = orc.rc_output_sensitivity_budget(th, x0, u, target, 3600.0)
    assert_loss_parity(loss, l64, lbud, what=(N, Tn, tile))
    assert_grad_parity(grad, g64, gbud, what=(N, Tn, tile))
    host.release()
