"""GPU: the *_host C-ABI entry points (NumPy buffers in, NumPy buffers out, tiled + overlapped)."""
import ctypes

import numpy as np
import pytest

from oracle import c_oracle, dynax_oracle as orc
from tests.util import assert_grad_parity, assert_loss_parity, rel_err

pytestmark = pytest.mark.gpu


def p(a):
    return None if a is None else a.ctypes.data


def test_forward_host_tiled_matches_oracle():
    from dynax_b200 import host
    N, Tn = 1000, 300
    th, x0, u = orc.synth_theta(N, 61), orc.synth_x0(N, 61), orc.synth_u(Tn, N, 61)
    xr, yr, xfr = c_oracle.rc_forward(th, x0, u, 3600.0)
    for tile in (0, 256, 384, 1000):
        xs, ys, xf = host.rc_forward(th, x0, u, 3600.0, tile_systems=tile, emit_final=True)
        assert rel_err(xs, xr) < 1e-5 and rel_err(ys, yr) < 1e-5 and np.array_equal(xf, xs[-1])
    xs1, _, _ = host.rc_forward(th[0], x0, u[:, 0], 3600.0, tile_systems=256)     # shared theta + u
    xr1, _, _ = c_oracle.rc_forward(th[:1], x0, u[:, :1], 3600.0)
    assert rel_err(xs1, xr1) < 1e-5


def test_loss_grad_host_tiled_matches_device_api():
    import torch
    from dynax_b200 import host, ops
    N, Tn = 900, 250
    th, x0, u = orc.synth_theta(N, 62), orc.synth_x0(N, 62), orc.synth_u(Tn, N, 62)
    target = (22 + np.random.default_rng(4).normal(size=(Tn, N))).astype(np.float32)
    th[3, 0] = 5e4
    l64, g64 = orc.rc_loss_grad(th, x0, u, target, 3600.0, orc.RC_LB, orc.RC_UB)
    lbud, gbud = orc.rc_output_sensitivity_budget(th, x0, u, target, 3600.0)
    for tile in (0, 256):
        loss, grad = host.rc_loss_grad(th, x0, u, target, 3600.0, orc.RC_LB, orc.RC_UB, tile_systems=tile)
        assert_loss_parity(loss, l64, lbud)
        assert_grad_parity(grad, g64, gbud)
    ld, gd, _ = ops.rc_loss_grad(*(torch.tensor(v).cuda() for v in (th, x0, u, target)), 3600.0, orc.RC_LB, orc.RC_UB)
    assert np.array_equal(ld.cpu().numpy(), loss) and np.array_equal(gd.cpu().numpy(), grad)
    # shared theta across tiles: sums reduced on the host, penalty once
    th1 = th[3].copy()
    loss1, grad1 = host.rc_loss_grad(th1, x0, u, target, 3600.0, orc.RC_LB, orc.RC_UB, tile_systems=256)
    lr, gr = orc.rc_loss_grad(np.repeat(th1[None], N, 0), x0, u, target, 3600.0)
    pen = orc.bound_penalty(th1[None], orc.RC_LB, orc.RC_UB)[0]
    gs = gr.sum(0) + orc.bound_penalty_grad(th1[None], orc.RC_LB, orc.RC_UB)[0]
    lbud, gbud = orc.rc_output_sensitivity_budget(np.repeat(th1[None], N, 0), x0, u, target, 3600.0)
    assert_loss_parity(loss1[0], lr.sum() + pen, lbud.sum())
    assert_grad_parity(grad1, gs, gbud.sum(0))


def test_loss_grad_ctrl_host_matches_device_api():
    import torch
    from dynax_b200 import host, ops
    N, Tn = 700, 200
    th, x0, u = orc.synth_theta(N, 63), orc.synth_x0(N, 63), orc.synth_u(Tn, N, 63)
    d, ctrl = u[:, 0].copy(), u[:, :, 2].copy()
    target = (22 + np.random.default_rng(8).normal(size=(Tn, N))).astype(np.float32)
    loss, grad = host.rc_loss_grad_ctrl(th, x0, d, ctrl, target, 3600.0, lb=orc.RC_LB, ub=orc.RC_UB, tile_systems=256)
    c = lambda a: torch.tensor(np.ascontiguousarray(a)).cuda()
    ld, gd, _ = ops.rc_loss_grad_ctrl(c(th), c(x0), c(d), c(ctrl), c(target), 3600.0, lb=orc.RC_LB, ub=orc.RC_UB)
    assert np.array_equal(ld.cpu().numpy(), loss) and np.array_equal(gd.cpu().numpy(), grad)


def test_resident_dataset_matches_host_call():
    """dynax_rc_dataset_*: u/target uploaded once, theta/x0 per call (the loop of examples/3-inverse.py:133-138) gives
    exactly what the one-shot host call gives, for per-system and shared theta / u / target, with and without bounds."""
    from dynax_b200 import host
    N, Tn = 300, 700
    th, x0, u = orc.synth_theta(N, 71), orc.synth_x0(N, 71), orc.synth_u(Tn, N, 71)
    tg = (22 + np.random.default_rng(3).normal(size=(Tn, N))).astype(np.float32)
    th[2, 3] = 50.0
    for uu, tt in ((u, tg), (u[:, :1], tg[:, 0])):
        ds = host.RcDataset(uu, tt, n_systems=N)
        for thx in (th, th[0]):
            for k in range(2):                                     # second call: parameters changed, data resident
                thk = (thx * (1.0 + 0.01 * k)).astype(np.float32)
                l1, g1 = ds.loss_grad(thk, x0, 3600.0, orc.RC_LB, orc.RC_UB)
                l2, g2 = host.rc_loss_grad(thk, x0, uu, tt, 3600.0, orc.RC_LB, orc.RC_UB)
                if thx.ndim == 2:
                    assert np.array_equal(l1, l2) and np.array_equal(g1, g2)
                else:   # shared theta: fp64 atomics across warps -> order-dependent last bits
                    assert np.allclose(l1, l2, rtol=1e-6) and np.allclose(g1, g2, rtol=1e-5, atol=1e-9 * np.abs(g2).max())
        ds.close()
    # the oracle on a sample
    ds = host.RcDataset(u, tg)
    l, g = ds.loss_grad(th, x0, 3600.0)
    l64, g64 = orc.rc_loss_grad(th[:16], x0[:16], u[:, :16], tg[:, :16], 3600.0, None, None, np.float64)
    assert np.allclose(l[:16], l64, rtol=1e-4)
    ds.close()


def test_host_register_pins_numpy_buffers():
    from dynax_b200 import host
    N, Tn = 256, 300
    th, x0, u = orc.synth_theta(N, 72), orc.synth_x0(N, 72), orc.synth_u(Tn, N, 72)
    tg = (22 + np.random.default_rng(4).normal(size=(Tn, N))).astype(np.float32)
    l0, g0 = host.rc_loss_grad(th, x0, u, tg, 3600.0)
    with host.pinned(u, tg):
        l1, g1 = host.rc_loss_grad(th, x0, u, tg, 3600.0)
    assert np.array_equal(l0, l1) and np.array_equal(g0, g1)
