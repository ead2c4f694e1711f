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
