"""GPU: row f-1 -- Tabular interpolation kernel vs the reference's golden table, and rollouts fed by a
shared disturbance series + one per-system control channel (how the RC env assembles its inputs)."""
import json
import os

import numpy as np
import pytest

from oracle import c_oracle, dynax_oracle as orc
from tests.util import TRAJ_RTOL, assert_grad_parity, assert_loss_parity, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def T():
    import torch
    return torch


def cu(T, a):
    return T.as_tensor(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def test_tabular_like_reference_test(T, golden_dir):
    """tests/test_agents.py:13-25 with torch tensors."""
    from dynax_b200.agents import Tabular
    g = json.load(open(os.path.join(golden_dir, "golden_tabular.json")))
    ts, xs = T.tensor(g["ts"]), T.tensor(g["xs"])
    t = T.tensor(g["at"]).reshape(-1, 1)
    const = Tabular(ts=ts, xs=xs, mode="constant")
    xt = const.apply(const.init(0, t), t)
    assert np.allclose(xt.cpu().numpy(), g["constant"]), "constant interpolation failed in Tabular agent"
    linear = Tabular(ts=ts, xs=xs, mode="linear")
    xt = linear.apply(linear.init(0, t), t)
    assert np.allclose(xt.cpu().numpy(), g["linear"]), "linear interpolation failed in Tabular agent"
    with pytest.raises(ValueError):
        Tabular(ts=ts, xs=xs, mode="cubic")


def test_tabular_random_vs_oracle(T):
    from dynax_b200 import ops
    rng = np.random.default_rng(121)
    ts = np.cumsum(rng.uniform(0.5, 2.0, size=300)).astype(np.float32)
    xs = rng.normal(size=(300, 4)).astype(np.float32)
    at = rng.uniform(ts[0] - 5, ts[-1] + 5, size=5000).astype(np.float32)
    for mode in ("linear", "constant"):
        got = ops.tabular_interp(cu(T, ts), cu(T, xs), cu(T, at), mode).cpu().numpy()
        ref = orc.tabular_interp(ts, xs, at, mode)
        if mode == "linear":
            assert np.abs(got - ref).max() < 2e-4      # fp32 index arithmetic on t up to ~400
        else:
            assert (np.abs(got - ref).max(1) < 1e-6).mean() > 0.999   # ties at exact halves may differ in fp32


@pytest.mark.parametrize("algo", ["1", "0"])     # DYNAX_GRAD_ALGO: tangent / adjoint kernel
@pytest.mark.parametrize("N,Tn", [(512, 700), (130, 96), (3, 10)])
def test_forward_and_grad_with_control_channel(T, monkeypatch, N, Tn, algo):
    from dynax_b200 import ops
    monkeypatch.setenv("DYNAX_GRAD_ALGO", algo)
    th, x0, u = orc.synth_theta(N, 122), orc.synth_x0(N, 122), orc.synth_u(Tn, N, 122)
    d = u[:, 0].copy()                      # one shared disturbance series [T,5] ...
    ctrl = u[:, :, 2].copy()                # ... and a per-system HVAC channel [T,N]
    u_full = np.repeat(d[:, None, :], N, 1); u_full[:, :, 2] = ctrl
    xs, ys, xf = ops.rc_forward_ctrl(cu(T, th), cu(T, x0), cu(T, d), cu(T, ctrl), 3600.0, emit_final=True)
    xr, yr, _ = c_oracle.rc_forward(th, x0, u_full, 3600.0)
    assert rel_err(xs.cpu().numpy(), xr) < TRAJ_RTOL and rel_err(ys.cpu().numpy(), yr) < TRAJ_RTOL
    # bit-identical to the explicit per-system-u kernel run on the assembled inputs
    x2, y2, _ = ops.rc_forward(cu(T, th), cu(T, x0), cu(T, u_full), 3600.0, time_chunks=0)
    assert T.equal(xs, x2) and T.equal(ys, y2)
    target = (22 + np.random.default_rng(7).normal(size=(Tn, N))).astype(np.float32)
    loss, grad, _ = ops.rc_loss_grad_ctrl(cu(T, th), cu(T, x0), cu(T, d), cu(T, ctrl), cu(T, target), 3600.0,
                                          lb=orc.RC_LB, ub=orc.RC_UB)
    l64, g64 = orc.rc_loss_grad(th, x0, u_full, target, 3600.0, orc.RC_LB, orc.RC_UB)
    lbud, gbud = orc.rc_output_sensitivity_budget(th, x0, u_full, target, 3600.0)
    assert_loss_parity(loss.cpu().numpy(), l64, lbud)
    assert_grad_parity(grad.cpu().numpy(), g64, gbud)
    # same arithmetic per step; only the checkpoint interval (fp32 partial-sum grouping) differs
    l2, g2, _ = ops.rc_loss_grad(cu(T, th), cu(T, x0), cu(T, u_full), cu(T, target), 3600.0, orc.RC_LB, orc.RC_UB,
                                 time_chunks=0)      # the serial kernel: same arithmetic as the control-channel launch
    assert np.allclose(loss.cpu().numpy(), l2.cpu().numpy(), rtol=1e-5)
    assert (np.abs(grad.cpu().numpy() - g2.cpu().numpy()) <= 2 * (1e-4 * np.abs(g64) + gbud)).all()
    if algo == "1":     # the tangent kernel performs identical arithmetic whichever way the inputs are staged
        assert T.equal(loss, l2) and T.equal(grad, g2)
    # another column, ONE shared target series
    col = 0
    u_full0 = np.repeat(d[:, None, :], N, 1); u_full0[:, :, col] = 15 + ctrl
    loss, grad, _ = ops.rc_loss_grad_ctrl(cu(T, th), cu(T, x0), cu(T, d), cu(T, 15 + ctrl), cu(T, target[:, 0]), 3600.0,
                                          ctrl_col=col)
    l64, g64 = orc.rc_loss_grad(th, x0, u_full0, target[:, :1], 3600.0, None, None)
    lbud, gbud = orc.rc_output_sensitivity_budget(th, x0, u_full0, target[:, :1], 3600.0)
    assert_loss_parity(loss.cpu().numpy(), l64, lbud)
    assert_grad_parity(grad.cpu().numpy(), g64, gbud)
