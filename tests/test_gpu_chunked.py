"""GPU parity: time-parallel (chunked) RC forward rollout vs the serial kernel and the fp64 oracle."""
import numpy as np
import pytest

from oracle import dynax_oracle as orc
from tests.util import TRAJ_RTOL, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def T():
    import torch
    return torch


def cu(T, a):
    return T.as_tensor(np.ascontiguousarray(a, dtype=np.float32)).cuda()


@pytest.mark.parametrize("N,Tn,chunks", [(64, 8760, -1), (1, 8760, 60), (130, 1000, 7), (33, 257, 4), (8, 100, 1), (5, 64, 64)])
def test_chunked_matches_fp64_and_serial(T, N, Tn, chunks):
    from dynax_b200 import ops
    th, x0, u = orc.synth_theta(N, 111), orc.synth_x0(N, 111), orc.synth_u(Tn, N, 111)
    a = [cu(T, v) for v in (th, x0, u)]
    xs, ys, xf = ops.rc_forward(*a, 3600.0, emit_final=True, time_chunks=chunks)
    x1, y1, xf1 = ops.rc_forward(*a, 3600.0, emit_final=True, time_chunks=0)
    x64, y64 = orc.rc_rollout(th, x0, u, 3600.0, np.float64)
    assert rel_err(xs.cpu().numpy(), x64) < TRAJ_RTOL and rel_err(ys.cpu().numpy(), y64) < TRAJ_RTOL
    assert rel_err(xs.cpu().numpy(), x1.cpu().numpy()) < TRAJ_RTOL
    assert rel_err(xf.cpu().numpy(), xf1.cpu().numpy()) < TRAJ_RTOL
    assert T.equal(xf, xs[-1])


def test_chunked_shared_inputs_and_flags(T):
    from dynax_b200 import ops
    N, Tn = 40, 3000
    th, x0, u = orc.synth_theta(N, 112), orc.synth_x0(N, 112), orc.synth_u(Tn, N, 112)
    xs, ys, _ = ops.rc_forward(cu(T, th[0]), cu(T, x0), cu(T, u[:, 0]), 3600.0, time_chunks=12)   # shared theta + u
    x64, y64 = orc.rc_rollout(th[:1], x0, u[:, :1], 3600.0, np.float64)
    assert rel_err(xs.cpu().numpy(), x64) < TRAJ_RTOL and rel_err(ys.cpu().numpy(), y64) < TRAJ_RTOL
    xo, yo, xf = ops.rc_forward(cu(T, th), cu(T, x0), cu(T, u), 3600.0, emit_x=False, emit_y=True, emit_final=True, time_chunks=9)
    x64, y64 = orc.rc_rollout(th, x0, u, 3600.0, np.float64)
    assert xo is None and rel_err(yo.cpu().numpy(), y64) < TRAJ_RTOL and rel_err(xf.cpu().numpy(), x64[-1]) < TRAJ_RTOL


@pytest.mark.parametrize("N,Tn,chunks", [(64, 8760, -1), (1, 8760, 40), (130, 1000, 7), (33, 257, 4), (8, 100, 1), (5, 70, 9)])
def test_chunked_loss_grad_matches_fp64_and_serial(T, N, Tn, chunks):
    from dynax_b200 import ops
    from tests.util import assert_grad_parity, assert_loss_parity
    th, x0, u = orc.synth_theta(N, 113), orc.synth_x0(N, 113), orc.synth_u(Tn, N, 113)
    target = (22 + np.random.default_rng(9).normal(size=(Tn, N))).astype(np.float32)
    th[0, 0] = 4.0e4                                                     # penalty branch
    a = [cu(T, v) for v in (th, x0, u, target)]
    loss, grad, gx0 = ops.rc_loss_grad(*a, 3600.0, orc.RC_LB, orc.RC_UB, want_gx0=True, time_chunks=chunks)
    l1, g1, gx1 = ops.rc_loss_grad(*a, 3600.0, orc.RC_LB, orc.RC_UB, want_gx0=True, time_chunks=0)
    l64, g64 = orc.rc_loss_grad(th, x0, u, target, 3600.0, orc.RC_LB, orc.RC_UB)
    lbud, gbud = orc.rc_output_sensitivity_budget(th, x0, u, target, 3600.0)
    assert_loss_parity(loss.cpu().numpy(), l64, lbud)
    assert_grad_parity(grad.cpu().numpy(), g64, gbud)
    assert rel_err(gx0.cpu().numpy(), gx1.cpu().numpy(), floor=1.0) < 1e-4
    assert np.allclose(loss.cpu().numpy(), l1.cpu().numpy(), rtol=2e-5)


def test_chunked_loss_grad_shared_modes(T):
    from dynax_b200 import ops
    from tests.util import assert_grad_parity, assert_loss_parity
    N, Tn = 48, 3000
    th, x0, u = orc.synth_theta(N, 114), orc.synth_x0(N, 114), orc.synth_u(Tn, N, 114)
    target = (22 + np.random.default_rng(10).normal(size=(Tn, N))).astype(np.float32)
    # shared theta (sum over systems, penalty once) + shared u + shared target
    th1 = th[0].copy(); th1[6] = 60.0
    loss, grad, _ = ops.rc_loss_grad(cu(T, th1), cu(T, x0), cu(T, u[:, 0]), cu(T, target[:, 0]), 3600.0,
                                     orc.RC_LB, orc.RC_UB, time_chunks=10)
    l64, g64 = orc.rc_loss_grad(np.repeat(th1[None], N, 0), x0, u[:, :1], target[:, :1], 3600.0)
    lbud, gbud = orc.rc_output_sensitivity_budget(np.repeat(th1[None], N, 0), x0, u[:, :1], target[:, :1], 3600.0)
    pen = orc.bound_penalty(th1[None], orc.RC_LB, orc.RC_UB)[0]
    peng = orc.bound_penalty_grad(th1[None], orc.RC_LB, orc.RC_UB)[0]
    assert_loss_parity(loss.cpu().numpy()[0], l64.sum() + pen, lbud.sum())
    assert_grad_parity(grad.cpu().numpy(), g64.sum(0) + peng, gbud.sum(0))


@pytest.mark.parametrize("N,Tn,chunks,use_gx", [(64, 4000, -1, True), (1, 8760, 40, False), (130, 1000, 7, True),
                                                (33, 257, 4, True), (8, 100, 1, False), (5, 70, 9, True)])
def test_chunked_vjp_matches_fp64_and_serial(T, N, Tn, chunks, use_gx):
    """Time-parallel general-cotangent VJP (what autograd through rc_rollout calls for few systems, long horizons)."""
    from dynax_b200 import ops
    rng = np.random.default_rng(115)
    th, x0, u = orc.synth_theta(N, 115), orc.synth_x0(N, 115), orc.synth_u(Tn, N, 115)
    gx = rng.normal(size=(Tn, N, 3)).astype(np.float32) if use_gx else None
    gy = rng.normal(size=(Tn, N, 1)).astype(np.float32)
    a = [cu(T, v) for v in (th, x0, u)]
    args = (None if gx is None else cu(T, gx), cu(T, gy))
    gth, gx0, gu = ops.rc_vjp(*a, 3600.0, *args, time_chunks=chunks)
    sth, sx0, su_ = ops.rc_vjp(*a, 3600.0, *args, time_chunks=0)
    rth, rx0, ru = orc.rc_vjp(th, x0, u, 3600.0, gx, gy, np.float64)
    fth, fx0, fu = orc.rc_vjp(th, x0, u, 3600.0, gx, gy, np.float32)       # yardstick: the same adjoint in fp32
    nrm = lambda p_, q_: np.linalg.norm(p_ - q_) / (np.linalg.norm(q_) + 1e-30)
    assert nrm(gth.cpu().numpy(), rth) <= max(1e-4, 2.0 * nrm(fth, rth))
    assert rel_err(gx0.cpu().numpy(), rx0, floor=1.0) <= max(1e-4, 2.0 * rel_err(fx0, rx0, floor=1.0))
    assert rel_err(gu.cpu().numpy(), ru, floor=1.0) <= max(1e-4, 2.0 * rel_err(fu, ru, floor=1.0))
    assert nrm(gth.cpu().numpy(), sth.cpu().numpy()) <= max(2e-4, 4.0 * nrm(fth, rth))
    # one shared parameter vector: gradient summed over systems
    g1, _, _ = ops.rc_vjp(cu(T, th[0]), a[1], a[2], 3600.0, *args, need_u=False, time_chunks=chunks)
    r1, _, _ = orc.rc_vjp(np.repeat(th[:1], N, 0), x0, u, 3600.0, gx, gy, np.float64)
    f1, _, _ = orc.rc_vjp(np.repeat(th[:1], N, 0), x0, u, 3600.0, gx, gy, np.float32)
    assert nrm(g1.cpu().numpy().reshape(-1), r1.sum(0)) <= max(1e-4, 2.0 * nrm(f1.sum(0), r1.sum(0)))
