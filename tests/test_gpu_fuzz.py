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

NS = [1, 2, 3, 4, 5, 31, 33, 127, 129, 224, 255, 257, 300]
TS = [1, 2, 7, 8, 9, 15, 16, 17, 31, 33, 100]


def cu(T, a):
    return T.as_tensor(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def _cases(seed, count):
    rng = np.random.default_rng(seed)
    all_ = list(itertools.product(NS, TS))
    idx = rng.choice(len(all_), size=count, replace=False)
    return [all_[i] for i in idx]


@pytest.mark.parametrize("algo", ["1", "0"])
@pytest.mark.parametrize("N,Tn", _cases(7, 14))
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
        loss, grad, _ = ops.rc_loss_grad(cu(T, th), cu(T, x0), cu(T, uu), cu(T, tt[:, 0] if stg else tt), 3600.0,
                                         orc.RC_LB, orc.RC_UB, time_chunks=0)
        l64, g64 = orc.rc_loss_grad(th, x0, uu, tt, 3600.0, orc.RC_LB, orc.RC_UB, np.float64)
        lbud, gbud = orc.rc_output_sensitivity_budget(th, x0, uu, tt, 3600.0)
        assert_loss_parity(loss.cpu().numpy(), l64, lbud, what=(N, Tn, su, stg))
        assert_grad_parity(grad.cpu().numpy(), g64, gbud, what=(N, Tn, su, stg))
    # shared theta: the sum over systems
    if N > 1:
        th1 = th[1:2]
        loss, grad, _ = ops.rc_loss_grad(cu(T, th1[0]), cu(T, x0), cu(T, u), cu(T, target), 3600.0, time_chunks=0)
        l64, g64 = orc.rc_loss_grad(np.repeat(th1, N, 0), x0, u, target, 3600.0, None, None, np.float64)
        lbud, gbud = orc.rc_output_sensitivity_budget(np.repeat(th1, N, 0), x0, u, target, 3600.0)
        assert_loss_parity(loss.cpu().numpy()[0], l64.sum(), lbud.sum(), what=(N, Tn, "shared theta"))
        assert_grad_parity(grad.cpu().numpy(), g64.sum(0), gbud.sum(0), what=(N, Tn, "shared theta"))


@pytest.mark.parametrize("N,Tn", _cases(8, 14))
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
