"""GPU: rows f-3 / f-4 -- LAMB update kernel, resample kernel, and the on-device calibration loop of
examples/3-inverse.py:117-138."""
import numpy as np
import pytest

from oracle import dynax_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def T():
    import torch
    return torch


def test_lamb_update_vs_oracle(T):
    from dynax_b200 import ops
    rng = np.random.default_rng(141)
    th = (orc.RC_NOMINAL[None] * np.exp(rng.uniform(-0.2, 0.2, size=(500, 7)))).astype(np.float32)
    th[0, 0] = 0.0                                           # zero parameter -> trust ratio 1
    tt = T.tensor(th).cuda(); m = T.zeros_like(tt); v = T.zeros_like(tt)
    rm, rv, rt = np.zeros_like(th, np.float64), np.zeros_like(th, np.float64), th.astype(np.float64)
    for step in range(1, 6):
        g = (rng.normal(size=th.shape) * np.abs(rng.normal(size=th.shape))).astype(np.float32)
        g[1, 1] = 0.0 if step == 1 else g[1, 1]              # zero update -> trust ratio 1
        ops.lamb_update_(tt, T.tensor(g).cuda(), m, v, step, lr=1e-3)
        rt, rm, rv = orc.lamb_update(rt, g, rm, rv, step, lr=1e-3)
        assert np.allclose(tt.cpu().numpy(), rt, rtol=2e-6, atol=1e-9)
    assert np.allclose(m.cpu().numpy(), rm, rtol=1e-4, atol=1e-6)      # fp32 moments vs fp64 replay


def test_resample_mean_vs_oracle(T, golden_dir):
    from dynax_b200 import ops
    rng = np.random.default_rng(142)
    x = rng.normal(size=(1003, 6)).astype(np.float32)        # ragged last group
    got = ops.resample_mean(T.tensor(x).cuda(), 60).cpu().numpy()
    assert got.shape == (17, 6) and np.allclose(got, orc.resample_mean(x, 60), rtol=1e-6, atol=1e-7)


def test_calibrate_recovers_loss_decrease(T):
    """examples/3-inverse.py:124-138 in miniature: start at the upper bounds (as the example does) and
    check the device loop drives the loss down and matches an oracle replay of the same loop."""
    from dynax_b200.estimators import calibrate
    from dynax_b200.models.RC import Continuous4R3C, PARAM_NAMES
    from dynax_b200.simulators import DifferentiableSimulator
    N, Tn = 64, 400
    x0, u = orc.synth_x0(N, 143), orc.synth_u(Tn, N, 143)
    hidden = orc.synth_theta(N, 144)
    _, y = orc.rc_rollout(hidden, x0, u, 3600.0, np.float64)
    target = y[..., 0].astype(np.float32)
    th0 = np.repeat((0.5 * (orc.RC_LB + orc.RC_UB))[None], N, 0).astype(np.float32)
    sim = DifferentiableSimulator(Continuous4R3C(), 3600)
    lbt = {"params": {"model": dict(zip(PARAM_NAMES, orc.RC_LB))}}
    ubt = {"params": {"model": dict(zip(PARAM_NAMES, orc.RC_UB))}}
    p0 = {"params": {"model": {k: T.tensor(th0[:, j]).cuda() for j, k in enumerate(PARAM_NAMES)}}}
    l0, _ = orc.rc_loss_grad(th0, x0, u, target, 3600.0, orc.RC_LB, orc.RC_UB)
    fitted, loss = calibrate(sim, p0, T.tensor(x0), T.tensor(u), T.tensor(target), lbt, ubt, n_epochs=40, lr=1e-2)
    th_fit = np.stack([fitted["params"]["model"][k].cpu().numpy() for k in PARAM_NAMES], -1)
    assert (loss.cpu().numpy() < l0).mean() > 0.9
    # oracle replay of the same 40 epochs (fp64 gradients): parameters agree to the accumulated fp32 noise
    th, m, v = th0.astype(np.float64), np.zeros((N, 7)), np.zeros((N, 7))
    for ep in range(1, 41):
        _, g = orc.rc_loss_grad(th, x0, u, target, 3600.0, orc.RC_LB, orc.RC_UB)
        th, m, v = orc.lamb_update(th, g, m, v, ep, lr=1e-2)
    assert np.allclose(th_fit, th, rtol=2e-3)


def test_calibrate_cuda_graph_matches_eager(T):
    """The captured-epoch loop gives the same parameters as the eager loop (same kernels, same order)."""
    import time
    from dynax_b200.estimators import calibrate
    from dynax_b200.models.RC import Continuous4R3C, PARAM_NAMES
    from dynax_b200.simulators import DifferentiableSimulator
    N, Tn = 4, 8760                      # launch-bound regime: few systems, one year (time-parallel kernels)
    x0, u = orc.synth_x0(N, 145), orc.synth_u(Tn, N, 145)
    target = (22 + np.random.default_rng(12).normal(size=(Tn, N))).astype(np.float32)
    th0 = orc.synth_theta(N, 146)
    sim = DifferentiableSimulator(Continuous4R3C(), 3600)
    mk = lambda: {"params": {"model": {k: T.tensor(th0[:, j]).cuda() for j, k in enumerate(PARAM_NAMES)}}}
    args = (T.tensor(x0), T.tensor(u), T.tensor(target))
    out = {}
    for graph in (False, True):
        T.cuda.synchronize(); t0 = time.perf_counter()
        fitted, loss = calibrate(sim, mk(), *args, n_epochs=60, lr=1e-2, cuda_graph=graph)
        T.cuda.synchronize(); out[graph] = (time.perf_counter() - t0,
                                            np.stack([fitted["params"]["model"][k].cpu().numpy() for k in PARAM_NAMES], -1))
    assert np.allclose(out[True][1], out[False][1], rtol=1e-6)
    print("calibrate 60 epochs N=4 T=8760: eager %.1f ms, cuda graph %.1f ms" % (out[False][0] * 1e3, out[True][0] * 1e3))
