"""GPU: the Flax-shaped facade (DifferentiableSimulator / models / estimators) reads like the
reference's own tests and examples."""
import json
import os

import numpy as np
import pytest

from oracle import dynax_oracle as orc
from tests.util import assert_grad_parity, assert_loss_parity, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def T():
    import torch
    return torch


def test_open_and_closed_loop_like_reference(T, golden_dir):
    """tests/test_simulator.py:38-84 with torch tensors in place of jnp arrays."""
    from dynax_b200.models.RC import Continuous4R3C
    from dynax_b200.simulators import DifferentiableSimulator
    g = json.load(open(os.path.join(golden_dir, "golden1.json")))
    RESULTS = T.tensor(g["RESULTS"], device="cuda")
    model = Continuous4R3C()
    x0 = T.tensor([20.0, 30.0, 26.0], device="cuda")
    dist = T.ones((10, 4)); policy = T.ones((10, 1))
    inputs = T.cat([dist[:, :2], policy, dist[:, 2:]], dim=1).cuda()
    simulator = DifferentiableSimulator(model=model, dt=1.0)
    inits = simulator.init(0, x0, inputs)
    assert set(inits["params"]["model"]) == {"Cai", "Cwe", "Cwi", "Re", "Ri", "Rw", "Rg"}
    xsol, ysol = simulator.apply(inits, x0, inputs)
    assert xsol.shape == (10, 3) and ysol.shape == (10, 1)
    assert T.allclose(xsol, RESULTS[1:])
    x_prev, res = x0, [x0]
    for _ in range(10):
        xsol, _ = simulator.apply(inits, x_prev, T.ones((1, 5)))
        x_prev = xsol[-1]
        res.append(x_prev)
    assert T.allclose(T.stack(res), RESULTS)


def test_example1_forward_on_bundled_series(T, golden_dir):
    """examples/1-forward.py:36-50 (config 1) on the rebuilt hourly series."""
    from dynax_b200.models.RC import Continuous4R3C
    from dynax_b200.simulators import DifferentiableSimulator
    d = np.load(os.path.join(golden_dir, "c1_hourly.npz"))
    rc = dict(zip(("Cai", "Cwe", "Cwi", "Re", "Ri", "Rw", "Rg"), orc.RC_NOMINAL))
    sim = DifferentiableSimulator(Continuous4R3C(), 3600)
    states, outputs = sim.apply({"params": {"model": rc}}, T.tensor([20.0, 30.0, 26.0]), T.tensor(d["u"]))
    xr, yr = orc.rc_rollout_single(orc.RC_NOMINAL, [20.0, 30.0, 26.0], d["u"], 3600.0, np.float64)
    assert rel_err(states.cpu().numpy(), xr) < 1e-5 and rel_err(outputs.cpu().numpy(), yr) < 1e-5


def test_model_apply_single_step_and_linear_ssm(T):
    """tests/test_forward_simulation.py:9-12,15-40: model.apply(params, x, u) -> (new_state, output)."""
    from dynax_b200.core import DiscreteLinearSSM
    from dynax_b200.models.RC import Discrete4R3C
    ssm = DiscreteLinearSSM(state_dim=4, input_dim=3, output_dim=4)
    params = ssm.init(0, T.zeros(4), T.zeros(3))
    assert params["params"]["_fxx"]["dense"]["kernel"].shape == (4, 4)
    assert params["params"]["_fxu"]["dense"]["kernel"].shape == (3, 4)
    assert params["params"]["_fyx"]["dense"]["kernel"].shape == (4, 4)
    assert params["params"]["_fyu"]["dense"]["kernel"].shape == (3, 4)
    state, inp = T.ones(4), T.ones(3)
    k = {n: params["params"][n]["dense"]["kernel"].cpu().double().numpy() for n in ("_fxx", "_fxu", "_fyx", "_fyu")}
    s = np.ones(4)
    for _ in range(20):
        state, out = ssm.apply(params, state, inp)
        o_ref = s @ k["_fyx"] + np.ones(3) @ k["_fyu"]      # nn.Dense: x @ kernel
        s = s @ k["_fxx"] + np.ones(3) @ k["_fxu"]
        assert np.allclose(out.cpu().numpy(), o_ref, rtol=1e-5, atol=1e-6)
    assert np.allclose(state.cpu().numpy(), s, rtol=1e-5, atol=1e-6)
    rc = Discrete4R3C()
    p = {"params": dict(zip(("Cai", "Cwe", "Cwi", "Re", "Ri", "Rw", "Rg"),
                            [6953.9422092947289, 21567.368048437285, 188064.81655062342, 1.4999822619982095,
                             0.55089086571081913, 5.6456475609117183, 3.9933826145529263]))}
    new_state, output = rc.apply(p, T.ones(3), T.full((5,), 0.5))
    A, B, C, D = orc.rc_abcd(np.array(list(p["params"].values())), np.float64)
    assert np.allclose(new_state.cpu().numpy(), A @ np.ones(3) + B @ np.full(5, 0.5), rtol=1e-5)
    assert output.cpu().numpy()[0] == 1.0


def test_missing_dynamics_raises(T):
    from dynax_b200.core import BaseBlockSSM
    from dynax_b200.simulators import DifferentiableSimulator
    m = BaseBlockSSM(state_dim=2, input_dim=1, output_dim=1)
    with pytest.raises(NotImplementedError):       # base_block_state_space.py:65
        DifferentiableSimulator(m, 1.0).apply({"params": {"model": {}}}, T.zeros(2), T.zeros(3, 1))


def test_inverse_step_like_example3(T):
    """examples/3-inverse.py:93-115: value_and_grad(mse_loss) -- fused kernel vs autograd-through-apply
    vs the fp64 oracle, for one shared model and for a batch of candidates."""
    from dynax_b200.estimators import mse_loss_and_grad
    from dynax_b200.models.RC import Continuous4R3C, PARAM_NAMES
    from dynax_b200.simulators import DifferentiableSimulator
    Tn = 600
    th = orc.synth_theta(1, 51)[0]; th[0] = 3.5e4
    x0 = np.array([21.0, 36.0, 25.0], np.float32)
    u = orc.synth_u(Tn, 1, 51)[:, 0]
    target = (22 + np.random.default_rng(2).normal(size=Tn)).astype(np.float32)
    sim = DifferentiableSimulator(Continuous4R3C(), 3600)
    lbt = {"params": {"model": dict(zip(PARAM_NAMES, orc.RC_LB))}}
    ubt = {"params": {"model": dict(zip(PARAM_NAMES, orc.RC_UB))}}
    leaves = {k: T.tensor(float(v), device="cuda", requires_grad=True) for k, v in zip(PARAM_NAMES, th)}
    params = {"params": {"model": leaves}}
    loss, grads = mse_loss_and_grad(sim, params, x0, u, target, lbt, ubt)
    l64, g64 = orc.rc_loss_grad(th[None], x0[None], u[:, None], target[:, None], 3600.0, orc.RC_LB, orc.RC_UB)
    assert np.allclose(loss.item(), l64[0], rtol=1e-5)
    got = np.array([grads["params"]["model"][k].item() for k in PARAM_NAMES])
    assert np.linalg.norm(got - g64[0]) / np.linalg.norm(g64[0]) < 1e-4
    # the same through simulator.apply + autograd (what jax.value_and_grad does in the example)
    _, y = sim.apply(params, T.tensor(x0), T.tensor(u))
    lb_t, ub_t = T.tensor(orc.RC_LB, device="cuda").float(), T.tensor(orc.RC_UB, device="cuda").float()
    thv = T.stack([leaves[k] for k in PARAM_NAMES])
    l2 = ((y.reshape(-1, 1) - T.tensor(target).cuda().reshape(-1, 1)) ** 2).mean() \
        + (T.relu(thv - ub_t) / ub_t).sum() + (T.relu(lb_t - thv) / ub_t).sum()
    l2.backward()
    got2 = np.array([leaves[k].grad.item() for k in PARAM_NAMES])
    assert np.allclose(l2.item(), l64[0], rtol=1e-5)
    assert np.linalg.norm(got2 - g64[0]) / np.linalg.norm(g64[0]) < 1e-4
    # batch of candidates: [N] leaves
    N = 64
    thN = orc.synth_theta(N, 52)
    x0N = orc.synth_x0(N, 52); uN = orc.synth_u(Tn, N, 52)
    tgN = (22 + np.random.default_rng(3).normal(size=(Tn, N))).astype(np.float32)
    pN = {"params": {"model": {k: T.tensor(thN[:, j]).cuda() for j, k in enumerate(PARAM_NAMES)}}}
    lossN, gN = mse_loss_and_grad(sim, pN, T.tensor(x0N), T.tensor(uN), T.tensor(tgN), lbt, ubt)
    l64, g64 = orc.rc_loss_grad(thN, x0N, uN, tgN, 3600.0, orc.RC_LB, orc.RC_UB)
    gotN = np.stack([gN["params"]["model"][k].cpu().numpy() for k in PARAM_NAMES], -1)
    lbud, gbud = orc.rc_output_sensitivity_budget(thN, x0N, uN, tgN, 3600.0)
    assert_loss_parity(lossN.cpu().numpy(), l64, lbud)
    assert_grad_parity(gotN, g64, gbud)


@pytest.mark.parametrize("hidden", [64, 32, 128])
def test_neural_ode_through_simulator(T, hidden):
    """General fx/fy path (base_block_state_space.py:62-63,72-73) through the same simulator facade,
    Euler (the reference's integrator) and RK4 (extension); gradients flow to the Dense kernels.  Hidden widths 64
    (tcgen05 gradient kernel), 32 and 128 (CUDA-core gradient kernels)."""
    from dynax_b200.models import NeuralODE
    from dynax_b200.simulators import DifferentiableSimulator
    model = NeuralODE(hidden=hidden)
    N, Tn = 96, 30
    rng = np.random.default_rng(171)
    x0 = T.tensor(rng.normal(size=(N, 3)), dtype=T.float32).cuda()
    u = T.tensor(rng.normal(size=(Tn, N, 5)), dtype=T.float32).cuda()
    for solver in ("euler", "rk4"):
        sim = DifferentiableSimulator(model, dt=0.1, solver=solver)
        variables = sim.init(3, x0, u)
        fx = variables["params"]["model"]["_fx"]
        assert fx["dense_0"]["kernel"].shape == (8, hidden) and fx["dense_2"]["kernel"].shape == (hidden, 3)
        for layer in fx.values():
            layer["kernel"].requires_grad_(True); layer["bias"].requires_grad_(True)
        xsol, ysol = sim.apply(variables, x0, u)
        W = [w.detach().cpu().numpy() for w in model.matrices(variables["params"]["model"])]
        xr, yr = orc.node_rk4_rollout(*W, x0.cpu().numpy(), u.cpu().numpy(), 0.1, np.float64, euler=(solver == "euler"))
        assert rel_err(xsol.detach().cpu().numpy(), xr, floor=1.0) < 1e-5
        assert rel_err(ysol.detach().cpu().numpy(), yr, floor=1.0) < 1e-5
        (ysol ** 2).mean().backward()
        assert all(layer["kernel"].grad is not None and bool(T.isfinite(layer["kernel"].grad).all()) for layer in fx.values())
    xs1, _ = DifferentiableSimulator(model, dt=0.1, solver="rk4").apply(variables, x0[0], u[:, 0])   # unbatched call shape
    assert xs1.shape == (Tn, 3)


def test_model_apply_and_solver_errors_for_every_model(T):
    """BaseBlockSSM.__call__ -> (rhs, y) (base_block_state_space.py:51-57) works for the general fx/fy model too, and
    solver='rk4' on a model without RK4 support is a clear ValueError, not a TypeError from a forwarded kwarg."""
    from dynax_b200.models import Continuous4R3C, NeuralODE
    from dynax_b200.simulators import DifferentiableSimulator
    rng = np.random.default_rng(181)
    model = NeuralODE()
    x = T.tensor(rng.normal(size=(5, 3)), dtype=T.float32).cuda()
    u = T.tensor(rng.normal(size=(5, 5)), dtype=T.float32).cuda()
    variables = model.init(4, x, u)
    rhs, y = model.apply(variables, x, u)
    W = [w.detach().cpu().double() for w in model.matrices(variables["params"])]
    z = T.cat([x.cpu().double(), u.cpu().double()], -1)
    ref = T.tanh(T.tanh(z @ W[0].T + W[1]) @ W[2].T + W[3]) @ W[4].T + W[5]
    assert rel_err(rhs.cpu().numpy(), ref.numpy(), floor=1.0) < 1e-5 and T.equal(y.cpu(), x.cpu()[:, :1])
    r1, y1 = model.apply(variables, x[0], u[0])                       # unbatched, like the reference's model(x, u)
    assert r1.shape == (3,) and T.allclose(r1, rhs[0])
    rc = Continuous4R3C()
    v = DifferentiableSimulator(rc, dt=60.0).init(0, x[0], u)            # x0[3], inputs[T=5,5]
    with pytest.raises(ValueError, match="rk4"):
        DifferentiableSimulator(rc, dt=60.0, solver="rk4").apply(v, x[0], u)
