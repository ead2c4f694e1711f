"""The C restatement (oracle/dynax_oracle.c) == NumPy restatement == reference goldens (CPU)."""
import json
import os

import numpy as np

from oracle import c_oracle, dynax_oracle as orc


def test_c_golden1(golden_dir):
    g = json.load(open(os.path.join(golden_dir, "golden1.json")))
    res = np.array(g["RESULTS"], dtype=np.float32)
    xs, ys, xf = c_oracle.rc_forward(np.array(g["theta"]), np.array([g["x0"]]), np.array(g["u"])[:, None, :], g["dt"])
    assert np.array_equal(xs[:, 0], res[1:])
    assert np.array_equal(ys[:, 0, 0], res[:-1, 0])
    assert np.array_equal(xf[0], res[-1])


def test_c_golden2(golden_dir):
    from tests.test_oracle_golden import golden2_inputs
    g = json.load(open(os.path.join(golden_dir, "golden2.json")))
    u, _ = golden2_inputs(g)
    xs, ys, _ = c_oracle.rc_forward(np.array(g["theta"]), np.array([g["x0"]]), u[:, None, :], g["dt"])
    assert [f"{v:.6f}" for v in xs[0, 0]] == [f"{v:.6f}" for v in np.array(g["STATES_NEXT"], dtype=np.float32)]
    assert ys[0, 0, 0] == 20.0


def test_c_matches_numpy_forward_ragged():
    T, N = 500, 37          # ragged: not a multiple of the SIMD block
    th, x0, u = orc.synth_theta(N, 3), orc.synth_x0(N, 3), orc.synth_u(T, N, 3)
    xs, ys, xf = c_oracle.rc_forward(th, x0, u, 3600.0)
    xn, yn = orc.rc_rollout(th, x0, u, 3600.0, np.float32)
    assert np.allclose(xs, xn, rtol=2e-6, atol=0)
    assert np.allclose(ys, yn, rtol=2e-6, atol=0)
    assert np.array_equal(xf, xs[-1])
    # shared theta / shared u broadcasting
    xs1, ys1, _ = c_oracle.rc_forward(th[:1], x0, u[:, :1], 3600.0)
    xn1, yn1 = orc.rc_rollout(th[:1], x0, u[:, :1], 3600.0, np.float32)
    assert np.allclose(xs1, xn1, rtol=2e-6, atol=0)


def test_c_loss_grad_close_to_fp64_truth():
    T, N = 2000, 21
    th, x0, u = orc.synth_theta(N, 4), orc.synth_x0(N, 4), orc.synth_u(T, N, 4)
    th[0, 0] = 4.0e4        # penalty branch
    target = (22 + np.random.default_rng(1).normal(size=(T, N))).astype(np.float32)
    l32, g32 = c_oracle.rc_loss_grad(th, x0, u, target, 3600.0, orc.RC_LB, orc.RC_UB)
    l64, g64 = orc.rc_loss_grad(th, x0, u, target, 3600.0, orc.RC_LB, orc.RC_UB, np.float64)
    assert np.allclose(l32, l64, rtol=5e-5)   # naive fp32 running sum of squares
    # fp32 BPTT noise (SURVEY.md finding #9) -> norm-wise check per system
    nrm = np.linalg.norm(g32 - g64, axis=1) / np.linalg.norm(g64, axis=1)
    assert nrm.max() < 2e-3, nrm.max()


def test_c_fp64_forward_mode_matches_numpy_adjoint_and_budgets():
    """oracle_rc_loss_grad_budget_f64 (forward-mode sensitivities, closed-form d rhs/d theta) == the NumPy hand
    adjoint + chain rule (reverse mode; itself pinned to torch-autograd fp64 in test_oracle_adjoint.py) == the NumPy
    conditioning budgets -- for per-system and shared theta / u / target."""
    T, N = 700, 19
    th, x0, u = orc.synth_theta(N, 5), orc.synth_x0(N, 5), orc.synth_u(T, N, 5)
    th[0, 0] = 4.0e4; th[1, 4] = 0.45      # above ub / below lb
    target = (22 + np.random.default_rng(2).normal(size=(T, N))).astype(np.float32)
    for thx, ux, tgx in ((th, u, target), (th[:1], u, target), (th, u[:, :1], target[:, 0])):
        l, g, lb, gb = c_oracle.rc_loss_grad_budget_f64(thx, x0, ux, tgx, 3600.0, orc.RC_LB, orc.RC_UB)
        thN = np.repeat(thx, N, 0) if thx.shape[0] == 1 else thx
        tgN = tgx if tgx.ndim == 2 else np.repeat(tgx[:, None], N, 1)
        l64, g64 = orc.rc_loss_grad(thN, x0, ux, tgN, 3600.0, orc.RC_LB, orc.RC_UB, np.float64)
        lb64, gb64 = orc.rc_output_sensitivity_budget(thN, x0, ux, tgN, 3600.0)
        assert np.allclose(l, l64, rtol=1e-12) and np.allclose(g, g64, rtol=1e-8, atol=1e-14 * np.abs(g64).max())
        assert np.allclose(lb, lb64, rtol=1e-9) and np.allclose(gb, gb64, rtol=1e-5)   # NumPy budget: central differences of A(theta)
