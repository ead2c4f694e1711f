"""Pin the NumPy oracle against the reference's own known-answer tests (CPU, no GPU).

golden #1: /root/reference/tests/test_simulator.py:24-57,60-84
golden #2: /root/reference/tests/test_gymwrapper_basics.py:7-43
"""
import json
import os

import numpy as np

from oracle import dynax_oracle as orc


def _g(golden_dir, name):
    return json.load(open(os.path.join(golden_dir, name)))


def test_golden1_open_loop(golden_dir):
    g = _g(golden_dir, "golden1.json")
    res = np.array(g["RESULTS"], dtype=np.float32)
    xsol, ysol = orc.rc_rollout_single(g["theta"], g["x0"], np.array(g["u"]), g["dt"])
    # test_simulator.py:57  jnp.allclose(xsol, RESULTS[1:])  -- all-integer table, exact in fp32
    assert np.array_equal(xsol, res[1:])
    # ysol[k] = C x_k (pre-update)  simulator.py:38-43
    assert np.array_equal(ysol[:, 0], res[:-1, 0])


def test_golden1_closed_loop(golden_dir):
    # test_simulator.py:60-84: ten chained T=1 rollouts give the same table
    g = _g(golden_dir, "golden1.json")
    res = np.array(g["RESULTS"], dtype=np.float32)
    x = np.array(g["x0"], dtype=np.float32)
    rows = [x]
    for _ in range(10):
        xs, _ = orc.rc_rollout_single(g["theta"], x, np.ones((1, 5)), g["dt"])
        x = xs[-1]
        rows.append(x)
    assert np.array_equal(np.stack(rows), res)


def golden2_inputs(g):
    d = g["disturbance_row0"]
    ctrl = g["action"] * (g["u_high"] - g["u_low"]) / (g["num_actions"] - 1) + g["u_low"]   # rc.py:166
    # rc.py:131  inputs = [dist0, dist1, action, dist2, dist3]
    return np.array([[d["out_temp"], d["qint_lump"], ctrl, d["qwin_lump"], d["qradin_lump"]]], dtype=np.float32), ctrl


def test_golden2_env_step(golden_dir):
    g = _g(golden_dir, "golden2.json")
    u, ctrl = golden2_inputs(g)
    assert ctrl == -3.0
    xs, ys = orc.rc_rollout_single(g["theta"], g["x0"], u, g["dt"])
    want = np.array(g["STATES_NEXT"], dtype=np.float32)
    # test_gymwrapper_basics.py:43 uses jnp.allclose; the fp32 restatement hits every printed digit
    assert np.allclose(xs[0], want, rtol=0, atol=2e-6), (xs[0], want)
    assert [f"{v:.6f}" for v in xs[0]] == [f"{v:.6f}" for v in want]
    # observation = [time, y (pre-update Tz), Tout, qradin, -u/COP]   rc.py:200-222
    obs = np.array([g["TIME_NEXT"], ys[0, 0], u[0, 0], u[0, 4], -ctrl / g["COP"]], dtype=np.float32)
    assert np.allclose(obs, np.array(g["OBS_NEXT"], dtype=np.float32), rtol=1e-6)
    # reward  rc.py:227-252 (hour index from post-update time = 60 s -> h=0)
    energy = np.float32(obs[4] * np.float32(g["dt"]) / np.float32(3600.0))
    cost = np.float32(g["price"][0]) * energy
    dTz = max(obs[1] - g["T_high_h0"], 0) + max(g["T_low_h0"] - obs[1], 0)
    reward = -np.float32(g["weights"][0]) * cost - np.float32(g["weights"][1]) * np.float32(dTz)
    assert abs(float(reward) - g["REWARD"]) < 1e-7


def test_fp32_vs_fp64_budget():
    # SURVEY.md finding #8: residual-form fp32 Euler stays within 1e-5 (max rel) of fp64 over a year
    T, N = 8760, 4
    th = orc.synth_theta(N, 7)
    x0 = orc.synth_x0(N, 7)
    u = orc.synth_u(T, N, 7)
    x32, _ = orc.rc_rollout(th, x0, u, 3600.0, np.float32)
    x64, _ = orc.rc_rollout(th, x0, u, 3600.0, np.float64)
    rel = np.abs(x32 - x64) / np.maximum(np.abs(x64), 1e-3 * np.abs(x64).max())
    assert rel.max() < 1e-5, rel.max()
