"""GPU parity: RC forward rollout through the C-ABI vs the oracle and the reference goldens.

Tolerance (BASELINE.json north_star): trajectories within rel 1e-5 (fp32).
"""
import json
import os

import numpy as np
import pytest

from oracle import c_oracle, dynax_oracle as orc
from tests.util import rel_err

pytestmark = pytest.mark.gpu

TRAJ_RTOL = 1e-5


@pytest.fixture(scope="module")
def T():
    import torch
    return torch


@pytest.fixture(scope="module")
def ops():
    from dynax_b200 import ops
    return ops


def cu(T, a):
    return T.as_tensor(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def test_golden1_exact(T, ops, golden_dir):
    g = json.load(open(os.path.join(golden_dir, "golden1.json")))
    res = np.array(g["RESULTS"], dtype=np.float32)
    xs, ys, xf = ops.rc_forward(cu(T, g["theta"]), cu(T, [g["x0"]]), cu(T, g["u"]), g["dt"], emit_final=True)
    assert np.array_equal(xs[:, 0].cpu().numpy(), res[1:])          # tests/test_simulator.py:57
    assert np.array_equal(ys[:, 0, 0].cpu().numpy(), res[:-1, 0])   # pre-update output
    assert np.array_equal(xf[0].cpu().numpy(), res[-1])


def test_golden1_closed_loop(T, ops, golden_dir):
    # tests/test_simulator.py:60-84: ten chained T=1 launches == one T=10 launch
    g = json.load(open(os.path.join(golden_dir, "golden1.json")))
    res = np.array(g["RESULTS"], dtype=np.float32)
    x = cu(T, [g["x0"]])
    rows = [x[0].cpu().numpy()]
    for _ in range(10):
        xs, _, _ = ops.rc_forward(cu(T, g["theta"]), x, cu(T, np.ones((1, 5))), g["dt"])
        x = xs[-1]
        rows.append(x[0].cpu().numpy())
    assert np.array_equal(np.stack(rows), res)


def test_golden2_env_step(T, ops, golden_dir):
    from tests.test_oracle_golden import golden2_inputs
    g = json.load(open(os.path.join(golden_dir, "golden2.json")))
    u, _ = golden2_inputs(g)
    xs, ys, _ = ops.rc_forward(cu(T, g["theta"]), cu(T, [g["x0"]]), cu(T, u), g["dt"])
    got = xs[0, 0].cpu().numpy()
    want = np.array(g["STATES_NEXT"], dtype=np.float32)
    assert np.allclose(got, want, rtol=0, atol=4e-6), (got, want)   # tests/test_gymwrapper_basics.py:43
    assert ys[0, 0, 0].item() == 20.0


@pytest.mark.parametrize("N,Tn", [(4096, 1000), (4099, 333), (130, 257), (1, 50), (3, 1)])
def test_matches_oracle(T, ops, N, Tn):
    # 4099 / 130 / 3: ragged N (tail tile, and N%4!=0 -> the non-TMA producer path)
    th, x0, u = orc.synth_theta(N, 21), orc.synth_x0(N, 21), orc.synth_u(Tn, N, 21)
    xs, ys, xf = ops.rc_forward(cu(T, th), cu(T, x0), cu(T, u), 3600.0, emit_final=True)
    xr, yr, xfr = c_oracle.rc_forward(th, x0, u, 3600.0)
    assert rel_err(xs.cpu().numpy(), xr) < TRAJ_RTOL
    assert rel_err(ys.cpu().numpy(), yr) < TRAJ_RTOL
    assert np.array_equal(xf.cpu().numpy(), xs[-1].cpu().numpy())


def test_year_long_vs_fp64(T, ops):
    N, Tn = 512, 8760
    th, x0, u = orc.synth_theta(N, 22), orc.synth_x0(N, 22), orc.synth_u(Tn, N, 22)
    xs, ys, _ = ops.rc_forward(cu(T, th), cu(T, x0), cu(T, u), 3600.0)
    x64, y64 = orc.rc_rollout(th, x0, u, 3600.0, np.float64)
    assert rel_err(xs.cpu().numpy(), x64) < TRAJ_RTOL
    assert rel_err(ys.cpu().numpy(), y64) < TRAJ_RTOL


def test_shared_theta_and_u(T, ops):
    N, Tn = 260, 400
    th, x0, u = orc.synth_theta(N, 23), orc.synth_x0(N, 23), orc.synth_u(Tn, N, 23)
    xs, ys, _ = ops.rc_forward(cu(T, th[:1]), cu(T, x0), cu(T, u[:, :1]), 3600.0)
    xr, yr, _ = c_oracle.rc_forward(th[:1], x0, u[:, :1], 3600.0)
    assert rel_err(xs.cpu().numpy(), xr) < TRAJ_RTOL and rel_err(ys.cpu().numpy(), yr) < TRAJ_RTOL
    xs2, _, _ = ops.rc_forward(cu(T, th), cu(T, x0), cu(T, u[:, 0]), 3600.0)        # u[T,5] shared
    xr2, _, _ = c_oracle.rc_forward(th, x0, u[:, :1], 3600.0)
    assert rel_err(xs2.cpu().numpy(), xr2) < TRAJ_RTOL


def test_emit_flags_and_no_tma_path(T, ops, monkeypatch):
    N, Tn = 512, 301
    th, x0, u = orc.synth_theta(N, 24), orc.synth_x0(N, 24), orc.synth_u(Tn, N, 24)
    a = [cu(T, v) for v in (th, x0, u)]
    xs, ys, xf = ops.rc_forward(*a, 3600.0, emit_final=True)
    xo, yo, _ = ops.rc_forward(*a, 3600.0, emit_x=True, emit_y=False)
    assert yo is None and T.equal(xo, xs)
    xo, yo, _ = ops.rc_forward(*a, 3600.0, emit_x=False, emit_y=True)
    assert xo is None and T.equal(yo, ys)
    xo, yo, xf2 = ops.rc_forward(*a, 3600.0, emit_x=False, emit_y=False, emit_final=True)
    assert xo is None and yo is None and T.equal(xf2, xf)
    monkeypatch.setenv("DYNAX_NO_BULK", "1")     # plain ld/st producer must give identical bits
    xs3, ys3, _ = ops.rc_forward(*a, 3600.0)
    assert T.equal(xs3, xs) and T.equal(ys3, ys)


@pytest.mark.parametrize("cfg", ["0", "3", "4", "5"])
def test_tile_configs_bit_identical(T, ops, monkeypatch, cfg):
    N, Tn = 1000, 203
    th, x0, u = orc.synth_theta(N, 25), orc.synth_x0(N, 25), orc.synth_u(Tn, N, 25)
    a = [cu(T, v) for v in (th, x0, u)]
    monkeypatch.setenv("DYNAX_FWD_CFG", "0")
    ref = ops.rc_forward(*a, 3600.0)
    monkeypatch.setenv("DYNAX_FWD_CFG", cfg)
    got = ops.rc_forward(*a, 3600.0)
    assert T.equal(ref[0], got[0]) and T.equal(ref[1], got[1])


@pytest.mark.parametrize("Tn", [1, 15, 16, 17, 64, 65, 203])
def test_shared_series_tile_shapes(T, ops, monkeypatch, Tn):
    """A shared input series (with and without a per-system control channel) on a small launch: the long-chunk tile
    (16 rows, 4-stage ring: the default for up to two CTAs per SM) against the fp64 oracle and, bit for bit, against the
    4-row tile (DYNAX_FWD_CFG=0); horizons around one chunk and one ring."""
    N = 517
    th, x0, u = orc.synth_theta(N, 26), orc.synth_x0(N, 26), orc.synth_u(Tn, N, 26)
    us = u[:, :1]
    ctrl = -5.0 * np.random.default_rng(27).random((Tn, N), dtype=np.float32)
    a = [cu(T, v) for v in (th, x0, us)]
    got = ops.rc_forward(*a, 3600.0, emit_final=True)
    gotc = ops.rc_forward_ctrl(cu(T, th), cu(T, x0), cu(T, us[:, 0]), cu(T, ctrl), 3600.0, ctrl_col=2)
    monkeypatch.setenv("DYNAX_FWD_CFG", "0")
    ref = ops.rc_forward(*a, 3600.0, emit_final=True)
    refc = ops.rc_forward_ctrl(cu(T, th), cu(T, x0), cu(T, us[:, 0]), cu(T, ctrl), 3600.0, ctrl_col=2)
    for g, r in zip(got + gotc[:2], ref + refc[:2]):
        assert T.equal(g, r)
    x64, y64 = orc.rc_rollout(th, x0, np.repeat(us, N, 1), 3600.0, np.float64)
    assert rel_err(got[0].cpu().numpy(), x64) < TRAJ_RTOL and rel_err(got[1].cpu().numpy(), y64) < TRAJ_RTOL
    uc = np.repeat(us, N, 1).copy()
    uc[:, :, 2] = ctrl
    xc64, _ = orc.rc_rollout(th, x0, uc, 3600.0, np.float64)
    assert rel_err(gotc[0].cpu().numpy(), xc64) < TRAJ_RTOL


def test_chaining_property_large(T, ops):
    """Size-independent property at a full-machine tile: T steps == T1 steps then T-T1 steps from
    x_final (tests/test_simulator.py:60-84 generalised), bit for bit; and zero in -> zero out."""
    N, Tn, T1 = 148 * 6 * 128, 96, 37
    g = T.Generator(device="cuda").manual_seed(5)
    th = cu(T, orc.RC_NOMINAL[None]) * T.exp(T.empty((N, 7), device="cuda").uniform_(-0.2, 0.2, generator=g))
    x0 = cu(T, [[20.0, 30.0, 26.0]]) + T.empty((N, 3), device="cuda").uniform_(-1, 1, generator=g)
    u = T.empty((Tn, N, 5), device="cuda").uniform_(-1, 1, generator=g)
    xs, ys, xf = ops.rc_forward(th, x0, u, 3600.0, emit_final=True)
    xa, ya, xfa = ops.rc_forward(th, x0, u[:T1].contiguous(), 3600.0, emit_final=True)
    xb, yb, xfb = ops.rc_forward(th, xfa, u[T1:].contiguous(), 3600.0, emit_final=True)
    assert T.equal(T.cat([xa, xb]), xs) and T.equal(T.cat([ya, yb]), ys) and T.equal(xfb, xf)
    z = ops.rc_forward(th, T.zeros_like(x0), T.zeros_like(u), 3600.0)
    assert float(z[0].abs().max()) == 0.0 and float(z[1].abs().max()) == 0.0
    # spot-check 64 random systems of the big launch against the oracle
    idx = np.random.default_rng(0).choice(N, 64, replace=False)
    xr, yr, _ = c_oracle.rc_forward(th[idx].cpu().numpy(), x0[idx].cpu().numpy(), u[:, idx].cpu().numpy(), 3600.0)
    assert rel_err(xs[:, idx].cpu().numpy(), xr) < TRAJ_RTOL


def test_empty_and_errors(T, ops):
    from dynax_b200._lib import DynaxError
    th, x0 = cu(T, orc.synth_theta(8, 1)), cu(T, orc.synth_x0(8, 1))
    xs, ys, xf = ops.rc_forward(th, x0, T.empty((0, 8, 5), device="cuda"), 1.0, emit_final=True)   # T = 0
    assert xs.shape == (0, 8, 3) and ys.shape == (0, 8, 1) and T.equal(xf, x0)
    xs, ys, _ = ops.rc_forward(th[:0], x0[:0], T.empty((5, 0, 5), device="cuda"), 1.0)             # N = 0
    assert xs.shape == (5, 0, 3)
    with pytest.raises(ValueError):
        ops.rc_forward(th, x0, T.empty((5, 7, 5), device="cuda"), 1.0)     # N mismatch
    with pytest.raises(ValueError):
        ops.rc_forward(th.cpu(), x0, T.empty((5, 8, 5), device="cuda"), 1.0)   # no CPU path
    with pytest.raises(TypeError):
        ops.rc_forward(th.double(), x0, T.empty((5, 8, 5), device="cuda"), 1.0)
