"""GPU: batched RC-env step kernel (row f-2) vs the reference's env golden and the oracle."""
import json
import os

import numpy as np
import pytest

from oracle import dynax_oracle as orc
from tests.test_oracle_env import golden2_table
from tests.util import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def T():
    import torch
    return torch


def test_basic_forward_like_reference(T, golden_dir):
    """tests/test_gymwrapper_basics.py:22-43 through the mirrored env API."""
    from dynax_b200.wrapper import RC
    g = json.load(open(os.path.join(golden_dir, "golden2.json")))
    ts, dist = golden2_table(g)
    env = RC(ts, dist, start_time=0, end_time=3600, dt=g["dt"], num_actions=11, name="RC-V1")
    obs, states, params = env.reset(0)
    obs_next, reward, terminated, truncated, info, states = env.step(7, states=states, params=params)
    assert T.allclose(obs_next.cpu(), T.tensor(g["OBS_NEXT"]))
    assert abs(reward.item() - g["REWARD"]) < 1e-6
    assert bool(terminated.item()) is False and truncated is False
    assert states.time.item() == g["TIME_NEXT"]
    assert T.allclose(states.x.cpu(), T.tensor(g["STATES_NEXT"]), rtol=0, atol=4e-6)


def test_batched_multi_step_vs_oracle(T):
    from dynax_b200.wrapper import RC, EnvStates
    rng = np.random.default_rng(131)
    R, N, K, dt = 500, 777, 30, 60.0
    ts = np.arange(R) * 60.0
    dist = np.stack([15 + 10 * rng.random(R), rng.random(R), 2 * rng.random(R), rng.random(R)], 1)
    env = RC(ts, dist, dt=dt, num_actions=11)
    x = (np.array([20.0, 30.0, 26.0])[None] + rng.uniform(-9, 15, size=(N, 3))).astype(np.float32)   # some envs terminate
    t = (rng.integers(0, 300, size=N) * 60.0 + rng.uniform(0, 60, size=N)).astype(np.float32)        # own clock per env
    actions = rng.integers(0, 11, size=(K, N))
    out = env.rollout(T.tensor(actions).cuda(), EnvStates(x=T.tensor(x).cuda(), time=T.tensor(t).cuda()))
    xr, tr = x.copy(), t.copy()
    for k in range(K):
        r = orc.rc_env_step(orc.RC_NOMINAL, xr, tr, actions[k], ts, dist, dt, num_actions=11, dtype=np.float64)
        ob = out["obs"][k].cpu().numpy()
        assert rel_err(ob[:, 1], r["obs"][:, 1]) < 1e-5                       # zone temperature (pre-update)
        # the fp32 lookup time (t ~ 2e4 s, ulp 2e-3 s) limits the interpolated disturbances to ~1e-4
        assert rel_err(ob[:, 2], r["obs"][:, 2], floor=1.0) < 3e-4 and rel_err(ob[:, 3], r["obs"][:, 3], floor=1.0) < 3e-4
        assert np.allclose(ob[:, 4], r["obs"][:, 4], rtol=1e-6) and np.allclose(ob[:, 0], r["obs"][:, 0])
        assert np.allclose(out["reward"][k].cpu().numpy(), r["reward"], rtol=1e-4, atol=1e-5)
        assert (out["terminated"][k].cpu().numpy() == r["terminated"]).mean() > 0.999
        xr, tr = r["x_next"], r["t_next"]
    assert rel_err(out["x_next"].cpu().numpy(), xr) < 1e-5
    assert out["terminated"].sum().item() > 0


def test_negative_and_late_start_times(T):
    """rc.py:238 h = int((int(t) % 86400)/3600) with Python/jnp semantics: a negative start time gives a non-negative
    remainder (C's % does not), so price/comfort tables are never read out of bounds (round-1 advisor finding)."""
    from dynax_b200.wrapper import RC, EnvStates
    rng = np.random.default_rng(141)
    R, N, K, dt = 300, 64, 5, 900.0
    ts = (np.arange(R) - 150) * 900.0                                   # table covers negative times too
    dist = np.stack([15 + 10 * rng.random(R), rng.random(R), 2 * rng.random(R), rng.random(R)], 1)
    env = RC(ts, dist, dt=dt, num_actions=11)
    x = (np.array([20.0, 30.0, 26.0])[None] + rng.uniform(-1, 1, size=(N, 3))).astype(np.float32)
    t = (rng.integers(-140, 100, size=N) * 900.0).astype(np.float32)     # many envs start before t = 0
    actions = rng.integers(0, 11, size=(K, N))
    out = env.rollout(T.tensor(actions).cuda(), EnvStates(x=T.tensor(x).cuda(), time=T.tensor(t).cuda()))
    xr, tr = x.copy(), t.copy()
    for k in range(K):
        r = orc.rc_env_step(orc.RC_NOMINAL, xr, tr, actions[k], ts, dist, dt, num_actions=11, dtype=np.float64)
        assert np.allclose(out["reward"][k].cpu().numpy(), r["reward"], rtol=1e-4, atol=1e-5)
        xr, tr = r["x_next"], r["t_next"]
    assert bool(T.isfinite(out["reward"]).all())
