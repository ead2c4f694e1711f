"""Mirror of dynax/wrapper/envs/rc.py: the differentiable RC environment with discrete actions,
batched over N environments (the reference batches with jax.vmap, tests/test_gymwrapper_basics.py:81-168).

``env.step(action, states, params)`` is ONE launch of ``dynax_rc_env_step_f32`` (action -> control,
disturbance lookup, Euler step, observation, reward, termination), ``env.rollout`` K steps in one launch.
The reference loads its disturbance CSV at import time (rc.py:52-64, a file missing from the checkout);
here the table is a constructor argument.
"""
from __future__ import annotations

from dataclasses import dataclass, replace

import torch

from ... import ops

# rc.py:25-42
PARAMS = [10384.31640625, 499089.09375, 1321535.125, 1.5348844528198242, 0.5000327825546265, 1.000040054321289,
          20.119935989379883]
T_LOW = [12.0] * 8 + [22.0] * 10 + [12.0] * 6
T_HIGH = [30.0] * 8 + [26.0] * 10 + [30.0] * 6
PRICE = [0.02987] * 6 + [0.04667] * 6 + [0.15877] * 7 + [0.04667] * 3 + [0.02987] * 2
WEIGHTS = [100.0, 1.0]


@dataclass
class EnvStates:
    """x: [N,3] (or [3]) zone / wall temperatures, time: [N] (or scalar) seconds (rc.py:66-70)."""
    x: torch.Tensor
    time: torch.Tensor

    def update(self, **kw):
        return replace(self, **kw)


class RC:
    def __init__(self, disturbance_ts, disturbance_xs, start_time=0.0, end_time=None, dt=60.0, COP=2.5, price=PRICE,
                 T_low=T_LOW, T_high=T_HIGH, weights=WEIGHTS, u_high=0.0, u_low=-10.0, num_actions=101, name="RC"):
        f = lambda v: torch.as_tensor(v, dtype=torch.float32).cuda().contiguous()
        self.ts, self.xs = f(disturbance_ts), f(disturbance_xs)
        assert self.xs.shape == (self.ts.numel(), 4), "disturbance columns: out_temp, qint_lump, qwin_lump, qradin_lump"
        self.start_time, self.dt, self.COP = float(start_time), float(dt), float(COP)
        self.end_time = float(self.ts[-1]) if end_time is None else float(end_time)
        self.price, self.T_low, self.T_high = f(price), f(T_low), f(T_high)
        self.weights = [float(w) for w in weights]
        self.u_high, self.u_low, self.num_actions, self.name = float(u_high), float(u_low), int(num_actions), name

    def default_params(self):
        keys = ("Cai", "Cwe", "Cwi", "Re", "Ri", "Rw", "Rg")
        return {"params": {"simulator": {"model": dict(zip(keys, PARAMS))}}}          # rc.py:28-31

    def reset(self, key=None, params=None, states=None, batch=None):
        """obs, states, params = env.reset(key)   (rc.py:266-300; deterministic reset)."""
        params = params or self.default_params()
        if states is None:
            n = batch or 1
            x = torch.tensor([20.0, 30.0, 26.0], device="cuda").repeat(n, 1)
            states = EnvStates(x=x if batch else x[0], time=torch.full((n,), self.start_time, device="cuda") if batch
                               else torch.tensor(self.start_time, device="cuda"))
        return None, states, params

    def _theta(self, params):
        leaf = params["params"]["simulator"]["model"]
        return torch.tensor([float(leaf[k]) for k in ("Cai", "Cwe", "Cwi", "Re", "Ri", "Rw", "Rg")], device="cuda")

    def rollout(self, actions, states, params=None):
        """K env steps in one launch: actions[K,N] -> dict of per-step obs/reward/terminated + final states."""
        params = params or self.default_params()
        batched = states.x.dim() == 2
        x = states.x if batched else states.x[None]
        t = torch.as_tensor(states.time, dtype=torch.float32, device=x.device).reshape(-1).expand(x.shape[0]).contiguous()
        a = torch.as_tensor(actions, device=x.device)
        a = a.reshape(-1, x.shape[0]) if a.dim() < 2 else a
        out = ops.rc_env_step(self._theta(params), x.contiguous().float(), t, a, self.ts, self.xs, self.price, self.T_low,
                              self.T_high, self.dt, self.COP, self.weights[0], self.weights[1], self.u_low, self.u_high,
                              self.num_actions)
        out["states"] = EnvStates(x=out["x_next"] if batched else out["x_next"][0],
                                  time=out["t_next"] if batched else out["t_next"][0])
        return out

    def step(self, action, states, params=None):
        """obs_next, reward, terminated, truncated, info, states = env.step(action, states, params)  (rc.py:103-158)."""
        batched = states.x.dim() == 2
        out = self.rollout(torch.as_tensor(action).reshape(1, -1), states, params)
        sel = (lambda v: v[0]) if batched else (lambda v: v[0, 0])
        info = {"time": out["states"].time, "cost": sel(out["cost"]), "dTz": sel(out["dTz"]), "states": out["states"].x}
        return sel(out["obs"]), sel(out["reward"]), sel(out["terminated"]), False, info, out["states"]
