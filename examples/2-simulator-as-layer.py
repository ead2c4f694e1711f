"""dynax examples/2-simulator-as-layer.py on the B200 engine: the simulator used as a layer of a larger module.

The reference wraps ``DifferentiableSimulator`` in a Flax module ``Forward`` whose parameter tree gains one
level (``{'params': {'simulator': {'model': rc}}}``, examples/2-simulator-as-layer.py:14-19,62).  The same
wrapper here: a plain Python class that owns the simulator, exposes ``init``/``apply`` and unwraps that level.
Then it is made a *trainable* layer: a torch ``nn.Module`` whose parameters are the seven RC coefficients, with
gradients flowing through the CUDA rollout's custom VJP.

    python examples/2-simulator-as-layer.py            (needs a B200; build first: python dynax_b200/build.py)
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dynax_b200 import ops                                   # noqa: E402
from dynax_b200.models.RC import Continuous4R3C            # noqa: E402
from dynax_b200.simulators import DifferentiableSimulator  # noqa: E402


class Forward:
    """Forward-simulation wrapper: the simulator is a sub-module named ``simulator``."""

    def __init__(self, simulator):
        self.simulator = simulator

    def init(self, key, states, inputs):
        return {"params": {"simulator": self.simulator.init(key, states, inputs)["params"]}}

    def apply(self, variables, states, inputs):
        return self.simulator.apply({"params": variables["params"]["simulator"]}, states, inputs)


dt = 3600
data = np.load(os.path.join(ROOT, "tests", "golden", "c1_hourly.npz"))   # the example's resampling pipeline, hourly means
u_dt, y_dt = torch.tensor(data["u"]).cuda(), torch.tensor(data["y"]).cuda()
model = Continuous4R3C()
simulator = DifferentiableSimulator(model, dt=dt, mode_interp="linear", start_time=0)
forward = Forward(simulator)
state = torch.tensor([20.0, 30.0, 26.0]).cuda()
params = [10384.31640625, 499089.09375, 1321535.125, 1.5348844528198242, 0.5000327825546265, 1.000040054321289,
          20.119935989379883]
names = ("Cai", "Cwe", "Cwi", "Re", "Ri", "Rw", "Rg")
params_true = {"params": {"simulator": {"model": dict(zip(names, params))}}}
tree = forward.init(0, state, u_dt)
assert set(tree["params"]["simulator"]["model"]) == set(names)

forward.apply(params_true, state, u_dt)
torch.cuda.synchronize(); start = time.time()
states, outputs = forward.apply(params_true, state, u_dt)
torch.cuda.synchronize(); end = time.time()
print("time elapsed for forward simulation: ", end - start)
print(states[0])
print(states.shape, outputs.shape)


class SimulatorLayer(torch.nn.Module):
    """The simulator as a trainable layer: log-parameters so that the coefficients stay positive."""

    def __init__(self, theta0, dt):
        super().__init__()
        self.log_theta = torch.nn.Parameter(torch.log(torch.tensor(theta0, dtype=torch.float32)))
        self.dt = dt

    def forward(self, x0, u):                       # x0[N,3], u[T,N,5] -> ysol[T,N,1]
        theta = torch.exp(self.log_theta)[None].expand(x0.shape[0], 7)
        return ops.rc_rollout(theta, x0, u, self.dt)[1]


layer = SimulatorLayer(np.array(params) * np.exp(np.random.default_rng(0).uniform(-0.3, 0.3, 7)), float(dt)).cuda()
opt = torch.optim.Adam(layer.parameters(), lr=2e-2)
x0, u = state[None], u_dt[:, None, :].contiguous()
losses = []
for epoch in range(60):
    opt.zero_grad()
    loss = ((layer(x0, u)[:, 0, 0] - y_dt) ** 2).mean()
    loss.backward()
    opt.step()
    losses.append(loss.item())
print("layer training: initial loss %.4f -> final loss %.4f" % (losses[0], losses[-1]))
