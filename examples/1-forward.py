"""dynax examples/1-forward.py on the B200 engine: forward rollout of the 4R3C zone model over the bundled
28-day disturbance series (hourly means), single system -- then the same call for 100 000 buildings.

    python examples/1-forward.py            (needs a B200; build first: python dynax_b200/build.py)
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dynax_b200.models.RC import Continuous4R3C            # noqa: E402
from dynax_b200.simulators import DifferentiableSimulator  # noqa: E402

model = Continuous4R3C()
data = np.load(os.path.join(ROOT, "tests", "golden", "c1_hourly.npz"))   # examples/1-forward.py:19-29 pipeline
u_dt, y_dt = torch.tensor(data["u"]).cuda(), torch.tensor(data["y"]).cuda()
dt = 3600
state = torch.tensor([20.0, 30.0, 26.0]).cuda()

simulator = DifferentiableSimulator(model, dt)
params = [10384.31640625, 499089.09375, 1321535.125, 1.5348844528198242, 0.5000327825546265, 1.000040054321289,
          20.119935989379883]
rc = dict(zip(("Cai", "Cwe", "Cwi", "Re", "Ri", "Rw", "Rg"), params))
params_true = {"params": {"model": rc}}

simulator.apply(params_true, state, u_dt)                      # warm-up
torch.cuda.synchronize(); start = time.time()
states, outputs = simulator.apply(params_true, state, u_dt)
torch.cuda.synchronize(); end = time.time()
print("time elapsed for forward simulation: ", end - start)
print(states[0])
print(states.shape, outputs.shape)
print("rmse vs measured zone temperature: %.3f K" % float(((outputs[:, 0] - y_dt) ** 2).mean().sqrt()))

# the same model for a city: N buildings, per-building parameters and inputs, one launch
N = 100_000
g = torch.Generator(device="cuda").manual_seed(0)
theta = {k: torch.tensor(v, device="cuda") * torch.exp(torch.empty(N, device="cuda").uniform_(-0.2, 0.2, generator=g))
         for k, v in rc.items()}
x0 = state[None] + torch.empty((N, 3), device="cuda").uniform_(-1, 1, generator=g)
u = u_dt[:, None, :].repeat(1, N, 1) * (1 + 0.05 * torch.randn((u_dt.shape[0], N, 5), device="cuda", generator=g))
simulator.apply({"params": {"model": theta}}, x0, u)
torch.cuda.synchronize(); start = time.time()
xs, ys = simulator.apply({"params": {"model": theta}}, x0, u)
torch.cuda.synchronize(); end = time.time()
print("%d buildings x %d steps: %.2f ms (%.2e system-steps/s)" % (N, u.shape[0], 1e3 * (end - start), N * u.shape[0] / (end - start)))
