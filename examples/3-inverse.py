"""dynax examples/3-inverse.py on the B200 engine: inverse parameter inference of the 4R3C model by
gradient descent (optax.lamb in the reference) on the MSE + box-penalty loss -- here for a whole population of
starting points at once, every epoch = one fused loss+gradient launch + one update launch on the device.

    python examples/3-inverse.py [n_epochs] [n_candidates]
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dynax_b200.estimators import calibrate, mse_loss_and_grad   # noqa: E402
from dynax_b200.models.RC import Continuous4R3C, PARAM_NAMES      # noqa: E402
from dynax_b200.simulators import DifferentiableSimulator         # noqa: E402

n_epochs = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 256

model = Continuous4R3C()
data = np.load(os.path.join(ROOT, "tests", "golden", "c1_hourly.npz"))
dt = 3600
n_train = int(len(data["u"]) * 0.75)                              # 3-inverse.py:37-42
u_train = torch.tensor(data["u"][:n_train]).cuda()
y_train = torch.tensor(data["y"][:n_train]).cuda()
state = torch.tensor([float(y_train[0]), 36.0, 25.0]).cuda()      # 3-inverse.py:51
simulator = DifferentiableSimulator(model, dt)

lb = dict(zip(PARAM_NAMES, [3.0e3, 5.0e4, 5.0e5, 1.0, 0.5, 1.0, 5.0]))          # 3-inverse.py:66-85
ub = dict(zip(PARAM_NAMES, [3.0e4, 5.0e5, 5.0e6, 1.0e1, 5.0, 1.0e1, 5.0e1]))
params_lb = {"params": {"model": lb}}
params_ub = {"params": {"model": ub}}

# the reference starts ONE run at the upper bounds; we start N candidates spread over the box
g = torch.Generator(device="cuda").manual_seed(0)
start = {k: torch.tensor(lb[k], device="cuda") + torch.rand(N, device="cuda", generator=g) * (ub[k] - lb[k]) for k in PARAM_NAMES}
for k in PARAM_NAMES:
    start[k][0] = ub[k]                                           # candidate 0 == the reference's initial guess
params0 = {"params": {"model": start}}
x0 = state[None].repeat(N, 1)

loss0, _ = mse_loss_and_grad(simulator, params0, x0, u_train, y_train, params_lb, params_ub)
fitted, loss = calibrate(simulator, params0, x0, u_train, y_train, params_lb, params_ub, n_epochs=n_epochs, lr=1e-3,
                         log_every=max(1, n_epochs // 5))
best = int(torch.argmin(loss))
print("initial loss (reference start) %.4f -> %.4f after %d epochs; best of %d candidates %.4f" %
      (float(loss0[0]), float(loss[0]), n_epochs, N, float(loss[best])))
trained = {k: float(fitted["params"]["model"][k][best]) for k in PARAM_NAMES}
print(json.dumps(trained, indent=1))
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "zone_coefficients.json"), "w") as f:
    json.dump(trained, f)                                         # 3-inverse.py:141-146
