"""examples/3-inverse.py as a timing: epochs/s of the on-device calibration loop (loss+gradient + LAMB per epoch)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynax_b200 import ops
from dynax_b200.estimators import calibrate
from dynax_b200.models.RC import Continuous4R3C
from dynax_b200.simulators.simulator import DifferentiableSimulator
from tools import _synth as orc

T = 672                                       # the bundled four weeks of hourly data (examples/3-inverse.py)
for N in (1, 1024, 65536):
    rng = np.random.default_rng(0)
    u = orc.synth_u(T, 1, 3)[:, 0]            # one measured building: shared series
    hidden = orc.synth_theta(1, 5)
    # the "measured" zone temperature: a rollout of the hidden building by the product path itself
    _, y, _ = ops.rc_forward(torch.tensor(hidden).cuda(), torch.tensor(orc.synth_x0(1, 3)).cuda(),
                             torch.tensor(u[:, None, :]).cuda(), 3600.0)
    target = y[:, 0, 0].cpu().numpy()
    th0 = orc.synth_theta(N, 7)
    sim = DifferentiableSimulator(Continuous4R3C(), dt=3600.0)
    names = ["Cai", "Cwe", "Cwi", "Re", "Ri", "Rw", "Rg"]
    LB = {k: float(orc.RC_LB[j]) for j, k in enumerate(names)}
    UB = {k: float(orc.RC_UB[j]) for j, k in enumerate(names)}
    variables = {"params": {"model": {k: torch.tensor(th0[:, j]).cuda() for j, k in enumerate(names)}}}
    x0 = torch.tensor(np.repeat(orc.synth_x0(1, 3), N, 0)).cuda()
    uu = torch.tensor(u).cuda() if N > 1 else torch.tensor(u[:, None, :]).cuda()
    tg = torch.tensor(target).cuda() if N > 1 else torch.tensor(target[:, None]).cuda()
    for graph in (False, True):
        epochs = 300
        calibrate(sim, variables, x0, uu, tg, LB, UB, n_epochs=20, cuda_graph=graph)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        _, loss = calibrate(sim, variables, x0, uu, tg, LB, UB, n_epochs=epochs, cuda_graph=graph)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"N={N:6d} T={T} cuda_graph={graph!s:5s}: {dt/epochs*1e3:7.3f} ms/epoch  ({epochs/dt:8.0f} epochs/s, "
              f"5000 epochs = {5000*dt/epochs:6.2f} s)  final mean loss {loss.mean().item():.4f}", flush=True)
