"""One launch of the MPC cost kernel for ncu (config 4: 4M candidates x 96 steps):
    ncu --set full --clock-control none --import-source on -k regex:rc_mpc_cost -s 1 -c 1 -o gpurun_out/r2_prof_mpc python tools/prof_mpc.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynax_b200 import ops
from tools import _synth as syn
dev = torch.device("cuda")
Nm, H = 1 << 22, 96
g = torch.Generator(device=dev).manual_seed(3456)
um = -10 * torch.rand((H, Nm), device=dev, generator=g)
d = torch.rand((H, 4), device=dev, generator=g); d[:, 0] = 15 + 10 * d[:, 0]
sch = [torch.tensor(x, dtype=torch.float32, device=dev) for x in (syn.RC_PRICE, syn.RC_T_LOW, syn.RC_T_HIGH)]
nom = torch.tensor(syn.RC_NOMINAL, dtype=torch.float32, device=dev); xm = torch.tensor([20.0, 35.8, 26.0], device=dev)
for _ in range(3):
    ops.rc_mpc_cost(nom, xm, d, um, *sch, 900.0)
torch.cuda.synchronize()
