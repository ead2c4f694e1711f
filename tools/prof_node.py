"""Three forward launches of the neural-ODE kernel (config 5 shape, short horizon) -- the command ncu profiles for
profiles/r02_node_kernel_ncu.csv."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynax_b200 import ops
from tools import _synth as orc
N, T, H = int(os.environ.get("PROF_N", 262144)), int(os.environ.get("PROF_T", 64)), int(os.environ.get("PROF_H", 64))
W = [torch.tensor(w).cuda() for w in orc.synth_mlp(7, hidden=H)]
g = torch.Generator(device="cuda").manual_seed(0)
x0 = torch.randn((N, 3), device="cuda", generator=g)
u = torch.randn((T, N, 5), device="cuda", generator=g)
for _ in range(3):
    xs, ys, _ = ops.node_rk4_forward(*W, x0, u, 0.25)
torch.cuda.synchronize()
print(float(xs.abs().max()))
