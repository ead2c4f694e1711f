"""Two VJP launches of the neural-ODE kernel on one wave of CTAs -- the command ncu profiles for
profiles/r02_node_vjp_tc_kernel_ncu.csv."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynax_b200 import ops
from tools import _synth as orc
N, T = 148 * 128, int(os.environ.get("PROF_T", 16))
W = [torch.tensor(w).cuda() for w in orc.synth_mlp(7)]
g = torch.Generator(device="cuda").manual_seed(0)
x0 = torch.randn((N, 3), device="cuda", generator=g)
u = torch.randn((T, N, 5), device="cuda", generator=g)
xs, ys, _, st = ops.node_rk4_forward(*W, x0, u, 0.25, emit_stages=True)
gx, gy = torch.randn(xs.shape, device="cuda", generator=g), torch.randn(ys.shape, device="cuda", generator=g)
for _ in range(2):
    out = ops.node_rk4_vjp(*W, x0, u, xs, 0.25, gx, gy, stage_x=st)
torch.cuda.synchronize()
print(float(out[0].abs().sum()))
