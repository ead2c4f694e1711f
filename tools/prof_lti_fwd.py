"""One launch of the dense (4,3,4) forward kernel (x and y emitted) for ncu, N = 113 664 x T = 1024:
    ncu --set full --clock-control none --import-source on -k regex:rollout_fwd -s 1 -c 1 -o gpurun_out/r2_prof_lti_fwd python tools/prof_lti_fwd.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynax_b200 import ops
g = torch.Generator(device="cuda").manual_seed(0)
n, m, p, N, T = 4, 3, 4, 148 * 3 * 128 * 2, 1024
A = -0.05 * torch.eye(n, device="cuda") + 0.01 * torch.randn((n, n), device="cuda", generator=g)
B = 0.1 * torch.randn((n, m), device="cuda", generator=g)
C = torch.randn((p, n), device="cuda", generator=g); D = 0.1 * torch.randn((p, m), device="cuda", generator=g)
x0 = torch.randn((N, n), device="cuda", generator=g)
u = torch.randn((T, N, m), device="cuda", generator=g)
for _ in range(3):
    ops.lti_forward(A, B, C, D, x0, u, 0.1)
torch.cuda.synchronize()
