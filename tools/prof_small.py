"""Five fused loss+gradient launches on a launch that cannot fill the GPU (N=8192 systems, T=8760, one shared theta:
BASELINE.json config 3 sharded over 8 GPUs) -- the command profiled by ncu for profiles/r02_tangent_split_kernel_ncu.csv."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynax_b200 import ops
from tools._synth import RC_NOMINAL
N, T = int(os.environ.get("PROF_N", 8192)), 8760
g = torch.Generator(device="cuda").manual_seed(0)
th = torch.tensor(RC_NOMINAL, device="cuda", dtype=torch.float32)
x0 = torch.tensor([[20.0, 30.0, 26.0]], device="cuda").repeat(N, 1)
u = torch.rand((T, N, 5), device="cuda", generator=g); u[..., 0] = 15 + 10 * u[..., 0]
tg = 20 + 2 * torch.rand((T, N), device="cuda", generator=g)
for _ in range(5):
    loss, grad, _ = ops.rc_loss_grad(th, x0, u, tg, 3600.0, time_chunks=0)
torch.cuda.synchronize()
print(float(loss), grad.tolist())
