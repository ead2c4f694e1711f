"""fwd+grad with shared disturbances + per-system control channel (row f-1): device and end-to-end."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynax_b200 import ops
NOM = [10384.31640625, 499089.09375, 1321535.125, 1.5348844528198242, 0.5000327825546265, 1.000040054321289, 20.119935989379883]
N, T = 148 * 2 * 128 * 3, 8760
g = torch.Generator(device="cuda").manual_seed(0)
th = torch.tensor(NOM, device="cuda")[None] * torch.exp(torch.empty((N, 7), device="cuda").uniform_(-0.2, 0.2, generator=g))
x0 = torch.tensor([[20.0, 30.0, 26.0]], device="cuda") + torch.empty((N, 3), device="cuda").uniform_(-1, 1, generator=g)
d = torch.rand((T, 5), device="cuda", generator=g); d[:, 0] = 15 + 10 * d[:, 0]
ctrl = -5 * torch.rand((T, N), device="cuda", generator=g)
target = 20 + 2 * torch.rand((T, N), device="cuda", generator=g)
for name, fn, B in (("fwd+grad ctrl, per-system target", lambda: ops.rc_loss_grad_ctrl(th, x0, d, ctrl, target, 3600.0), 16),
                    ("fwd+grad ctrl, shared target", lambda: ops.rc_loss_grad_ctrl(th, x0, d, ctrl, target[:, 0].contiguous(), 3600.0), 8),
                    ("forward ctrl, y only", lambda: ops.rc_forward_ctrl(th, x0, d, ctrl, 3600.0, emit_x=False), 8)):
    for _ in range(2): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): fn()
    e1.record(); torch.cuda.synchronize()
    s = e0.elapsed_time(e1) / 3 * 1e-3
    print(f"{name:36s} N={N} T={T}: {s*1e3:8.2f} ms  {N*T/s:.3e} steps/s  {N*T*B/s/1e9:7.1f} GB/s ({B} B/step)", flush=True)
