"""Small-N / long-T forward rollout: serial-in-time kernel vs the time-parallel chunked kernel."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynax_b200 import ops
NOM = [10384.31640625, 499089.09375, 1321535.125, 1.5348844528198242, 0.5000327825546265, 1.000040054321289, 20.119935989379883]
T = 8760
for N in (1, 128, 1024, 8192, 32768):
    th = torch.tensor(NOM, device="cuda")[None].repeat(N, 1)
    x0 = torch.tensor([[20.0, 30.0, 26.0]], device="cuda").repeat(N, 1)
    u = torch.rand((T, N, 5), device="cuda")
    for name, tc in (("serial", 0), ("chunked", -1)):
        for _ in range(2):
            ops.rc_forward(th, x0, u, 3600.0, time_chunks=tc)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ops.rc_forward(th, x0, u, 3600.0, time_chunks=tc)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"N={N:6d} T={T} {name:8s} {ms:8.3f} ms  {N*T/ms*1e3:.3e} steps/s", flush=True)

print("--- loss+grad")
for N in (1, 128, 512, 1024, 2048, 4096, 8192):
    th = torch.tensor(NOM, device="cuda")[None].repeat(N, 1)
    x0 = torch.tensor([[20.0, 30.0, 26.0]], device="cuda").repeat(N, 1)
    u = torch.rand((T, N, 5), device="cuda"); tg = 20 + torch.rand((T, N), device="cuda")
    for name, tc in (("serial", 0), ("chunked", -1)):
        for _ in range(2):
            ops.rc_loss_grad(th, x0, u, tg, 3600.0, time_chunks=tc)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ops.rc_loss_grad(th, x0, u, tg, 3600.0, time_chunks=tc)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"N={N:6d} T={T} loss+grad {name:8s} {ms:8.3f} ms  {N*T/ms*1e3:.3e} steps/s", flush=True)
