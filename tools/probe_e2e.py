"""Scratch probe: host-API loss+gradient vs tile size, and the raw pinned H2D copy bandwidth of this box."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynax_b200 import host

T, Ne = 8760, 32768
hu = torch.empty((T, Ne, 5), dtype=torch.float32, pin_memory=True).uniform_(0, 1)
ht = torch.empty((T, Ne), dtype=torch.float32, pin_memory=True).uniform_(20, 22)
hth = torch.tensor([10384.3, 499089.1, 1321535.1, 1.5348, 0.5, 1.0, 20.12]).repeat(Ne, 1).pin_memory()
hx0 = torch.tensor([20.0, 30.0, 26.0]).repeat(Ne, 1).pin_memory()
d = torch.empty((T * Ne * 5,), dtype=torch.float32, device="cuda")
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    d.copy_(hu.view(-1), non_blocking=True); torch.cuda.synchronize()
    dt = time.perf_counter() - t0
print(f"raw contiguous pinned H2D: {hu.numel()*4/dt/1e9:.1f} GB/s", flush=True)
del d
args = (hth.numpy(), hx0.numpy(), hu.numpy(), ht.numpy(), 3600.0)
for tile in (0, 2048, 4096, 8192, 16384, 32768):
    host.rc_loss_grad(*args, tile_systems=tile)
    t0 = time.perf_counter()
    for _ in range(2):
        host.rc_loss_grad(*args, tile_systems=tile)
    dt = (time.perf_counter() - t0) / 2
    print(f"tile {tile:6d}: {dt*1e3:7.1f} ms  {Ne*T/dt:.3e} steps/s  {Ne*T*24/dt/1e9:.1f} GB/s H2D", flush=True)
