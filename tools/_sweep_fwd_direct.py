"""Aligned launches of the forward kernel: TMA out-tile vs per-thread stores (DYNAX_FWD_DIRECT), per tile shape."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynax_b200 import ops
from tools._synth import RC_NOMINAL
def t(fn, it=6):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(it):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
g = torch.Generator(device="cuda").manual_seed(0)
T = 8760
for N in (1024, 4096, 16384, 18944, 32768, 37888, 56832, 75776, 151552):
    th = torch.tensor(RC_NOMINAL, device="cuda", dtype=torch.float32)[None].repeat(N, 1)
    x0 = torch.tensor([[20.0, 30.0, 26.0]], device="cuda").repeat(N, 1)
    u = torch.rand((T, N, 5), device="cuda", generator=g)
    for cfg in (-1, 0, 1, 3, 4):
        row = []
        for direct in (0, 1):
            os.environ["DYNAX_FWD_CFG"] = str(cfg); os.environ["DYNAX_FWD_DIRECT"] = str(direct)
            a = t(lambda: ops.rc_forward(th, x0, u, 3600.0, time_chunks=0))
            b = t(lambda: ops.rc_forward(th, x0, u, 3600.0, emit_x=False, time_chunks=0))
            row.append((a, b))
        print(f"N={N:6d} cfg={cfg:2d}  x+y tma {row[0][0]:7.3f} direct {row[1][0]:7.3f} | y tma {row[0][1]:7.3f} direct {row[1][1]:7.3f} ms", flush=True)
    del u
