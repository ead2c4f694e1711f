"""General-cotangent RC VJP: serial kernel vs time-parallel chunked adjoint around the N threshold (ops.rc_vjp: N <= 4096)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynax_b200 import ops
from tools._synth import RC_NOMINAL
g = torch.Generator(device="cuda").manual_seed(0)
T = 8760
def t(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10
for N in (1024, 4096, 6144, 8192, 12288, 16384):
    th = torch.tensor(RC_NOMINAL, device="cuda", dtype=torch.float32)[None] * torch.exp(torch.empty((N, 7), device="cuda").uniform_(-0.2, 0.2, generator=g))
    x0 = torch.tensor([[20.0, 30.0, 26.0]], device="cuda").repeat(N, 1)
    u = torch.rand((T, N, 5), device="cuda", generator=g); u[..., 0] = 15 + 10 * u[..., 0]
    gy = torch.randn((T, N, 1), device="cuda", generator=g)
    line = f"N={N}: serial {t(lambda: ops.rc_vjp(th, x0, u, 3600.0, None, gy, time_chunks=0)):.3f} ms"
    for ch in (-1, 2, 4, 8, 16):
        try:
            line += f"  chunks={ch}: {t(lambda: ops.rc_vjp(th, x0, u, 3600.0, None, gy, time_chunks=ch)):.3f}"
        except Exception as e:
            line += f"  chunks={ch}: {str(e)[:40]}"
    print(line, flush=True)
    del u, gy
