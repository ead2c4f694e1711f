"""General-cotangent RC VJP and dense LTI VJP on launches below one wave, per-system vs shared input series."""
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
for N in (128, 8192, 18944):
    th = torch.tensor(RC_NOMINAL, device="cuda", dtype=torch.float32)[None] * torch.exp(torch.empty((N, 7), device="cuda").uniform_(-0.2, 0.2, generator=g))
    x0 = torch.tensor([[20.0, 30.0, 26.0]], device="cuda").repeat(N, 1)
    u = torch.rand((T, N, 5), device="cuda", generator=g); u[..., 0] = 15 + 10 * u[..., 0]
    us = u[:, 0].contiguous()
    gy = torch.randn((T, N, 1), device="cuda", generator=g)
    print(f"N={N} rc_vjp per-system u: {t(lambda: ops.rc_vjp(th, x0, u, 3600.0, None, gy, need_u=False, time_chunks=0)):.4f} ms   shared u: {t(lambda: ops.rc_vjp(th, x0, us, 3600.0, None, gy, need_u=False, time_chunks=0)):.4f}  shared u + g_u: {t(lambda: ops.rc_vjp(th, x0, us, 3600.0, None, gy, need_u=True, time_chunks=0)):.4f} ms", flush=True)
    os.environ["DYNAX_VJP_CFG"] = "1"
    print(f"N={N} (8-step segments forced) shared u: {t(lambda: ops.rc_vjp(th, x0, us, 3600.0, None, gy, need_u=False, time_chunks=0)):.4f}  shared u + g_u: {t(lambda: ops.rc_vjp(th, x0, us, 3600.0, None, gy, need_u=True, time_chunks=0)):.4f} ms", flush=True)
    os.environ.pop("DYNAX_VJP_CFG")
