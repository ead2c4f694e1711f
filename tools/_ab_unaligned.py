"""A/B of the launches whose rows TMA cannot move (N % 4 != 0) between two builds of the library (tools/_ab_fwd.py says how
the second build gets to dynax_b200/_old/): forward x+y, fused loss+gradient (tangent and adjoint kernels), general VJP.
Prints ms, fraction of the measured HBM copy peak (algorithmic bytes) and fp64 checksums (equal sums = same results)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynax_b200 import ops, _lib
from tools._synth import RC_NOMINAL, RC_LB, RC_UB
PEAK = 6549.4e9
name = os.environ.get("AB_NAME", "new")   # one build per process: DYNAX_B200_LIB=... AB_NAME=old python tools/_ab_unaligned.py
def best(fn, it=6):
    for _ in range(2): fn()
    torch.cuda.synchronize(); b = 1e9
    for _ in range(it):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); b = min(b, e0.elapsed_time(e1))
    return b
g = torch.Generator(device="cuda").manual_seed(0)
T = 2000
for N in [int(x) for x in os.environ.get('AB_N', '151553,151552,75777').split(',')]:
    th = torch.tensor(RC_NOMINAL, device="cuda", dtype=torch.float32)[None] * torch.exp(torch.empty((N, 7), device="cuda").uniform_(-0.2, 0.2, generator=g))
    x0 = torch.tensor([[20.0, 30.0, 26.0]], device="cuda").repeat(N, 1)
    u = torch.rand((T, N, 5), device="cuda", generator=g); u[..., 0] = 15 + 10 * u[..., 0]
    tg = 20 + 2 * torch.rand((T, N), device="cuda", generator=g)
    gy = torch.randn((T, N, 1), device="cuda", generator=g)
    lb, ub = torch.tensor(RC_LB, device="cuda", dtype=torch.float32), torch.tensor(RC_UB, device="cuda", dtype=torch.float32)
    if True:
        os.environ.pop("DYNAX_GRAD_ALGO", None)
        a = best(lambda: ops.rc_forward(th, x0, u, 3600.0, time_chunks=0))
        xs, ys, _ = ops.rc_forward(th, x0, u, 3600.0, time_chunks=0)
        b = best(lambda: ops.rc_loss_grad(th, x0, u, tg, 3600.0, lb, ub, time_chunks=0))
        l1, g1, _ = ops.rc_loss_grad(th, x0, u, tg, 3600.0, lb, ub, time_chunks=0)
        os.environ["DYNAX_GRAD_ALGO"] = "0"
        c = best(lambda: ops.rc_loss_grad(th, x0, u, tg, 3600.0, lb, ub, time_chunks=0))
        l2, g2, _ = ops.rc_loss_grad(th, x0, u, tg, 3600.0, lb, ub, time_chunks=0)
        os.environ.pop("DYNAX_GRAD_ALGO")
        d = best(lambda: ops.rc_vjp(th, x0, u, 3600.0, None, gy, need_u=True, time_chunks=0))
        v = ops.rc_vjp(th, x0, u, 3600.0, None, gy, need_u=True, time_chunks=0)
        fr = lambda ms, B: N * T * B / (ms * 1e-3) / PEAK
        print(f"N={N} T={T} {name}: fwd x+y {a:6.3f} ms ({fr(a,36):.2f}) | tangent {b:6.3f} ms ({fr(b,24):.2f}) | adjoint {c:6.3f} ms ({fr(c,48):.2f}) | "
              f"vjp+g_u {d:6.3f} ms ({fr(d,68):.2f})  chk {float(xs.double().sum()):.9e} {float(ys.double().sum()):.9e} "
              f"{float(l1.double().sum()):.9e} {float(g1.double().abs().sum()):.9e} {float(l2.double().sum()):.9e} "
              f"{float(g2.double().abs().sum()):.9e} {float(v[0].double().abs().sum()):.9e} {float(v[2].double().abs().sum()):.9e}", flush=True)
    del u, tg, gy
