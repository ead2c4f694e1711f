"""A/B timing of the fused loss+gradient (tangent kernel) between two builds of the library (see tools/_ab_fwd.py for how
the second build gets to dynax_b200/_old/; one build per process -- two copies of the library in one process crashed).  Prints the best single launch, the mean of 30 launches back to back (the
power-capped steady state) and fp64 sums of loss and gradient (equal sums = same results)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynax_b200 import ops, _lib
from tools._synth import RC_NOMINAL
name = os.environ.get("AB_NAME", "new")   # one build per process: DYNAX_B200_LIB=dynax_b200/_old/libdynax_old.so AB_NAME=old python tools/_ab_tangent.py
g = torch.Generator(device="cuda").manual_seed(0)
T = 8760
for N in (227328, 65536, 16384):
    th = torch.tensor(RC_NOMINAL, device="cuda", dtype=torch.float32)[None] * torch.exp(torch.empty((N, 7), device="cuda").uniform_(-0.2, 0.2, generator=g))
    x0 = torch.tensor([[20.0, 30.0, 26.0]], device="cuda").repeat(N, 1)
    u = torch.rand((T, N, 5), device="cuda", generator=g); u[..., 0] = 15 + 10 * u[..., 0]
    tg = 20 + 2 * torch.rand((T, N), device="cuda", generator=g)
    fn = lambda: ops.rc_loss_grad(th, x0, u, tg, 3600.0, time_chunks=0)
    for rep in range(2):
        if True:
            for _ in range(3): fn()
            torch.cuda.synchronize()
            best = 1e9
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(30): out = fn()
            e1.record(); torch.cuda.synchronize()
            loss, grad = out[0], out[1]
            print(f"N={N:6d} {name}: best {best:7.3f} ms  sustained {e0.elapsed_time(e1)/30:7.3f} ms  "
                  f"chk {float(loss.double().sum()):.12e} {float(grad.double().abs().sum()):.12e}", flush=True)
    del u, tg
