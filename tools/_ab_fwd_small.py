"""Forward kernel on launches of at most one 128-system CTA per SM: out-tile count / chunk length (DYNAX_FWD_CFG)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynax_b200 import ops
from tools._synth import RC_NOMINAL
g = torch.Generator(device="cuda").manual_seed(0)
T = 8760
for N in (16384, 8192, 18944, 1024):
    th = torch.tensor(RC_NOMINAL, device="cuda", dtype=torch.float32)[None] * torch.exp(torch.empty((N, 7), device="cuda").uniform_(-0.2, 0.2, generator=g))
    x0 = torch.tensor([[20.0, 30.0, 26.0]], device="cuda").repeat(N, 1)
    u = torch.rand((T, N, 5), device="cuda", generator=g); u[..., 0] = 15 + 10 * u[..., 0]
    ref = None
    for cfg in ("4", "5", "4", "5"):
        os.environ["DYNAX_FWD_CFG"] = cfg
        fn = lambda: ops.rc_forward(th, x0, u, 3600.0, time_chunks=0)
        out = fn(); torch.cuda.synchronize()
        if ref is None: ref = [o.clone() for o in out[:2]]
        same = all(torch.equal(a, b) for a, b in zip(out[:2], ref))
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        print(f"N={N} cfg={cfg}: {e0.elapsed_time(e1)/10:.4f} ms  bitwise equal to cfg 4: {same}", flush=True)
    del u
