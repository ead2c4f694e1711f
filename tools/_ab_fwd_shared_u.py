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
for N in (128, 8192, 18944, 65536):
    th = torch.tensor(RC_NOMINAL, device="cuda", dtype=torch.float32)[None] * torch.exp(torch.empty((N, 7), device="cuda").uniform_(-0.2, 0.2, generator=g))
    x0 = torch.tensor([[20.0, 30.0, 26.0]], device="cuda").repeat(N, 1)
    us = torch.rand((T, 5), device="cuda", generator=g); us[..., 0] = 15 + 10 * us[..., 0]
    ctrl = -5 * torch.rand((T, N), device="cuda", generator=g)
    ref = None
    for cfg in ("0", "-1"):   # 0 = the short chunks everywhere, -1 = the library default
        os.environ["DYNAX_FWD_CFG"] = cfg
        out = ops.rc_forward(th, x0, us, 3600.0, time_chunks=0); torch.cuda.synchronize()
        if ref is None: ref = [o.clone() for o in out[:2]]
        same = all(torch.equal(a, b) for a, b in zip(out[:2], ref))
        print(f"N={N} shared u cfg={cfg}: x+y {t(lambda: ops.rc_forward(th, x0, us, 3600.0, time_chunks=0)):.4f}  final only {t(lambda: ops.rc_forward(th, x0, us, 3600.0, emit_x=False, emit_y=False, emit_final=True, time_chunks=0)):.4f}  ctrl x+y {t(lambda: ops.rc_forward_ctrl(th, x0, us, ctrl, 3600.0)):.4f} ms  same {same}", flush=True)
