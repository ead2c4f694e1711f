"""Config 5 gradient timing: VJP of the MLP-dynamics RK4 rollout, tcgen05 kernel vs CUDA-core kernel (scratch tool)."""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynax_b200 import ops
from tools import _synth as orc

ap = argparse.ArgumentParser(); ap.add_argument("--N", type=int, default=148 * 128 * 2); ap.add_argument("--T", type=int, default=64)
ap.add_argument("--impls", default="1,0")
a = ap.parse_args()
W = [torch.tensor(w).cuda() for w in orc.synth_mlp(7)]
g = torch.Generator(device="cuda").manual_seed(0)
x0 = torch.randn((a.N, 3), device="cuda", generator=g)
u = torch.randn((a.T, a.N, 5), device="cuda", generator=g)
xs, ys, _, stages = ops.node_rk4_forward(*W, x0, u, 0.25, emit_stages=True)
gx, gy = torch.randn(xs.shape, device="cuda", generator=g), torch.randn(ys.shape, device="cuda", generator=g)
outs = {}
for impl in a.impls.split(","):
    os.environ["DYNAX_NODE_VJP_IMPL"] = impl
    for st, label in ((None, "stage points recomputed"), (stages, "stage points saved by the forward kernel")):
        for it in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); out = ops.node_rk4_vjp(*W, x0, u, xs, 0.25, gx, gy, stage_x=st); e1.record(); torch.cuda.synchronize()
            s = e0.elapsed_time(e1) * 1e-3
        print(f"node rk4 VJP impl={'tcgen05' if impl != '0' else 'cuda-core'} N={a.N} T={a.T} ({label}): {s*1e3:.2f} ms  "
              f"{a.N*a.T/s:.3e} steps/s", flush=True)
    outs[impl] = out
if len(outs) == 2:
    for nm, p, q in zip(["gW1", "gb1", "gW2", "gb2", "gW3", "gb3", "g_x0", "g_u"], outs["1"], outs["0"]):
        print(f"  {nm}: max|tc - cuda_core| / max|cuda_core| = {(p - q).abs().max().item() / q.abs().max().item():.2e}")
