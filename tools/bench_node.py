"""Config 5 timing: MLP-dynamics RK4 rollout, N=262144 trajectories, T steps (scratch tool)."""
import argparse, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynax_b200 import ops
from tools import _synth as orc

ap = argparse.ArgumentParser(); ap.add_argument("--N", type=int, default=262144); ap.add_argument("--T", type=int, default=256)
ap.add_argument("--impl", default=None); ap.add_argument("--hidden", type=int, default=64)
a = ap.parse_args()
if a.impl is not None:
    os.environ["DYNAX_NODE_IMPL"] = a.impl
W = [torch.tensor(w).cuda() for w in orc.synth_mlp(7, hidden=a.hidden)]
g = torch.Generator(device="cuda").manual_seed(0)
x0 = torch.randn((a.N, 3), device="cuda", generator=g)
u = torch.randn((a.T, a.N, 5), device="cuda", generator=g)
for it in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); xs, ys, _ = ops.node_rk4_forward(*W, x0, u, 0.25); e1.record(); torch.cuda.synchronize()
    s = e0.elapsed_time(e1) * 1e-3
    flop = 2 * (8 * a.hidden + a.hidden * a.hidden + 3 * a.hidden) * 4
    print(f"node rk4 hidden={a.hidden} N={a.N} T={a.T}: {s*1e3:.2f} ms  {a.N*a.T/s:.3e} steps/s  {a.N*a.T*flop/s/1e12:.1f} TFLOP/s(MLP)", flush=True)
# parity of this kernel against the fp64 restatement: tests/test_gpu_node.py
# --- VJP timing: see tools/bench_node_vjp.py (tcgen05 kernel vs CUDA-core kernel)
