"""Diagnostic (not collected by pytest): element-wise gradient error of the CUDA path and of the fp32 C oracle against the
fp64 oracle, per case of tests/test_gpu_rc_grad.py.  Run on a GPU box: python tests/diag_grad.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynax_b200 import ops
from oracle import c_oracle, dynax_oracle as orc
from tests.test_gpu_rc_grad import make_case

def cu(a): return torch.tensor(np.ascontiguousarray(a, dtype=np.float32)).cuda()

for (N, Tn, noise, kind) in [(256, 8760, 0.1, "hidden"), (256, 8760, 1.0, "hidden"), (256, 8760, None, "flat22"), (130, 1000, 0.1, "hidden")]:
    if kind == "hidden":
        th, x0, u, target = make_case(N, Tn, 31, noise)
    else:
        th, x0, u = orc.synth_theta(N, 31), orc.synth_x0(N, 31), orc.synth_u(Tn, N, 31)
        target = (22 + np.random.default_rng(0).normal(size=(Tn, N))).astype(np.float32)
    loss, grad, _ = ops.rc_loss_grad(cu(th), cu(x0), cu(u), cu(target), 3600.0)
    l64, g64 = orc.rc_loss_grad(th, x0, u, target, 3600.0, None, None, np.float64)
    l32, g32 = c_oracle.rc_loss_grad(th, x0, u, target, 3600.0)
    g = grad.cpu().numpy().astype(np.float64); l = loss.cpu().numpy().astype(np.float64)
    y64 = orc.rc_rollout(th, x0, u, 3600.0, np.float64)[1][..., 0]
    r = y64 - target
    nrm_o = np.linalg.norm(g - g64, axis=1) / np.linalg.norm(g64, axis=1)
    nrm_r = np.linalg.norm(g32 - g64, axis=1) / np.linalg.norm(g64, axis=1)
    el_o = np.abs(g - g64) / np.abs(g64); el_r = np.abs(g32 - g64) / np.abs(g64)
    print(f"--- N={N} T={Tn} noise={noise} {kind}: mean|r|={np.abs(r).mean():.3f} loss~{l64.mean():.3f}")
    print(f" loss rel err ours max {np.abs(l/l64-1).max():.2e} med {np.median(np.abs(l/l64-1)):.2e} | ref32 max {np.abs(l32/l64-1).max():.2e} med {np.median(np.abs(l32/l64-1)):.2e}")
    print(f" loss bound from traj tol: max ratio {(np.abs(l-l64)/(2*np.abs(r).mean(0)*1e-5*np.abs(y64).max(0))).max():.3f}")
    print(f" grad normwise ours max {nrm_o.max():.2e} med {np.median(nrm_o):.2e} | ref32 max {nrm_r.max():.2e} med {np.median(nrm_r):.2e}")
    print(f" grad per-elem ours max {el_o.max():.2e} p99 {np.quantile(el_o,0.99):.2e} med {np.median(el_o):.2e} | ref32 max {el_r.max():.2e} p99 {np.quantile(el_r,0.99):.2e} med {np.median(el_r):.2e}")
    print(" per-param max ours", np.array2string(el_o.max(0), precision=1), "\n per-param max ref ", np.array2string(el_r.max(0), precision=1))
    worse = (np.abs(g - g64) > np.maximum(1e-4*np.abs(g64), np.abs(g32-g64))).mean()
    print(f" frac elems failing H2 per-elem rule: {worse:.4f}")
