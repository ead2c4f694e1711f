"""Tiny invocation of every kernel family, meant to run under `compute-sanitizer --tool memcheck`
(one tool per gpurun call; see B200_PROFILING.md)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynax_b200 import ops
from tools import _synth as orc

c = lambda a: torch.tensor(np.ascontiguousarray(a, dtype=np.float32)).cuda()
for N, T in ((132, 37), (131, 20)):      # TMA path with a ragged tail tile / plain-load path
    th, x0, u = orc.synth_theta(N, 1), orc.synth_x0(N, 1), orc.synth_u(T, N, 1)
    tg = (22 + np.zeros((T, N))).astype(np.float32)
    ops.rc_forward(c(th), c(x0), c(u), 3600.0, emit_final=True, time_chunks=0)
    ops.rc_forward(c(th), c(x0), c(u), 3600.0, time_chunks=3)
    ops.rc_forward(c(th[0]), c(x0), c(u[:, 0]), 3600.0, time_chunks=0)
    ops.rc_loss_grad(c(th), c(x0), c(u), c(tg), 3600.0, orc.RC_LB, orc.RC_UB, want_gx0=True, time_chunks=0)
    ops.rc_loss_grad(c(th), c(x0), c(u), c(tg), 3600.0, time_chunks=3)
    ops.rc_loss_grad(c(th[0]), c(x0), c(u[:, 0]), c(tg[:, 0]), 3600.0, time_chunks=0)
    ops.rc_vjp(c(th), c(x0), c(u), 3600.0, c(np.ones((T, N, 3))), c(np.ones((T, N, 1))))
    ops.rc_forward_ctrl(c(th), c(x0), c(u[:, 0]), c(u[:, :, 2]), 3600.0)
    ops.rc_loss_grad_ctrl(c(th), c(x0), c(u[:, 0]), c(u[:, :, 2]), c(tg), 3600.0)
    rng = np.random.default_rng(0)
    A = -np.eye(4)[None] + 0.1 * rng.normal(size=(N, 4, 4)); B = rng.normal(size=(N, 4, 3))
    C = rng.normal(size=(N, 4, 4)); D = rng.normal(size=(N, 4, 3))
    xl, ul = rng.normal(size=(N, 4)), rng.normal(size=(T, N, 3))
    ops.lti_forward(c(A), c(B), c(C), c(D), c(xl), c(ul), 0.05)
    ops.lti_vjp(c(A), c(B), c(C), c(D), c(xl), c(ul), 0.05, c(np.ones((T, N, 4))), c(np.ones((T, N, 4))))
    W = [c(w) for w in orc.synth_mlp(2)]
    xn, un = rng.normal(size=(N, 3)), rng.normal(size=(T, N, 5))
    xs, ys, _ = ops.node_rk4_forward(*W, c(xn), c(un), 0.1)
    ops.node_rk4_vjp(*W, c(xn), c(un), xs, 0.1, torch.ones_like(xs), torch.ones_like(ys))
    os.environ["DYNAX_NODE_IMPL"] = "0"
    ops.node_rk4_forward(*W, c(xn), c(un), 0.1)
    os.environ.pop("DYNAX_NODE_IMPL")
    d = np.stack([15 + rng.random(T), rng.random(T), rng.random(T), rng.random(T)], 1)
    sch = [c(v) for v in (orc.RC_PRICE, orc.RC_T_LOW, orc.RC_T_HIGH)]
    ops.rc_mpc_cost(c(orc.RC_NOMINAL), c([20.0, 35.8, 26.0]), c(d), c(-rng.random((T, N))), *sch, 900.0)
    ops.tabular_interp(c(np.arange(T) * 60.0), c(d), c(rng.uniform(-10, T * 60 + 10, size=50)), "linear")
    ops.rc_env_step(c(orc.RC_NOMINAL), c(x0), c(np.zeros(N)), torch.zeros((3, N), dtype=torch.int32, device="cuda"),
                    c(np.arange(T) * 60.0), c(d), *sch, 60.0, num_actions=11)
    ops.resample_mean(c(rng.normal(size=(T, 6))), 4)
torch.cuda.synchronize()
print("sanitize_smoke: all kernels ran")
