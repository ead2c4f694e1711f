"""Secondary workloads of BASELINE.json's `configs` on one B200 (the headline fwd+grad line is bench.py).
Writes one JSON line per workload to stdout; CUDA-event timing, best of `--iters` after warm-up.

    python tools/bench_all.py > profiles/r01_all_workloads.jsonl
"""
import argparse, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynax_b200 import ops
from tools import _synth as orc

PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6650.0


def timeit(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) * 1e-3)
    return best


def emit(name, units, sec, bytes_per_unit=None, **kw):
    line = {"workload": name, "units_per_call": units, "ms": sec * 1e3, "units_per_s": units / sec}
    if bytes_per_unit:
        line.update(alg_bytes_per_unit=bytes_per_unit, achieved_gbs=units * bytes_per_unit / sec / 1e9,
                    frac_of_measured_hbm_peak=units * bytes_per_unit / sec / 1e9 / PEAK)
    line.update(kw)
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser(); ap.add_argument("--iters", type=int, default=5); a = ap.parse_args()
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    nom = torch.tensor(orc.RC_NOMINAL, dtype=torch.float32, device=dev)
    # --- config 2: forward, T=8760, one full-machine tile (inputs 23 GB >> L2)
    N, T = 148 * 2 * 256 * 2, 8760
    th = nom[None] * torch.exp(torch.empty((N, 7), device=dev).uniform_(-0.2, 0.2, generator=g))
    x0 = torch.tensor([[20.0, 30.0, 26.0]], device=dev) + torch.empty((N, 3), device=dev).uniform_(-1, 1, generator=g)
    u = torch.empty((T, N, 5), device=dev).uniform_(0, 1, generator=g); u[..., 0] = 15 + 10 * u[..., 0]
    s = timeit(lambda: ops.rc_forward(th, x0, u, 3600.0), a.iters)
    emit("config2 forward x+y (reference semantics)", N * T, s, 36, N=N, T=T, kernel="rollout_fwd_kernel<RCModel,256,4,4>")
    s = timeit(lambda: ops.rc_forward(th, x0, u, 3600.0, emit_x=False), a.iters)
    emit("config2 forward y only", N * T, s, 24, N=N, T=T, kernel="rollout_fwd_kernel<RCModel,256,4,4>")
    del u
    # --- config 3: inverse, N=65536 candidates, per-system u/target (HBM-bound variant i) and shared (variant ii)
    N3 = 65536
    u3 = torch.empty((T, N3, 5), device=dev).uniform_(0, 1, generator=g); u3[..., 0] = 15 + 10 * u3[..., 0]
    tg3 = 20 + 2 * torch.empty((T, N3), device=dev).uniform_(0, 1, generator=g)
    us, ts_ = u3[:, 0].contiguous(), tg3[:, 0].contiguous()
    for algo, env, B, kern in (("tangent", "1", 24, "rc_fwdsens_kernel<224,8,2,2,packed>"),
                               ("adjoint", "0", 48, "rollout_adj_kernel<RCModel,128,12|16,2,MSE>")):
        os.environ["DYNAX_GRAD_ALGO"] = env
        s = timeit(lambda: ops.rc_loss_grad(th[:N3], x0[:N3], u3, tg3, 3600.0), a.iters)
        emit(f"config3 inverse loss+grad, per-system u/target [{algo}]", N3 * T, s, B, N=N3, T=T, kernel=kern)
        s = timeit(lambda: ops.rc_loss_grad(th[:N3], x0[:N3], us, ts_, 3600.0), a.iters)
        emit(f"config3 inverse loss+grad, shared u/target (compute-bound) [{algo}]", N3 * T, s, None, N=N3, T=T)
    os.environ["DYNAX_GRAD_ALGO"] = "1"
    s = timeit(lambda: ops.rc_loss_grad(nom, x0[:N3], u3, tg3, 3600.0), a.iters)
    emit("config3 shared-theta loss+grad (tangent kernel, fp64 warp sums + atomics)", N3 * T, s, 24, N=N3, T=T)
    del u3, tg3
    # --- config 4: MPC
    Nm, H = 1 << 22, 96
    um = -10 * torch.rand((H, Nm), device=dev, generator=g)
    d = torch.rand((H, 4), device=dev, generator=g); d[:, 0] = 15 + 10 * d[:, 0]
    sch = [torch.tensor(v, dtype=torch.float32, device=dev) for v in (orc.RC_PRICE, orc.RC_T_LOW, orc.RC_T_HIGH)]
    xm = torch.tensor([20.0, 35.8, 26.0], device=dev)
    s = timeit(lambda: ops.rc_mpc_cost(nom, xm, d, um, *sch, 900.0), a.iters)
    emit("config4 MPC cost + argmin", Nm * H, s, 4, N=Nm, H=H, kernel="rc_mpc_cost_kernel", bound="fp32 issue slots")
    del um
    # --- config 5: neural ODE
    Nn, Tn = 262144, 2048
    W = [torch.tensor(w).to(dev) for w in orc.synth_mlp(7)]
    xn = torch.randn((Nn, 3), device=dev, generator=g); un = torch.randn((Tn, Nn, 5), device=dev, generator=g)
    s = timeit(lambda: ops.node_rk4_forward(*W, xn, un, 0.25), 2, warm=1)
    emit("config5 neural-ODE RK4 forward (tcgen05, 3xTF32)", Nn * Tn, s, None, N=Nn, T=Tn, mlp_tflops=Nn * Tn * 38400 / s / 1e12,
         tensor_tflops_issued=3 * Nn * Tn * 4 * 2 * (8 * 64 + 64 * 64) / s / 1e12, kernel="node_tc_kernel")
    os.environ["DYNAX_NODE_IMPL"] = "0"
    s0 = timeit(lambda: ops.node_rk4_forward(*W, xn[:32768], un[:256, :32768].contiguous(), 0.25), 1, warm=1)
    emit("config5 neural-ODE RK4 forward, CUDA-core kernel (reference point)", 32768 * 256, s0, None, N=32768, T=256)


if __name__ == "__main__":
    main()
