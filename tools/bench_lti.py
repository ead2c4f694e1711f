"""Dense-LTI rollouts (row A4) and the general-cotangent RC VJP on one B200: CUDA-event timing vs the HBM roofline.

    python tools/bench_lti.py
"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynax_b200 import ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6650.0


def timeit(fn, iters=4, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) * 1e-3)
    return best


def emit(name, N, T, s, B, **kw):
    print(json.dumps(dict(workload=name, N=N, T=T, ms=s * 1e3, steps_per_s=N * T / s, alg_bytes_per_step=B,
                          achieved_gbs=N * T * B / s / 1e9, frac_of_measured_hbm_peak=N * T * B / s / 1e9 / PEAK, **kw)), flush=True)


def main():
    g = torch.Generator(device="cuda").manual_seed(0)
    for (n, m, p) in ((4, 3, 4), (3, 5, 1), (2, 1, 1), (4, 4, 4)):
        try:
            N, T = 148 * 3 * 128 * 2, 1024
            A = -0.05 * torch.eye(n, device="cuda") + 0.01 * torch.randn((n, n), device="cuda", generator=g)
            Bm = 0.1 * torch.randn((n, m), device="cuda", generator=g)
            C = torch.randn((p, n), device="cuda", generator=g); D = 0.1 * torch.randn((p, m), device="cuda", generator=g)
            x0 = torch.randn((N, n), device="cuda", generator=g)
            u = torch.randn((T, N, m), device="cuda", generator=g)
            s = timeit(lambda: ops.lti_forward(A, Bm, C, D, x0, u, 0.1))
            emit(f"lti forward x+y ({n},{m},{p}) shared matrices", N, T, s, 4 * (m + n + p))
            s = timeit(lambda: ops.lti_forward(A, Bm, C, D, x0, u, 0.1, emit_x=False))
            emit(f"lti forward y only ({n},{m},{p})", N, T, s, 4 * (m + p))
            gy = torch.randn((T, N, p), device="cuda", generator=g)
            s = timeit(lambda: ops.lti_vjp(A, Bm, C, D, x0, u, 0.1, None, gy, need_u=False))
            emit(f"lti vjp, g_ysol only, no g_u ({n},{m},{p})", N, T, s, 4 * (2 * m + p) + 24.0 * n / 8 / 1, note="u twice + g_y once + checkpoints")
            s = timeit(lambda: ops.lti_vjp(A, Bm, C, D, x0, u, 0.1, None, gy, need_u=True))
            emit(f"lti vjp, g_ysol only, with g_u ({n},{m},{p})", N, T, s, 4 * (3 * m + p))
            del u, gy
        except Exception as e:  # dims not instantiated
            print(json.dumps(dict(workload=f"lti ({n},{m},{p})", error=str(e)[:200])), flush=True)
    # RC general-cotangent VJP (what jax.grad of a user loss on simulator outputs needs)
    N, T = 148 * 3 * 128 * 2, 2048
    NOM = torch.tensor([10384.31640625, 499089.09375, 1321535.125, 1.5348844528198242, 0.5000327825546265, 1.000040054321289, 20.119935989379883], device="cuda")
    th = NOM[None] * torch.exp(torch.empty((N, 7), device="cuda").uniform_(-0.2, 0.2, generator=g))
    x0 = torch.tensor([[20.0, 30.0, 26.0]], device="cuda") + torch.empty((N, 3), device="cuda").uniform_(-1, 1, generator=g)
    u = torch.rand((T, N, 5), device="cuda", generator=g); u[..., 0] = 15 + 10 * u[..., 0]
    gy = torch.randn((T, N, 1), device="cuda", generator=g)
    s = timeit(lambda: ops.rc_vjp(th, x0, u, 3600.0, None, gy, need_u=False))
    emit("rc vjp, g_ysol only, no g_u", N, T, s, 4 * (2 * 5 + 1) + 3)
    s = timeit(lambda: ops.rc_vjp(th, x0, u, 3600.0, None, gy, need_u=True))
    emit("rc vjp, g_ysol only, with g_u", N, T, s, 4 * (3 * 5 + 1) + 3)


if __name__ == "__main__":
    main()
