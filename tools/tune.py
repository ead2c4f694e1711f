"""Sweep the instantiated tile configurations on one GPU (CUDA-event timing).  Scratch tool for
choosing defaults; numbers for the record come from bench.py.

    python tools/tune.py [--T 4380] [--wave 7] [--what fwd,adj]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynax_b200 import ops  # noqa: E402

NOMINAL = [10384.31640625, 499089.09375, 1321535.125, 1.5348844528198242, 0.5000327825546265, 1.000040054321289,
           20.119935989379883]


def timeit(fn, iters=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e-3)
    return min(ts), sum(ts) / len(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--T", type=int, default=4380)
    ap.add_argument("--N", type=int, default=148 * 7 * 128)
    ap.add_argument("--what", default="fwd,tan,adj")
    ap.add_argument("--fwd-cfgs", default="0,3,4")
    ap.add_argument("--adj-cfgs", default="0,1")
    ap.add_argument("--fs-cfgs", default="0,2,3")
    a = ap.parse_args()
    N, T = a.N, a.T
    g = torch.Generator(device="cuda").manual_seed(0)
    if "mpc" in a.what:
        Nm, H = 1 << 22, 96
        um = -10 * torch.rand((H, Nm), device="cuda", generator=g)
        d = torch.rand((H, 4), device="cuda", generator=g)
        d[:, 0] = 15 + 10 * d[:, 0]
        price = torch.full((24,), 0.05, device="cuda"); lo = torch.full((24,), 22.0, device="cuda"); hi = torch.full((24,), 26.0, device="cuda")
        th1 = torch.tensor(NOMINAL, device="cuda"); x01 = torch.tensor([20.0, 35.8, 26.0], device="cuda")
        best, avg = timeit(lambda: ops.rc_mpc_cost(th1, x01, d, um, price, lo, hi, 900.0), iters=5)
        st = Nm * H
        print(f"mpc N={Nm} H={H} best {best*1e3:8.3f} ms avg {avg*1e3:8.3f} ms  {st/best:.3e} cand-steps/s  {st*4/best/1e9:7.1f} GB/s", flush=True)
        del um
        if a.what == "mpc":
            return
    th = torch.tensor(NOMINAL, device="cuda")[None] * torch.exp(torch.empty((N, 7), device="cuda").uniform_(-0.2, 0.2, generator=g))
    x0 = torch.tensor([[20.0, 30.0, 26.0]], device="cuda") + torch.empty((N, 3), device="cuda").uniform_(-1, 1, generator=g)
    u = torch.empty((T, N, 5), device="cuda").uniform_(0, 1, generator=g)
    u[..., 0] = 15 + 10 * u[..., 0]
    target = 20 + 2 * torch.empty((T, N), device="cuda").uniform_(-1, 1, generator=g)
    steps = N * T
    print(f"N={N} T={T} u={u.numel()*4/1e9:.2f} GB", flush=True)
    if "fwd" in a.what:
        for cfg in a.fwd_cfgs.split(","):
            os.environ["DYNAX_FWD_CFG"] = cfg
            for name, kw, B in (("y-only", dict(emit_x=False), 24), ("x+y", {}, 36)):
                best, avg = timeit(lambda: ops.rc_forward(th, x0, u, 3600.0, **kw))
                print(f"fwd cfg={cfg} {name:7s} best {best*1e3:8.2f} ms avg {avg*1e3:8.2f} ms  "
                      f"{steps/best:.3e} steps/s  {steps*B/best/1e9:7.1f} GB/s", flush=True)
    if "tan" in a.what:
        os.environ["DYNAX_GRAD_ALGO"] = "1"
        for cfg in a.fs_cfgs.split(","):
            os.environ["DYNAX_FS_CFG"] = cfg
            best, avg = timeit(lambda: ops.rc_loss_grad(th, x0, u, target, 3600.0), iters=4)
            print(f"tan cfg={cfg} best {best*1e3:8.2f} ms avg {avg*1e3:8.2f} ms  {steps/best:.3e} steps/s  "
                  f"{steps*24/best/1e9:7.1f} GB/s (24 B/step)", flush=True)
    if "adj" in a.what:
        os.environ["DYNAX_GRAD_ALGO"] = "0"
        for cfg in a.adj_cfgs.split(","):
            os.environ["DYNAX_ADJ_CFG"] = cfg
            best, avg = timeit(lambda: ops.rc_loss_grad(th, x0, u, target, 3600.0))
            print(f"adj cfg={cfg} best {best*1e3:8.2f} ms avg {avg*1e3:8.2f} ms  {steps/best:.3e} steps/s  "
                  f"{steps*48/best/1e9:7.1f} GB/s (48 B/step)", flush=True)


if __name__ == "__main__":
    main()
