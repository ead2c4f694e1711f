"""A/B timing of the fused loss+gradient between builds of the library, ONE PROCESS PER BUILD (DYNAX_B200_LIB),
alternating:  python tools/ab_tangent3.py new,old,diaglast [N,N,...]   ('new' = the in-tree library, others =
dynax_b200/_old/libdynax_<name>.so).  Child mode (AB_CHILD=1) prints best / sustained (30 launches back to back)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if os.environ.get("AB_CHILD"):
    import torch
    from dynax_b200 import ops
    from tools._synth import RC_NOMINAL
    g = torch.Generator(device="cuda").manual_seed(0)
    T = 8760
    for N in [int(s) for s in sys.argv[1].split(",")]:
        th = torch.tensor(RC_NOMINAL, device="cuda", dtype=torch.float32)[None] * torch.exp(torch.empty((N, 7), device="cuda").uniform_(-0.2, 0.2, generator=g))
        x0 = torch.tensor([[20.0, 30.0, 26.0]], device="cuda").repeat(N, 1)
        u = torch.rand((T, N, 5), device="cuda", generator=g); u[..., 0] = 15 + 10 * u[..., 0]
        tg = 20 + 2 * torch.rand((T, N), device="cuda", generator=g)
        fn = lambda: ops.rc_loss_grad(th, x0, u, tg, 3600.0, time_chunks=0)
        for _ in range(3): fn()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        reps = 60 if N > 100000 else 30
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): out = fn()
        e1.record(); torch.cuda.synchronize()
        print(f"N={N:7d} {os.environ.get('AB_NAME'):34s}: best {best:7.3f} ms  sustained {e0.elapsed_time(e1)/reps:7.3f} ms = {N*T/(e0.elapsed_time(e1)/reps*1e-3):.4e} steps/s  "
              f"chk {float(out[0].double().sum()):.10e} {float(out[1].double().abs().sum()):.10e}", flush=True)
        del u, tg
    sys.exit(0)
names = sys.argv[1].split(",")
sizes = sys.argv[2] if len(sys.argv) > 2 else "227328,65536"
for rep in range(2):
    for nm in names:
        nm, _, own = nm.partition("#")        # name#N,N: this variant's own sizes
        env = dict(os.environ, AB_CHILD="1", AB_NAME=nm)
        lib, *sets = nm.split("@")          # e.g. new@DYNAX_FS_SPLIT=2@DYNAX_NO_TENSOR=1
        for kv in sets:
            k, v = kv.split("=")
            env[k] = v
        if lib != "new":
            env["DYNAX_B200_LIB"] = os.path.join(ROOT, "dynax_b200", "_old", f"libdynax_{lib}.so")
        subprocess.run([sys.executable, os.path.abspath(__file__), own or sizes], env=env)
