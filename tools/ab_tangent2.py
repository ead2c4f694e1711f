"""A/B timing of the fused loss+gradient between the round-1 library (dynax_b200/_old/libdynax_old.so, if present) and
the current one, incl. the threads-per-system variants of the current tangent kernel (DYNAX_FS_SPLIT) on launches that
cannot fill the GPU.  Prints best single launch, mean of `reps` launches back to back (power-capped steady state)
and fp64 checksums."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynax_b200 import ops, _lib
from tools._synth import RC_NOMINAL
NEW = _lib.LIB_PATH
OLD = os.path.join(os.path.dirname(NEW), "_old", "libdynax_old.so")
g = torch.Generator(device="cuda").manual_seed(0)
T = 8760


def use(path):
    _lib._LIB = None; _lib.LIB_PATH = path; _lib.lib()


def timeit(fn, reps):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): out = fn()
    e1.record(); torch.cuda.synchronize()
    return best, e0.elapsed_time(e1) / reps, out


def data(N, shared_series=False):
    th = torch.tensor(RC_NOMINAL, device="cuda", dtype=torch.float32)[None] * torch.exp(torch.empty((N, 7), device="cuda").uniform_(-0.2, 0.2, generator=g))
    x0 = torch.tensor([[20.0, 30.0, 26.0]], device="cuda").repeat(N, 1)
    Nu = 1 if shared_series else N
    u = torch.rand((T, Nu, 5), device="cuda", generator=g); u[..., 0] = 15 + 10 * u[..., 0]
    tg = 20 + 2 * torch.rand((T, Nu), device="cuda", generator=g)
    return th, x0, u, (tg[:, 0].contiguous() if shared_series else tg)


sizes = [int(s) for s in (sys.argv[1].split(",") if len(sys.argv) > 1 else "227328,65536,16384,8192,4096".split(","))]
for N in sizes:
    th, x0, u, tg = data(N)
    for shared_theta in (False, True):
        thx = th[0].contiguous() if shared_theta else th
        fn = lambda: ops.rc_loss_grad(thx, x0, u, tg, 3600.0, time_chunks=0)
        variants = [("old", OLD, None)] if os.path.exists(OLD) else []
        variants += [("new", NEW, None)]
        for extra in os.environ.get("AB_LIBS", "").split(","):
            pth = os.path.join(os.path.dirname(NEW), "_old", f"libdynax_{extra}.so")
            if extra and os.path.exists(pth):
                variants.append((extra, pth, None))
        if N <= 40000 and not os.environ.get("AB_NO_SPLIT"):
            variants += [("new split1", NEW, "1"), ("new split2", NEW, "2"), ("new split4", NEW, "4")]
        for rep in range(2 if N > 100000 else 1):
            for name, path, split in variants:
                use(path)
                if split is None: os.environ.pop("DYNAX_FS_SPLIT", None)
                else: os.environ["DYNAX_FS_SPLIT"] = split
                best, mean, out = timeit(fn, 30 if N > 100000 else 20)
                print(f"N={N:6d} shared_theta={int(shared_theta)} {name:11s}: best {best:7.3f} ms  sustained {mean:7.3f} ms  "
                      f"chk {float(out[0].double().sum()):.10e} {float(out[1].double().abs().sum()):.10e}", flush=True)
        if N <= 40000 and not shared_theta and not os.environ.get("AB_NO_SPLIT"):   # the time-parallel path for comparison
            use(NEW); os.environ.pop("DYNAX_FS_SPLIT", None)
            best, mean, out = timeit(lambda: ops.rc_loss_grad(thx, x0, u, tg, 3600.0, time_chunks=-1), 20)
            print(f"N={N:6d} chunked(auto)          : best {best:7.3f} ms  sustained {mean:7.3f} ms", flush=True)
    del u, tg
# config 3 as launched: candidates against ONE measured series
for N in ((65536, 8192) if not os.environ.get("AB_NO_SPLIT") else ()):
    th, x0, u, tg = data(N, shared_series=True)
    for name, path in ([("old", OLD)] if os.path.exists(OLD) else []) + [("new", NEW)]:
        use(path); os.environ.pop("DYNAX_FS_SPLIT", None)
        best, mean, out = timeit(lambda: ops.rc_loss_grad(th, x0, u, tg, 3600.0, time_chunks=0), 20)
        print(f"N={N:6d} shared u+target {name}: best {best:7.3f} ms  sustained {mean:7.3f} ms  chk {float(out[0].double().sum()):.10e}", flush=True)
