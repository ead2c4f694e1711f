"""A/B timing of the forward kernel between two builds of the library, same process, same inputs.

    git stash; python -m dynax_b200.build; mkdir -p dynax_b200/_old; cp dynax_b200/libdynax_b200.so dynax_b200/_old/libdynax_old.so
    git stash pop; python -m dynax_b200.build
    gpurun -- 'python tools/_ab_fwd.py; DYNAX_B200_LIB=dynax_b200/_old/libdynax_old.so AB_NAME=old python tools/_ab_fwd.py'
(one build per process: two copies of the library in one process crashed on this image)

Prints ms for x+y and y-only per shape and the fp64 sums of both outputs (equal sums = same results)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynax_b200 import ops, _lib
NOM = [10384.31640625, 499089.09375, 1321535.125, 1.5348844528198242, 0.5000327825546265, 1.000040054321289, 20.119935989379883]
def t(fn, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(it):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
name = os.environ.get("AB_NAME", "new")
g = torch.Generator(device="cuda").manual_seed(0)
for N, T, tc in ((1, 8760, -1), (1, 8760, 0), (129, 8760, -1), (1023, 8760, 0), (75777, 2000, 0), (151553, 2000, 0), (151552, 2000, 0), (16384, 8760, 0)):
    th = torch.tensor(NOM, device="cuda")[None].repeat(N, 1)
    x0 = torch.tensor([[20.0, 30.0, 26.0]], device="cuda").repeat(N, 1)
    u = torch.rand((T, N, 5), device="cuda", generator=g)
    for rep in range(2):
        if True:
            a = t(lambda: ops.rc_forward(th, x0, u, 3600.0, time_chunks=tc))
            b = t(lambda: ops.rc_forward(th, x0, u, 3600.0, emit_x=False, time_chunks=tc))
            xs, ys, _ = ops.rc_forward(th, x0, u, 3600.0, time_chunks=tc)
            print(f"N={N:6d} T={T} chunks={tc:2d} {name}: x+y {a:8.4f} ms | y only {b:8.4f} ms  chk {float(xs.double().sum()):.10e} {float(ys.double().sum()):.10e}", flush=True)
    del u
