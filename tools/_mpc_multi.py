"""Config 4 on N ranks: MPC cost + sharded argmin, time per step (torchrun)."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynax_b200 import ops
from dynax_b200.distributed import sharded_argmin
from tools import _synth as syn
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"])); dev = torch.device("cuda")
dist.init_process_group("nccl")
Nm, H = 1 << 22, 96
g = torch.Generator(device=dev).manual_seed(3456 + rank)
um = -10 * torch.rand((H, Nm), device=dev, generator=g)
d = torch.rand((H, 4), device=dev, generator=g); d[:, 0] = 15 + 10 * d[:, 0]
sch = [torch.tensor(x, dtype=torch.float32, device=dev) for x in (syn.RC_PRICE, syn.RC_T_LOW, syn.RC_T_HIGH)]
nom = torch.tensor(syn.RC_NOMINAL, dtype=torch.float32, device=dev); xm = torch.tensor([20.0, 35.8, 26.0], device=dev)
def mpc(sh):
    cost, amin, cmin = ops.rc_mpc_cost(nom, xm, d, um, *sch, 900.0)
    if sh: return sharded_argmin(cmin, amin, rank * Nm)
for sh in (False, True, False, True):
    for _ in range(5): mpc(sh)
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): r = mpc(sh)
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 50], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0: print(f"world {world} sharded_argmin={sh}: {float(t):.4f} ms/step", (float(r[0]), int(r[1])) if sh else "", flush=True)
dist.destroy_process_group()
