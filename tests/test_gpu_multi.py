"""GPU, world_size 2 (skipped on a 1-GPU box): sharding the system axis over ranks gives per-system
results identical to one GPU, and the NCCL all-reduce of shared-parameter gradients equals the
1-GPU sum up to fp32 reassociation."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, N, T, ret):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from dynax_b200 import ops
    from dynax_b200.distributed import shard_range, sharded_argmin, sharded_shared_loss_grad
    from oracle import dynax_oracle as orc
    th, x0, u = orc.synth_theta(N, 81), orc.synth_x0(N, 81), orc.synth_u(T, N, 81)
    target = (22 + np.random.default_rng(6).normal(size=(T, N))).astype(np.float32)
    lo, hi = shard_range(N, rank, world)
    c = lambda a: torch.tensor(np.ascontiguousarray(a, dtype=np.float32)).cuda()
    d = (c(x0[lo:hi]), c(u[:, lo:hi]), c(target[:, lo:hi]))
    l_i, g_i, _ = ops.rc_loss_grad(c(th[lo:hi]), *d, 3600.0)                    # per-system theta: no comm
    th1 = c(orc.RC_NOMINAL.astype(np.float32))
    loss, grad = sharded_shared_loss_grad(lambda: ops.rc_loss_grad(th1, *d, 3600.0)[:2], th1, orc.RC_LB, orc.RC_UB)
    # config 4: candidates sharded over ranks, all-reduce-min of (cost, index)
    dm, um = _mpc_inputs(N)
    cost, amin, cmin = ops.rc_mpc_cost(th1, c(np.array([20, 30, 26], np.float32)), c(dm), c(um[:, lo:hi]),
                                       c(orc.RC_PRICE), c(orc.RC_T_LOW), c(orc.RC_T_HIGH), 900.0)
    best, idx = sharded_argmin(cmin, amin, lo)
    ret[rank] = (lo, hi, l_i.cpu().numpy(), g_i.cpu().numpy(), float(loss), grad.cpu().numpy(), float(best), int(idx))
    dist.destroy_process_group()


def _mpc_inputs(N, H=48):
    rng = np.random.default_rng(12)
    d = np.stack([rng.uniform(10, 30, H), rng.uniform(0, 2, H), rng.uniform(0, 2, H), rng.uniform(0, 2, H)], 1).astype(np.float32)
    return d, rng.uniform(-5, 0, (H, N)).astype(np.float32)


def test_two_gpu_sharding_matches_one_gpu():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from dynax_b200 import ops
    from oracle import dynax_oracle as orc
    N, T, world = 1000, 200, 2
    mgr = mp.Manager(); ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), N, T, ret), nprocs=world, join=True)
    th, x0, u = orc.synth_theta(N, 81), orc.synth_x0(N, 81), orc.synth_u(T, N, 81)
    target = (22 + np.random.default_rng(6).normal(size=(T, N))).astype(np.float32)
    c = lambda a: torch.tensor(np.ascontiguousarray(a, dtype=np.float32)).cuda()
    l1, g1, _ = ops.rc_loss_grad(c(th), c(x0), c(u), c(target), 3600.0)
    l_cat = np.concatenate([ret[r][2] for r in range(world)]); g_cat = np.concatenate([ret[r][3] for r in range(world)])
    assert np.array_equal(l_cat, l1.cpu().numpy()) and np.array_equal(g_cat, g1.cpu().numpy())
    ls, gs, _ = ops.rc_loss_grad(c(orc.RC_NOMINAL.astype(np.float32)), c(x0), c(u), c(target), 3600.0, orc.RC_LB, orc.RC_UB)
    for r in range(world):
        assert np.isclose(ret[r][4], ls.item(), rtol=1e-5)
        assert np.allclose(ret[r][5], gs.cpu().numpy(), rtol=1e-5, atol=1e-12)
    dm, um = _mpc_inputs(N)
    _, amin, cmin = ops.rc_mpc_cost(c(orc.RC_NOMINAL.astype(np.float32)), c(np.array([20, 30, 26], np.float32)), c(dm), c(um),
                                    c(orc.RC_PRICE), c(orc.RC_T_LOW), c(orc.RC_T_HIGH), 900.0)
    for r in range(world):
        assert ret[r][6:] == (float(cmin), int(amin))
