"""CPU, world_size 2, gloo: host-side logic of the sharded shared-parameter loss/gradient
(dynax_b200/distributed.py).  The per-rank partial sums come from the oracle here -- on the GPU
they come from dynax_rc_loss_grad_f32 with DYNAX_SHARED_THETA."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, N, T, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dynax_b200.distributed import shard_range, sharded_shared_loss_grad
    from oracle import dynax_oracle as orc
    th = orc.RC_NOMINAL.copy(); th[6] = 60.0          # above ub: penalty must be counted once
    x0, u = orc.synth_x0(N, 71), orc.synth_u(T, N, 71)
    target = (22 + np.random.default_rng(5).normal(size=(T, N))).astype(np.float32)
    lo, hi = shard_range(N, rank, world)

    def local_fn():
        l, g = orc.rc_loss_grad(np.repeat(th[None], hi - lo, 0), x0[lo:hi], u[:, lo:hi], target[:, lo:hi], 3600.0)
        return torch.tensor([l.sum()], dtype=torch.float64), torch.tensor(g.sum(0), dtype=torch.float64)

    loss, grad = sharded_shared_loss_grad(local_fn, th, orc.RC_LB, orc.RC_UB)
    ret[rank] = (float(loss), grad.numpy().copy(), (lo, hi))
    dist.destroy_process_group()


def test_shard_range_partitions():
    from dynax_b200.distributed import shard_range
    for N in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            r = [shard_range(N, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == N
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_box_penalty_matches_oracle():
    from dynax_b200.distributed import box_penalty
    from oracle import dynax_oracle as orc
    th = orc.RC_NOMINAL.copy(); th[0] = 1e3; th[6] = 60.0
    pen, g = box_penalty(th, orc.RC_LB, orc.RC_UB)
    assert np.isclose(float(pen), orc.bound_penalty(th[None], orc.RC_LB, orc.RC_UB)[0])
    assert np.allclose(g.numpy(), orc.bound_penalty_grad(th[None], orc.RC_LB, orc.RC_UB)[0])


@pytest.mark.timeout(120)
def test_two_rank_shared_theta_allreduce():
    from oracle import dynax_oracle as orc
    N, T, world = 11, 60, 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), N, T, ret), nprocs=world, join=True)
    th = orc.RC_NOMINAL.copy(); th[6] = 60.0
    x0, u = orc.synth_x0(N, 71), orc.synth_u(T, N, 71)
    target = (22 + np.random.default_rng(5).normal(size=(T, N))).astype(np.float32)
    l, g = orc.rc_loss_grad(np.repeat(th[None], N, 0), x0, u, target, 3600.0)
    want_l = l.sum() + orc.bound_penalty(th[None], orc.RC_LB, orc.RC_UB)[0]
    want_g = g.sum(0) + orc.bound_penalty_grad(th[None], orc.RC_LB, orc.RC_UB)[0]
    assert ret[0][2] == (0, 6) and ret[1][2] == (6, 11)
    for r in range(world):
        assert np.isclose(ret[r][0], want_l, rtol=1e-12)
        assert np.allclose(ret[r][1], want_g, rtol=1e-10)


def _argmin_worker(rank, world, port, N, H, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dynax_b200.distributed import shard_range, sharded_argmin
    cost = _mpc_costs(N, H)
    lo, hi = shard_range(N, rank, world)
    local = torch.tensor(cost[lo:hi])
    best, idx = sharded_argmin(local.min(), local.argmin(), lo)
    ret[rank] = (float(best), int(idx))
    dist.destroy_process_group()


def _mpc_costs(N, H):
    from oracle import dynax_oracle as orc
    rng = np.random.default_rng(9)
    d = np.stack([rng.uniform(10, 30, H), rng.uniform(0, 2, H), rng.uniform(0, 2, H), rng.uniform(0, 2, H)], 1).astype(np.float32)
    u = rng.uniform(-5, 0, (H, N)).astype(np.float32)
    cost = np.asarray(orc.mpc_cost(orc.RC_NOMINAL, np.array([20, 30, 26], np.float32), d, u, 900.0), np.float32)
    cost[1] = cost[N - 2] = cost.min() - 1  # the best cost twice, once per shard: the lower global index must win
    return cost


@pytest.mark.timeout(120)
def test_two_rank_mpc_argmin():
    """Config 4 on two ranks: every rank ends with the global best candidate (cost, index)."""
    N, H, world = 13, 24, 2
    cost = _mpc_costs(N, H)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_argmin_worker, args=(world, _free_port(), N, H, ret), nprocs=world, join=True)
    for r in range(world):
        assert ret[r] == (float(cost.min()), 1)


def test_sharded_argmin_tie_goes_to_lowest_index():
    from dynax_b200.distributed import sharded_argmin
    best, idx = sharded_argmin(torch.tensor(1.5), torch.tensor(3), 10)
    assert float(best) == 1.5 and int(idx) == 13
