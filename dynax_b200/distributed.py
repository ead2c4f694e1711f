"""Multi-GPU: the system axis is sharded across ranks (one process per GPU); rollouts need no
communication.  Only gradients of parameters SHARED by all systems are all-reduced (NCCL over
NVLink): 7 + 1 floats per loss/grad call for the RC model.

The host logic is independent of the device kernels (``local_fn`` computes the per-rank partial
sums), so it is covered on CPU with world_size-2 gloo tests (tests/test_dist_gloo.py).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

__all__ = ["shard_range", "allreduce_sum_", "sharded_shared_loss_grad", "sharded_argmin", "box_penalty"]


def shard_range(N, rank, world):
    """Contiguous block [lo, hi) of the system axis owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(int(N), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_sum_(t, group=None):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def box_penalty(theta, lb, ub):
    """sum_j relu(theta_j-ub_j)/ub_j + relu(lb_j-theta_j)/ub_j and its gradient
    (examples/3-inverse.py:103-107; normaliser is ub for both terms)."""
    th, lb, ub = (torch.as_tensor(v, dtype=torch.float64).reshape(-1).cpu() for v in (theta, lb, ub))
    over, under = th > ub, th < lb
    pen = (torch.where(over, (th - ub) / ub, torch.zeros_like(th)) + torch.where(under, (lb - th) / ub, torch.zeros_like(th))).sum()
    g = over.double() / ub - under.double() / ub
    return pen, g


def sharded_shared_loss_grad(local_fn, theta, lb=None, ub=None, group=None, device=None):
    """Loss and gradient of ONE shared parameter set over a system batch sharded across ranks.

    ``local_fn()`` -> (loss_sum, grad_sum[P]) over this rank's systems, WITHOUT the box penalty
    (e.g. ``ops.rc_loss_grad(theta, x0_local, u_local, target_local, dt)[:2]``).  The 1+P partial
    sums are packed into one fp32 buffer and summed with a single all-reduce; the penalty is added
    once, identically on every rank.  Returns (loss, grad[P]) as tensors on the partials' device.
    """
    loss, grad = local_fn()
    grad = grad.reshape(-1)
    buf = torch.cat([loss.reshape(1).to(grad.dtype), grad])
    if device is not None:
        buf = buf.to(device)
    allreduce_sum_(buf, group)
    loss, grad = buf[0], buf[1:]
    if lb is not None:
        pen, g = box_penalty(theta, lb, ub)
        loss = loss + pen.to(buf.dtype).to(buf.device)
        grad = grad + g.to(buf.dtype).to(buf.device)
    return loss, grad


def sharded_argmin(local_min, local_argmin, offset, group=None):
    """Best candidate over a candidate axis sharded across ranks (config 4).

    ``local_min`` / ``local_argmin``: what ``ops.rc_mpc_cost`` returns for this rank's candidates
    (0-d tensors); ``offset``: global index of the rank's first candidate (``shard_range(...)[0]``).
    One all-gather of two 8-byte values per rank; ties go to the lowest global index, like ``argmin``
    over the unsharded axis.  Returns (min_cost, global_index) as 0-d tensors, identical on every rank.
    """
    # few launches: this runs inside the control loop next to a 0.3 ms kernel (bench.py workloads.config4_mpc)
    pair = torch.empty((2,), dtype=torch.float64, device=local_min.device)   # indices < 2^53 are exact
    pair[0].copy_(local_min.reshape(()))
    pair[1].copy_(local_argmin.reshape(()))
    pair[1] += int(offset)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        allp = torch.empty((dist.get_world_size(group), 2), dtype=torch.float64, device=pair.device)
        if pair.is_cuda:   # NCCL: one collective into the [world, 2] buffer
            dist.all_gather_into_tensor(allp, pair, group=group)
        else:              # gloo (CPU tests) has no all_gather_into_tensor
            dist.all_gather(list(allp.unbind(0)), pair, group=group)
    else:
        allp = pair[None]
    # argmin returns the FIRST minimal entry = the lowest rank = the lowest global index among equal costs
    # (within a rank the kernel has already kept the lowest index)
    # (index_select: indexing with a 0-d tensor would synchronise the stream to read it)
    win = allp.index_select(0, torch.argmin(allp[:, 0]).reshape(1)).reshape(2)
    return win[0].to(local_min.dtype), win[1].to(torch.int64)
