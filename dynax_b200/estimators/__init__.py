"""Inverse-inference helpers: the loss/gradient step of examples/3-inverse.py:93-115.

The reference's ``dynax/estimators`` package is empty; "inverse inference" lives in the example
script as ``jax.value_and_grad(mse_loss)``.  ``mse_loss_and_grad`` is that call for a whole batch of
parameter candidates, served by ONE fused kernel (forward sweep + checkpointed adjoint + loss +
box penalty), without materialising ysol or any cotangent array.
"""
from __future__ import annotations

import torch

from .. import ops
from ..models.RC import PARAM_NAMES, _RC4R3C, stack_theta, unstack_theta


def _bounds(tree, device):
    if tree is None:
        return None
    leaf = tree["params"]["model"] if "params" in tree else tree
    return torch.tensor([float(leaf[k]) for k in PARAM_NAMES], dtype=torch.float32, device=device)


def mse_loss_and_grad(simulator, variables, x0, u, target, params_lb=None, params_ub=None):
    """(loss, grads) = value_and_grad(mse_loss)(params)   [examples/3-inverse.py:96-112]

    variables: ``{'params': {'model': {Cai..Rg}}}`` with scalar leaves (one shared model: loss is the
    SUM over the batch of per-system MSEs + one penalty) or [N] leaves (N candidates: loss[N]).
    Returns ``loss`` and a gradient tree shaped like ``variables``.
    """
    model = simulator.model
    if not isinstance(model, _RC4R3C):
        raise NotImplementedError("fused loss/grad is implemented for the 4R3C model; "
                                  "use simulator.apply + torch autograd for other block SSMs")
    x0 = torch.as_tensor(x0, dtype=torch.float32)
    if not x0.is_cuda:
        x0 = x0.cuda()
    dev = x0.device
    u = torch.as_tensor(u, dtype=torch.float32).to(dev)
    target = torch.as_tensor(target, dtype=torch.float32).to(dev)
    leaves = variables["params"]["model"]
    theta = stack_theta(leaves, dev).detach()
    if x0.dim() == 1:
        x0 = x0.reshape(1, 3)
        if u.dim() == 2:
            u = u[:, None, :]
        target = target.reshape(u.shape[0], -1)
    loss, grad, _ = ops.rc_loss_grad(theta, x0, u, target, float(simulator.dt), _bounds(params_lb, dev),
                                     _bounds(params_ub, dev))
    grads = {"params": {"model": unstack_theta(grad.reshape(-1, 7), leaves)}}
    return loss, grads
