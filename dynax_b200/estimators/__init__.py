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


def calibrate(simulator, variables, x0, u, target, params_lb=None, params_ub=None, n_epochs=5000, lr=1e-3,
              log_every=0, cuda_graph=False):
    """The training loop of examples/3-inverse.py:117-138 kept entirely on the device (row f-3):
    every epoch = one fused loss+gradient launch + one LAMB update launch, no host round trip.

    ``variables`` holds [N] leaves (N candidates calibrated independently) or scalar leaves (one
    candidate).  Returns (variables_fitted, loss[N] of the last epoch).
    ``cuda_graph=True`` captures one epoch (all kernels of the loss/gradient + the update, epoch counter on
    the device) into a CUDA graph and replays it: for small batches the epoch is launch-bound and the graph
    removes the per-launch host cost.
    """
    model = simulator.model
    if not isinstance(model, _RC4R3C):
        raise NotImplementedError("calibrate is implemented for the 4R3C model")
    x0 = torch.as_tensor(x0, dtype=torch.float32)
    x0 = x0.cuda() if not x0.is_cuda else x0
    dev = x0.device
    u = torch.as_tensor(u, dtype=torch.float32).to(dev)
    target = torch.as_tensor(target, dtype=torch.float32).to(dev)
    leaves = variables["params"]["model"]
    theta = stack_theta(leaves, dev).detach().clone().contiguous()
    if x0.dim() == 1:
        x0 = x0.reshape(1, 3)
        u = u[:, None, :] if u.dim() == 2 else u
        target = target.reshape(u.shape[0], -1)
    if theta.shape[0] == 1 and x0.shape[0] > 1:
        raise NotImplementedError("calibrate updates per-candidate parameters; pass [N] leaves")
    lb, ub = _bounds(params_lb, dev), _bounds(params_ub, dev)
    m, v = torch.zeros_like(theta), torch.zeros_like(theta)
    loss = None
    if cuda_graph and int(n_epochs) > 3:
        counter = torch.zeros((), dtype=torch.int64, device=dev)

        def epoch_fn():
            l, g, _ = ops.rc_loss_grad(theta, x0, u, target, float(simulator.dt), lb, ub)
            ops.lamb_update_dev_(theta, g.contiguous(), m, v, counter, lr)
            return l

        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):          # warm-up outside capture (lazy module loads, attribute calls)
            for _ in range(2):
                loss = epoch_fn()
        torch.cuda.current_stream(dev).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            loss = epoch_fn()
        for _ in range(int(n_epochs) - 2):      # capture records without executing: 2 warm-up epochs ran so far
            graph.replay()
        fitted = {"params": {"model": unstack_theta(theta, leaves)}}
        return fitted, loss
    for epoch in range(1, int(n_epochs) + 1):
        loss, grad, _ = ops.rc_loss_grad(theta, x0, u, target, float(simulator.dt), lb, ub)
        ops.lamb_update_(theta, grad.contiguous(), m, v, epoch, lr)
        if log_every and epoch % log_every == 0:
            print("loss at epoch %d: %.4f, grad norm %.4f" % (epoch, float(loss.mean()), float(grad.norm())))
    fitted = {"params": {"model": unstack_theta(theta, leaves)}}
    return fitted, loss
