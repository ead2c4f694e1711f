"""Mirror of dynax/models/RC.py: the 4R3C thermal zone model.

States [Tai, Twe, Twi], inputs [Tao, qCon_i, qHVAC, qRad_e, qRad_i], output Tai (RC.py:8-19).
Parameters are the seven Flax scalars Cai,Cwe,Cwi,Re,Ri,Rw,Rg initialised to ones (RC.py:176-182);
A(theta), B(theta) are built inside the CUDA kernel (csrc/models.cuh, following RC.py:201-210,222-230).
Each parameter leaf is a 0-d tensor (one model shared by every system) or an [N] tensor (one value
per system -- what a caller of the reference would express with jax.vmap over the params).
"""
from __future__ import annotations

import torch

from .. import ops
from ..core.base_block_state_space import BaseContinuousBlockSSM, BaseDiscreteBlockSSM

PARAM_NAMES = ("Cai", "Cwe", "Cwi", "Re", "Ri", "Rw", "Rg")


def stack_theta(params, device=None):
    """Flax-style dict of 7 leaves -> theta[N,7] (or [1,7] when every leaf is a scalar)."""
    if not any(isinstance(params[k], torch.Tensor) for k in PARAM_NAMES):
        # plain Python / NumPy scalars (how the reference's examples pass a fitted model): ONE host-to-device copy
        return torch.tensor([[float(params[k]) for k in PARAM_NAMES]], dtype=torch.float32).to(device or "cuda")
    leaves = []
    for k in PARAM_NAMES:
        v = params[k]
        if not isinstance(v, torch.Tensor):
            v = torch.tensor(float(v), dtype=torch.float32, device=device or "cuda")
        leaves.append(v.to(torch.float32).reshape(-1))
    L = max(v.numel() for v in leaves)
    leaves = [v if v.numel() == L else v.expand(L) for v in leaves]
    return torch.stack(leaves, dim=-1)


def unstack_theta(theta, like):
    """theta-shaped tensor [L,7] -> dict with leaves shaped like ``like``'s leaves."""
    out = {}
    for j, k in enumerate(PARAM_NAMES):
        ref = like[k]
        col = theta[..., j]
        if isinstance(ref, torch.Tensor):
            out[k] = col.sum().reshape(ref.shape) if ref.numel() == 1 and col.numel() > 1 else col.reshape(ref.shape)
        else:
            out[k] = col.sum() if col.numel() > 1 else col.reshape(())
    return out


class _RC4R3C:
    state_dim = 3
    input_dim = 5
    output_dim = 1

    def init_params(self, seed, device):
        return {k: torch.ones((), dtype=torch.float32, device=device) for k in PARAM_NAMES}   # RC.py:176-182

    def _rollout(self, params, x0, u, dt, discrete):
        theta = stack_theta(params, x0.device)
        return ops.rc_rollout(theta, x0, u, dt, discrete)


class Continuous4R3C(_RC4R3C, BaseContinuousBlockSSM):
    """dynax/models/RC.py:137-249."""


class Discrete4R3C(_RC4R3C, BaseDiscreteBlockSSM):
    """dynax/models/RC.py:21-134 (identical matrices; stepped as x_next = rhs when applied directly)."""
