"""Shared implementation of Discrete/ContinuousLinearSSM: four bias-free nn.Dense layers
(dynax/core/discrete_block_state_space.py:21-54, continuous_block_state_space.py:21-54).

Flax ``nn.Dense`` computes ``x @ kernel`` with ``kernel[in, out]``, so the state matrix is the
TRANSPOSE of the stored kernel; parameter paths are ``params/_fxx/dense/kernel`` [n,n],
``_fxu`` [m,n], ``_fyx`` [n,p], ``_fyu`` [m,p].  Kernels may carry a leading system axis [N,...].
"""
from __future__ import annotations

import math

import torch

from .. import ops


def lecun_normal(gen, fan_in, shape, device):
    """flax.linen.Dense default kernel_init: truncated normal, variance 1/fan_in."""
    std = math.sqrt(1.0 / fan_in) / 0.87962566103423978
    w = torch.empty(shape, dtype=torch.float32)
    torch.nn.init.trunc_normal_(w, mean=0.0, std=std, a=-2 * std, b=2 * std, generator=gen)
    return w.to(device)


class LinearSSMMixin:
    def init_params(self, seed, device):
        n, m, p = self.state_dim, self.input_dim, self.output_dim
        gen = torch.Generator().manual_seed(int(seed))
        mk = lambda i, o: {"dense": {"kernel": lecun_normal(gen, i, (i, o), device)}}
        return {"_fxx": mk(n, n), "_fxu": mk(m, n), "_fyx": mk(n, p), "_fyu": mk(m, p)}

    @staticmethod
    def matrices(params):
        """Flax kernels -> row-major matrices A[.,n,n], B[.,n,m], C[.,p,n], D[.,p,m]."""
        k = lambda name: params[name]["dense"]["kernel"].transpose(-1, -2).contiguous()
        return k("_fxx"), k("_fxu"), k("_fyx"), k("_fyu")

    def _rollout(self, params, x0, u, dt, discrete):
        A, B, C, D = self.matrices(params)
        return ops.lti_rollout(A, B, C, D, x0, u, dt, discrete)
