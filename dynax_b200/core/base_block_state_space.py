"""Host-side mirror of dynax/core/base_block_state_space.py.

The reference's BaseBlockSSM is a Flax module whose ``__call__(x, u)`` returns ``(rhs, y)`` with
``rhs = fxx(x)+fxu(u)`` / ``fx(x,u)`` and ``y = fyx(x)+fyu(u)`` / ``fy(x,u)`` and raises
NotImplementedError when neither form is defined (base_block_state_space.py:51-76).  Here a model is
a small spec object: it names which CUDA kernel family implements that split form and how its Flax
parameter tree maps onto the kernel's arguments.  There is no Python evaluation of the model.
"""
from __future__ import annotations

import torch


def _seed_of(key):
    """Accept an int seed, a torch.Generator, a (jax-style) uint32 key array, or None."""
    if key is None:
        return 0
    if isinstance(key, torch.Generator):
        return key.initial_seed()
    if isinstance(key, int):
        return key
    try:
        flat = torch.as_tensor(key).reshape(-1).tolist()
        return int(flat[-1]) & 0x7FFFFFFF
    except Exception:
        return 0


class BaseBlockSSM:
    """state_dim / input_dim / output_dim are the dataclass fields of the reference (:39-41)."""

    state_dim: int = None
    input_dim: int = None
    output_dim: int = None

    def __init__(self, state_dim=None, input_dim=None, output_dim=None, name=None):
        if state_dim is not None:
            self.state_dim = state_dim
        if input_dim is not None:
            self.input_dim = input_dim
        if output_dim is not None:
            self.output_dim = output_dim
        self.name = name

    # -- to be provided by concrete models ------------------------------------------------------
    def init_params(self, seed, device):
        raise NotImplementedError("dynamic equation is not implemented")      # :65

    def _rollout(self, params, x0, u, dt, discrete):
        """Differentiable batched rollout (xsol[T,N,n], ysol[T,N,p]) on the CUDA library."""
        raise NotImplementedError("dynamic equation is not implemented")      # :65

    # -- Flax-shaped API ------------------------------------------------------------------------
    def init(self, key, x, u):
        """``model.init(key, x, u)`` -> ``{'params': {...}}`` (tests/test_forward_simulation.py:26,60)."""
        dev = x.device if isinstance(x, torch.Tensor) and x.is_cuda else torch.device("cuda")
        return {"params": self.init_params(_seed_of(key), dev)}

    def apply(self, variables, x, u):
        """``model.apply(params, x, u)`` -> ``(rhs, y)``: ONE evaluation of the block SSM (:51-57).

        Runs as a T=1 launch with the discrete update (x_next = rhs), so the value returned is rhs.
        """
        x, u, batched = _as_batch_step(x, u, self.state_dim, self.input_dim)
        xsol, ysol = self._rollout(variables["params"], x, u[None], 1.0, True)
        rhs, y = xsol[0], ysol[0]
        return (rhs, y) if batched else (rhs[0], y[0])



class BaseDiscreteBlockSSM(BaseBlockSSM):
    """Pass-through subclass, as in the reference (:79-84)."""


class BaseContinuousBlockSSM(BaseBlockSSM):
    """Pass-through subclass, as in the reference (:86-90)."""


def _as_batch_step(x, u, n, m):
    x = torch.as_tensor(x, dtype=torch.float32)
    u = torch.as_tensor(u, dtype=torch.float32)
    if not x.is_cuda:
        x = x.cuda()
    if not u.is_cuda:
        u = u.to(x.device)
    batched = x.dim() == 2
    if not batched:
        x, u = x.reshape(1, n), u.reshape(1, m)
    return x, u, batched
