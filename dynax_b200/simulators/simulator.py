"""Mirror of dynax/simulators/simulator.py:12-55 -- ``DifferentiableSimulator``.

Same constructor fields (model, dt, mode_interp, start_time), same ``init`` / ``apply`` call shape
and parameter tree (``{'params': {'model': {...}}}``, examples/1-forward.py:44-47), same output
convention (``xsol[k] = x_{k+1}``, ``ysol[k] = y(x_k,u_k)``, simulator.py:38-43).  The nn.scan is one
kernel launch; ``apply`` is differentiable w.r.t. params, x0 and inputs through torch autograd
(one adjoint-kernel launch), which is what jax.grad gives the reference.

Batching: the reference is single-system and batches through jax.vmap; here ``x0[N,n]`` with
``inputs[T,N,m]`` (north_star layout) runs N systems in one launch, ``x0[n]`` with ``inputs[T,m]``
is the reference's unbatched signature.
"""
from __future__ import annotations

import torch

from ..core.base_block_state_space import BaseBlockSSM


class DifferentiableSimulator:
    def __init__(self, model: BaseBlockSSM, dt: float = 1.0, mode_interp: str = "linear", start_time: float = 0.0,
                 solver: str = "euler"):
        self.model = model
        self.dt = dt
        self.mode_interp = mode_interp      # dead field in the reference too (simulator.py:15)
        self.start_time = start_time
        # The reference integrates with explicit Euler only (simulator.py:40).  "rk4" is an extension for
        # models on the general fx/fy path (NeuralODE, BASELINE.json config 5).
        if solver not in ("euler", "rk4"):
            raise ValueError("solver must be 'euler' or 'rk4'")
        self.solver = solver

    # Flax: simulator.init(key, x0, inputs) -> {'params': {'model': ...}}
    def init(self, key, states_init, inputs):
        inner = self.model.init(key, _to_cuda(states_init), inputs)
        return {"params": {"model": inner["params"]}}

    def apply(self, variables, states_init, inputs):
        """xsol[T,(N,)n], ysol[T,(N,)p] = simulator.apply(params, x0, inputs)."""
        model = self.model
        if isinstance(states_init, (int, float)):
            assert model.state_dim == 1                                   # simulator.py:33-34
            states_init = torch.tensor([float(states_init)])
        x0 = _to_cuda(states_init)
        u = _to_cuda(inputs, x0.device)
        batched = x0.dim() == 2
        if not batched:
            x0 = x0.reshape(1, model.state_dim)
            if u.dim() != 2:
                raise ValueError(f"unbatched call expects inputs[T,{model.input_dim}], got {tuple(u.shape)}")
            u = u[:, None, :]
        params = variables["params"]["model"]
        if self.solver == "rk4":
            if not getattr(model, "supports_rk4", False):
                raise ValueError(f"solver='rk4' is available for models with general fx/fy dynamics (NeuralODE); "
                                 f"{type(model).__name__} is stepped with explicit Euler like the reference (simulator.py:40)")
            xsol, ysol = model._rollout(params, x0, u, float(self.dt), False, solver="rk4")
        else:
            xsol, ysol = model._rollout(params, x0, u, float(self.dt), False)   # simulator.py:40 always Euler
        if not batched:
            xsol, ysol = xsol[:, 0], ysol[:, 0]
        return xsol, ysol

    __call__ = apply


def _to_cuda(t, device=None):
    t = torch.as_tensor(t)
    if t.dtype != torch.float32:
        t = t.to(torch.float32)
    if not t.is_cuda:
        t = t.to(device or "cuda")
    return t
