"""General ``fx/fy`` block SSM with MLP dynamics (the form examples/4-neural-ode.py:16-44 sketches):

    rhs = fx(x, u) = W3 tanh(W2 tanh(W1 [x;u] + b1) + b2) + b3          y = fy(x, u) = x[0:output_dim]

dispatched through ``BaseBlockSSM._call_state`` / ``_call_observation``'s general branch
(dynax/core/base_block_state_space.py:62-63,72-73).  Parameters follow Flax ``nn.Dense`` naming and
layout (``params/_fx/dense_{0,1,2}/{kernel[in,out],bias}``); the CUDA kernels take the transposed
(row-major [out,in]) matrices.  Instantiated for (state, input, output) = (3, 5, 1) and hidden widths 32, 64, 128.
"""
from __future__ import annotations

import torch

from .. import ops
from ..core._linear import lecun_normal
from ..core.base_block_state_space import BaseContinuousBlockSSM


class NeuralODE(BaseContinuousBlockSSM):
    state_dim = 3
    input_dim = 5
    output_dim = 1
    hidden = 64

    def __init__(self, state_dim=3, input_dim=5, output_dim=1, hidden=64, name=None):
        super().__init__(state_dim, input_dim, output_dim, name)
        self.hidden = hidden

    def init_params(self, seed, device):
        gen = torch.Generator().manual_seed(int(seed))
        dims = [(self.state_dim + self.input_dim, self.hidden), (self.hidden, self.hidden), (self.hidden, self.state_dim)]
        fx = {}
        for k, (i, o) in enumerate(dims):
            fx[f"dense_{k}"] = {"kernel": lecun_normal(gen, i, (i, o), device),
                                "bias": torch.zeros(o, dtype=torch.float32, device=device)}
        return {"_fx": fx}

    @staticmethod
    def matrices(params):
        fx = params["_fx"]
        out = []
        for k in range(3):
            out += [fx[f"dense_{k}"]["kernel"].transpose(-1, -2).contiguous(), fx[f"dense_{k}"]["bias"].contiguous()]
        return out  # W1, b1, W2, b2, W3, b3  (row-major [out,in])

    supports_rk4 = True

    def _rollout(self, params, x0, u, dt, discrete, solver="euler"):
        if discrete:
            # BaseBlockSSM.apply evaluates the model ONCE (rhs, y) as a T=1 launch with x_next = rhs
            # (base_block_state_space.py:51-57).  For the general fx/fy model: one Euler step of size 1 gives
            # x + rhs, so rhs = xsol - x; y comes from the pre-update state as in every rollout.
            if u.shape[0] != 1:
                raise NotImplementedError("NeuralODE is a continuous model: discrete stepping is the single evaluation of model.apply only")
            xsol, ysol = ops.node_rollout(*self.matrices(params), x0, u, 1.0, euler=True)
            return xsol - x0[None], ysol
        return ops.node_rollout(*self.matrices(params), x0, u, dt, euler=(solver == "euler"))
