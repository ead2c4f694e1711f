"""Mirror of dynax/core/discrete_block_state_space.py:5-54."""
from ._linear import LinearSSMMixin
from .base_block_state_space import BaseDiscreteBlockSSM


class DiscreteLinearSSM(LinearSSMMixin, BaseDiscreteBlockSSM):
    """x_{k+1} = A x_k + B u_k,  y_k = C x_k + D u_k  with learnable dense A,B,C,D."""
