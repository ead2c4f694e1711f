"""Mirror of dynax/core/continuous_block_state_space.py:5-54."""
from ._linear import LinearSSMMixin
from .base_block_state_space import BaseContinuousBlockSSM


class ContinuousLinearSSM(LinearSSMMixin, BaseContinuousBlockSSM):
    """dx/dt = A x + B u,  y = C x + D u  with learnable dense A,B,C,D."""
