from .base_block_state_space import BaseBlockSSM, BaseContinuousBlockSSM, BaseDiscreteBlockSSM
from .continuous_block_state_space import ContinuousLinearSSM
from .discrete_block_state_space import DiscreteLinearSSM

__all__ = ["BaseBlockSSM", "BaseDiscreteBlockSSM", "BaseContinuousBlockSSM", "DiscreteLinearSSM",
           "ContinuousLinearSSM"]
