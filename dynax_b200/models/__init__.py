from .mlp import NeuralODE
from .RC import Continuous4R3C, Discrete4R3C

__all__ = ["Continuous4R3C", "Discrete4R3C", "NeuralODE"]
