from .simulator import DifferentiableSimulator

__all__ = ["DifferentiableSimulator"]
