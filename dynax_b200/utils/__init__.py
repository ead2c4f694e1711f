from .interpolate import LinearInterpolation, PiecewiseConstantInterpolation

__all__ = ["LinearInterpolation", "PiecewiseConstantInterpolation"]
