"""Mirror of dynax/agents/tabular.py:7-44: a time table of actions / disturbances."""
from __future__ import annotations

from ..utils.interpolate import LinearInterpolation, PiecewiseConstantInterpolation


class Tabular:
    def __init__(self, ts, xs, mode: str = "linear"):
        if mode == "linear":
            self.interp = LinearInterpolation(ts, xs)
        elif mode == "constant":
            self.interp = PiecewiseConstantInterpolation(ts, xs)
        else:
            raise ValueError("Interpolation mode in Tabular has to be either 'linear' or 'constant' !!!")   # tabular.py:39
        self.ts, self.xs, self.mode = self.interp.ts, self.interp.xs, mode

    # Flax call shape of the reference tests: agent.init(key, t); agent.apply(params, t)
    def init(self, key, at):
        return {}

    def apply(self, variables, at):
        return self(at)

    def __call__(self, at):
        return self.interp(at)
