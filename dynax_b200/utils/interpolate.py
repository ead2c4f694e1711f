"""Mirror of dynax/utils/interpolate.py:26-68 (the two implemented interpolators)."""
from __future__ import annotations

import torch

from .. import ops


class _Interp:
    mode = None

    def __init__(self, ts, xs):
        self.ts = torch.as_tensor(ts, dtype=torch.float32).cuda()
        self.xs = torch.as_tensor(xs, dtype=torch.float32).cuda()
        assert self.ts.dim() == 1, "rank of ts has to be 1 !!!"      # interpolate.py:15
        assert self.xs.dim() == 2, "rank of xs has to be 2 !!!"      # interpolate.py:16

    def __call__(self, at):
        at = torch.as_tensor(at, dtype=torch.float32).to(self.xs.device)
        return ops.tabular_interp(self.ts, self.xs, at, self.mode)


class PiecewiseConstantInterpolation(_Interp):
    """order-0 map_coordinates: the NEAREST table row (interpolate.py:26-46)."""
    mode = "constant"

    def order(self):
        return 0


class LinearInterpolation(_Interp):
    """order-1 map_coordinates on the fractional row index (interpolate.py:48-68)."""
    mode = "linear"

    def order(self):
        return 1
