"""GPU parity: time-parallel (chunked) RC forward rollout vs the serial kernel and the fp64 oracle."""
import numpy as np
import pytest

from oracle import dynax_oracle as orc
from tests.util import TRAJ_RTOL, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def T():
    import torch
    return torch


def cu(T, a):
    return T.as_tensor(np.ascontiguousarray(a, dtype=np.float32)).cuda()


@pytest.mark.parametrize("N,Tn,chunks", [(64, 8760, -1), (1, 8760, 60), (130, 1000, 7), (33, 257, 4), (8, 100, 1), (5, 64, 64)])
def test_chunked_matches_fp64_and_serial(T, N, Tn, chunks):
    from dynax_b200 import ops
    th, x0, u = orc.synth_theta(N, 111), orc.synth_x0(N, 111), orc.synth_u(Tn, N, 111)
    a = [cu(T, v) for v in (th, x0, u)]
    xs, ys, xf = ops.rc_forward(*a, 3600.0, emit_final=True, time_chunks=chunks)
    x1, y1, xf1 = ops.rc_forward(*a, 3600.0, emit_final=True, time_chunks=0)
    x64, y64 = orc.rc_rollout(th, x0, u, 3600.0, np.float64)
    assert rel_err(xs.cpu().numpy(), x64) < TRAJ_RTOL and rel_err(ys.cpu().numpy(), y64) < TRAJ_RTOL
    assert rel_err(xs.cpu().numpy(), x1.cpu().numpy()) < TRAJ_RTOL
    assert rel_err(xf.cpu().numpy(), xf1.cpu().numpy()) < TRAJ_RTOL
    assert T.equal(xf, xs[-1])


def test_chunked_shared_inputs_and_flags(T):
    from dynax_b200 import ops
    N, Tn = 40, 3000
    th, x0, u = orc.synth_theta(N, 112), orc.synth_x0(N, 112), orc.synth_u(Tn, N, 112)
    xs, ys, _ = ops.rc_forward(cu(T, th[0]), cu(T, x0), cu(T, u[:, 0]), 3600.0, time_chunks=12)   # shared theta + u
    x64, y64 = orc.rc_rollout(th[:1], x0, u[:, :1], 3600.0, np.float64)
    assert rel_err(xs.cpu().numpy(), x64) < TRAJ_RTOL and rel_err(ys.cpu().numpy(), y64) < TRAJ_RTOL
    xo, yo, xf = ops.rc_forward(cu(T, th), cu(T, x0), cu(T, u), 3600.0, emit_x=False, emit_y=True, emit_final=True, time_chunks=9)
    x64, y64 = orc.rc_rollout(th, x0, u, 3600.0, np.float64)
    assert xo is None and rel_err(yo.cpu().numpy(), y64) < TRAJ_RTOL and rel_err(xf.cpu().numpy(), x64[-1]) < TRAJ_RTOL
