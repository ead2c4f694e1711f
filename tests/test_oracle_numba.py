"""The numba cross-check of the CPU baseline (oracle/numba_oracle.py) against the C/OpenMP port and
the fp64 NumPy oracle: two independently written all-fp32 BPTT restatements of the reference path agree with each other
far better than either agrees with fp64 truth."""
import numpy as np
import pytest

from oracle import dynax_oracle as orc

numba_oracle = pytest.importorskip("oracle.numba_oracle")


def test_numba_port_matches_c_port_and_fp64():
    from oracle import c_oracle
    N, Tn = 48, 700
    th, x0, u = orc.synth_theta(N, 5), orc.synth_x0(N, 5), orc.synth_u(Tn, N, 5)
    target = (22 + np.random.default_rng(6).normal(size=(Tn, N))).astype(np.float32)
    ln, gn = numba_oracle.rc_loss_grad(th, x0, u, target, 3600.0)
    lc, gc = c_oracle.rc_loss_grad(th, x0, u, target, 3600.0)
    l64, g64 = orc.rc_loss_grad(th, x0, u, target, 3600.0, None, None, np.float64)
    nw = lambda a, b: np.linalg.norm(a.astype(np.float64) - b, axis=1) / np.linalg.norm(b, axis=1)
    assert np.abs(ln - lc).max() / np.abs(lc).max() < 1e-5
    assert nw(gn, gc.astype(np.float64)).max() < 5e-4          # same algorithm, different compilers / contraction
    assert np.abs(ln - l64).max() / np.abs(l64).max() < 1e-4
    assert nw(gn, g64).max() < 5e-3                            # all-fp32 BPTT vs fp64 (SURVEY.md finding #9: ~1e-3)
