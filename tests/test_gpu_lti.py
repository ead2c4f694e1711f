"""GPU parity: dense linear block SSMs (Discrete/ContinuousLinearSSM) forward + VJP vs the oracle."""
import numpy as np
import pytest

from oracle import dynax_oracle as orc
from tests.util import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def T():
    import torch
    return torch


@pytest.fixture(scope="module")
def ops():
    from dynax_b200 import ops
    return ops


def cu(T, a):
    return T.as_tensor(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def make(N, Tn, n, m, p, seed, L=None):
    rng = np.random.default_rng(seed)
    L = N if L is None else L
    A = (-np.eye(n)[None] + 0.3 * rng.normal(size=(L, n, n))).astype(np.float32)
    B = rng.normal(size=(L, n, m)).astype(np.float32)
    C = rng.normal(size=(L, p, n)).astype(np.float32)
    D = rng.normal(size=(L, p, m)).astype(np.float32)
    x0 = rng.normal(size=(N, n)).astype(np.float32)
    u = rng.normal(size=(Tn, N, m)).astype(np.float32)
    return A, B, C, D, x0, u


DIMS = [(1, 1, 1), (2, 1, 1), (2, 2, 1), (2, 2, 2), (3, 1, 1), (3, 3, 1), (3, 3, 3), (3, 5, 1), (4, 1, 1), (4, 2, 2),
        (4, 3, 4), (4, 4, 4)]


@pytest.mark.parametrize("n,m,p", DIMS)
def test_forward_all_dims(T, ops, n, m, p):
    N, Tn, dt = 260, 150, 0.05
    A, B, C, D, x0, u = make(N, Tn, n, m, p, 41)
    xs, ys, xf = ops.lti_forward(*(cu(T, v) for v in (A, B, C, D, x0, u)), dt, emit_final=True)
    xr, yr = orc.lti_rollout(A, B, C, D, x0, u, dt, np.float64)
    # zero-mean signals: inf-norm relative error (floor=1.0); 1e-5 is ~170 fp32 ulps of the signal scale
    assert rel_err(xs.cpu().numpy(), xr, floor=1.0) < 1e-5 and rel_err(ys.cpu().numpy(), yr, floor=1.0) < 1e-5
    assert np.array_equal(xf.cpu().numpy(), xs[-1].cpu().numpy())


@pytest.mark.parametrize("discrete", [False, True])
def test_forward_ragged_shared_discrete(T, ops, discrete):
    N, Tn, dt = 131, 90, 0.05
    n, m, p = 4, 3, 4                       # tests/test_forward_simulation.py:17-20
    A, B, C, D, x0, u = make(N, Tn, n, m, p, 42, L=1)
    A *= 0.5
    xs, ys, _ = ops.lti_forward(*(cu(T, v) for v in (A, B, C, D, x0, u[:, :1])), dt, discrete=discrete)
    xr, yr = orc.lti_rollout(A, B, C, D, x0, u[:, :1], dt, np.float64, discrete=discrete)
    assert rel_err(xs.cpu().numpy(), xr, floor=1.0) < 1e-5 and rel_err(ys.cpu().numpy(), yr, floor=1.0) < 1e-5


@pytest.mark.parametrize("n,m,p", [(4, 3, 4), (3, 5, 1), (2, 2, 2), (1, 1, 1), (4, 4, 4)])
@pytest.mark.parametrize("discrete", [False, True])
def test_vjp(T, ops, n, m, p, discrete):
    N, Tn, dt = 140, 120, 0.05
    A, B, C, D, x0, u = make(N, Tn, n, m, p, 43)
    if discrete:
        A *= 0.5
    rng = np.random.default_rng(44)
    gx = rng.normal(size=(Tn, N, n)).astype(np.float32)
    gy = rng.normal(size=(Tn, N, p)).astype(np.float32)
    out = ops.lti_vjp(*(cu(T, v) for v in (A, B, C, D, x0, u)), dt, cu(T, gx), cu(T, gy), discrete=discrete)
    r = orc.lti_vjp(A, B, C, D, x0, u, dt, gx, gy, np.float64, discrete=discrete)
    for got, key in zip(out, ("gA", "gB", "gC", "gD", "gx0", "gu")):
        assert rel_err(got.cpu().numpy(), r[key], floor=1.0) < 1e-4, key


@pytest.mark.parametrize("n,m,p,N", [(4, 3, 4, 19201), (4, 3, 4, 19396), (4, 4, 4, 19203), (4, 4, 4, 19364)])
def test_vjp_wide_tile(T, ops, n, m, p, N, monkeypatch):
    """Large dense models on launches of more than one 128-system CTA per SM take 192-/160-system tiles
    (lti_api.cu): parity with the fp64 oracle, ragged and 16-byte-aligned N, and the same numbers as the
    128-system tile (per-system outputs bit for bit: the per-thread arithmetic does not depend on the tile)."""
    Tn, dt = 37, 0.05
    A, B, C, D, x0, u = make(N, Tn, n, m, p, 48)
    rng = np.random.default_rng(49)
    gx = rng.normal(size=(Tn, N, n)).astype(np.float32)
    gy = rng.normal(size=(Tn, N, p)).astype(np.float32)
    a = [cu(T, v) for v in (A, B, C, D, x0, u)]
    out = ops.lti_vjp(*a, dt, cu(T, gx), cu(T, gy), time_chunks=0)
    r = orc.lti_vjp(A, B, C, D, x0, u, dt, gx, gy, np.float64)
    for got, key in zip(out, ("gA", "gB", "gC", "gD", "gx0", "gu")):
        assert rel_err(got.cpu().numpy(), r[key], floor=1.0) < 1e-4, key
    monkeypatch.setenv("DYNAX_LTI_VJP_CFG", "0")
    ref = ops.lti_vjp(*a, dt, cu(T, gx), cu(T, gy), time_chunks=0)
    for got, want in zip(out, ref):
        assert T.equal(got, want)
    monkeypatch.delenv("DYNAX_LTI_VJP_CFG")
    # one shared set of matrices: gradients summed over the systems (fp64 warp sums + atomics)
    sh = [cu(T, v[:1]) for v in (A, B, C, D)] + a[4:]
    A1, B1, C1, D1 = (np.repeat(v[:1], N, 0) for v in (A, B, C, D))
    outs = ops.lti_vjp(*sh, dt, cu(T, gx), cu(T, gy), time_chunks=0)
    rs = orc.lti_vjp(A1, B1, C1, D1, x0, u, dt, gx, gy, np.float64)
    for got, key in zip(outs[:4], ("gA", "gB", "gC", "gD")):
        assert rel_err(got.cpu().numpy()[0], rs[key].sum(0), floor=1.0) < 1e-4, key


@pytest.mark.parametrize("n,m,p", DIMS)
@pytest.mark.parametrize("Tn", [1, 7, 8, 9, 65, 203])
def test_vjp_forward_sweep_substages(T, ops, n, m, p, Tn):
    """The forward sweep of the general-cotangent kernel runs through sub-stages of the stage buffers (AdjCfg::NSUB, up to
    four per stage: the deeper ring wraps several times at T = 203) and hands the buffers back before the reverse sweep:
    every instantiated shape, horizons of one segment, one segment + 1 step, and many segments, ragged N."""
    N, dt = 131, 0.05
    A, B, C, D, x0, u = make(N, Tn, n, m, p, 60 + Tn)
    rng = np.random.default_rng(61 + Tn)
    gx = rng.normal(size=(Tn, N, n)).astype(np.float32)
    gy = rng.normal(size=(Tn, N, p)).astype(np.float32)
    out = ops.lti_vjp(*(cu(T, v) for v in (A, B, C, D, x0, u)), dt, cu(T, gx), cu(T, gy), time_chunks=0)
    r = orc.lti_vjp(A, B, C, D, x0, u, dt, gx, gy, np.float64)
    for got, key in zip(out, ("gA", "gB", "gC", "gD", "gx0", "gu")):
        assert rel_err(got.cpu().numpy(), r[key], floor=1.0) < 1e-4, key


def test_vjp_shared_matrices_sums_over_systems(T, ops):
    N, Tn, dt = 256, 64, 0.05
    A, B, C, D, x0, u = make(N, Tn, 3, 3, 1, 45, L=1)
    gy = np.random.default_rng(46).normal(size=(Tn, N, 1)).astype(np.float32)
    out = ops.lti_vjp(*(cu(T, v) for v in (A, B, C, D, x0, u)), dt, None, cu(T, gy))
    r = orc.lti_vjp(A, B, C, D, x0, u, dt, None, gy, np.float64)
    for got, key in zip(out[:4], ("gA", "gB", "gC", "gD")):
        assert rel_err(got.cpu().numpy()[0], r[key].sum(0)) < 1e-4, key


def test_unsupported_dims(T, ops):
    from dynax_b200._lib import DynaxError
    A, B, C, D, x0, u = make(4, 5, 5, 2, 1, 47)
    with pytest.raises(DynaxError) as e:
        ops.lti_forward(*(cu(T, v) for v in (A, B, C, D, x0, u)), 0.1)
    assert e.value.code == -2


@pytest.mark.parametrize("n,m,p,N,Tn,chunks", [(4, 3, 4, 1, 2000, -1), (3, 5, 1, 130, 700, 7), (2, 1, 1, 33, 257, 4),
                                               (4, 4, 4, 64, 900, -1), (1, 1, 1, 5, 300, 3), (3, 3, 3, 260, 512, 16)])
def test_forward_time_parallel(T, ops, n, m, p, N, Tn, chunks):
    """Time-parallel (chunked) dense-LTI forward: same outputs as the serial kernel within the trajectory tolerance."""
    dt = 0.02
    A, B, C, D, x0, u = make(N, Tn, n, m, p, 47)
    a = [cu(T, v) for v in (A, B, C, D, x0, u)]
    xs, ys, xf = ops.lti_forward(*a, dt, emit_final=True, time_chunks=chunks)
    x1, y1, _ = ops.lti_forward(*a, dt, emit_final=True, time_chunks=0)
    xr, yr = orc.lti_rollout(A, B, C, D, x0, u, dt, np.float64)
    assert rel_err(xs.cpu().numpy(), xr, floor=1.0) < 1e-5 and rel_err(ys.cpu().numpy(), yr, floor=1.0) < 1e-5
    assert rel_err(xs.cpu().numpy(), x1.cpu().numpy(), floor=1.0) < 1e-5
    assert T.equal(xf, xs[-1])
    # one set of matrices, one shared input series
    A1, B1, C1, D1 = A[:1], B[:1], C[:1], D[:1]
    xs, ys, _ = ops.lti_forward(*(cu(T, v) for v in (A1, B1, C1, D1, x0, u[:, :1])), dt, time_chunks=chunks)
    xr, yr = orc.lti_rollout(A1, B1, C1, D1, x0, u[:, :1], dt, np.float64)
    assert rel_err(xs.cpu().numpy(), xr, floor=1.0) < 1e-5 and rel_err(ys.cpu().numpy(), yr, floor=1.0) < 1e-5


@pytest.mark.parametrize("n,m,p,N,Tn,chunks,shared", [(4, 3, 4, 1, 1500, -1, False), (3, 5, 1, 130, 600, 7, False),
                                                      (2, 2, 2, 33, 257, 4, True), (4, 4, 4, 8, 640, -1, True),
                                                      (1, 1, 1, 5, 300, 3, False)])
def test_vjp_time_parallel(T, ops, n, m, p, N, Tn, chunks, shared):
    """Time-parallel (chunked) dense-LTI VJP vs the fp64 hand adjoint and the serial kernel."""
    dt = 0.02
    A, B, C, D, x0, u = make(N, Tn, n, m, p, 48, L=1 if shared else None)
    rng = np.random.default_rng(49)
    gx = rng.normal(size=(Tn, N, n)).astype(np.float32)
    gy = rng.normal(size=(Tn, N, p)).astype(np.float32)
    a = [cu(T, v) for v in (A, B, C, D, x0, u)]
    out = ops.lti_vjp(*a, dt, cu(T, gx), cu(T, gy), time_chunks=chunks)
    ser = ops.lti_vjp(*a, dt, cu(T, gx), cu(T, gy), time_chunks=0)
    r = orc.lti_vjp(A, B, C, D, x0, u, dt, gx, gy, np.float64)
    for got, s_, key in zip(out, ser, ("gA", "gB", "gC", "gD", "gx0", "gu")):
        ref = r[key].sum(0, keepdims=True) if (shared and key in ("gA", "gB", "gC", "gD")) else r[key]
        assert rel_err(got.cpu().numpy().reshape(ref.shape), ref, floor=1.0) < 1e-4, key
        assert rel_err(got.cpu().numpy(), s_.cpu().numpy(), floor=1.0) < 1e-4, key


@pytest.mark.parametrize("n,m,p", [(4, 3, 4), (3, 5, 1), (2, 1, 1)])
def test_vjp_shared_input_series(T, ops, n, m, p):
    """g_u of a shared input series u[T,1,m]: summed over systems inside the kernel (dense models)."""
    N, Tn, dt = 150, 60, 0.05
    A, B, C, D, x0, u = make(N, Tn, n, m, p, 47)
    u1 = u[:, :1]
    rng = np.random.default_rng(48)
    gx = rng.normal(size=(Tn, N, n)).astype(np.float32)
    gy = rng.normal(size=(Tn, N, p)).astype(np.float32)
    out = ops.lti_vjp(*(cu(T, v) for v in (A, B, C, D, x0, u1)), dt, cu(T, gx), cu(T, gy), time_chunks=0)
    r = orc.lti_vjp(A, B, C, D, x0, np.repeat(u1, N, 1), dt, gx, gy, np.float64)
    assert out[5].shape == (Tn, 1, m)
    assert rel_err(out[5].cpu().numpy()[:, 0], r["gu"].sum(1), floor=1.0) < 1e-4
    for got, key in zip(out[:5], ("gA", "gB", "gC", "gD", "gx0")):
        assert rel_err(got.cpu().numpy(), r[key], floor=1.0) < 1e-4, key
