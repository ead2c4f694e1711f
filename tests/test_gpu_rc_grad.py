"""GPU parity: fused loss+gradient and general VJP of the RC rollout vs the fp64 oracle.

Tolerance (north_star): gradients rel 1e-4 -- applied as SURVEY.md section 7 H2 makes it
meaningful: norm-wise <= 1e-4 per system, and per element no further from fp64 truth than
max(1e-4*|g|, error of the all-fp32 reference-order BPTT).  Loss rel 1e-5.
"""
import numpy as np
import pytest

from oracle import c_oracle, dynax_oracle as orc
from tests.util import assert_grad_parity, assert_loss_parity, normwise, rel_err

pytestmark = pytest.mark.gpu

from tests.util import GRAD_RTOL  # noqa: E402


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


def make_case(N, Tn, seed, noise=0.1):
    th, x0, u = orc.synth_theta(N, seed), orc.synth_x0(N, seed), orc.synth_u(Tn, N, seed)
    if noise is None:
        return th, x0, u, (22 + np.random.default_rng(seed).normal(size=(Tn, N))).astype(np.float32)
    hidden = orc.synth_theta(N, seed + 100)
    _, y = orc.rc_rollout(hidden, x0, u, 3600.0, np.float64)
    target = (y[..., 0] + noise * np.random.default_rng(seed).normal(size=(Tn, N))).astype(np.float32)
    return th, x0, u, target


ALGO_ENV = {"tangent": "1", "adjoint": "0"}   # DYNAX_GRAD_ALGO: forward-mode single sweep / checkpointed adjoint


@pytest.mark.parametrize("algo", ["tangent", "adjoint"])
@pytest.mark.parametrize("N,Tn,noise", [(256, 8760, 0.1), (256, 8760, None), (130, 1000, 0.1), (7, 100, 1.0),
                                        (1, 17, 0.1), (5, 16, 0.1), (3, 1, 0.1)])
def test_loss_grad_vs_fp64(T, ops, monkeypatch, N, Tn, noise, algo):
    monkeypatch.setenv("DYNAX_GRAD_ALGO", ALGO_ENV[algo])
    # noise=0.1: target = a hidden model's output + N(0,0.1) (small residual: ill-conditioned, the
    # realistic near-optimum regime); noise=None: target = 22 + N(0,1) (large residual)
    th, x0, u, target = make_case(N, Tn, 31, noise)
    th[0, 0] = 4.0e4            # above ub
    if N > 1:
        th[1, 4] = 0.45         # below lb
    # (time_chunks=0: the serial-in-time kernels are what this file tests; the chunked ones have test_gpu_chunked.py)
    loss, grad, gx0 = ops.rc_loss_grad(cu(T, th), cu(T, x0), cu(T, u), cu(T, target), 3600.0,
                                       orc.RC_LB, orc.RC_UB, want_gx0=True, time_chunks=0)
    l64, g64 = orc.rc_loss_grad(th, x0, u, target, 3600.0, orc.RC_LB, orc.RC_UB, np.float64)
    l32, g32 = c_oracle.rc_loss_grad(th, x0, u, target, 3600.0, orc.RC_LB, orc.RC_UB)
    lbud, gbud = orc.rc_output_sensitivity_budget(th, x0, u, target, 3600.0)
    assert_loss_parity(loss.cpu().numpy(), l64, lbud)
    assert_grad_parity(grad.cpu().numpy(), g64, gbud, g32 if N >= 64 else None)
    # d loss / d x0 against the oracle's general VJP (tangent kernel: three more sensitivity columns)
    gy = (2.0 / Tn) * (orc.rc_rollout(th, x0, u, 3600.0, np.float64)[1][..., 0] - target)
    _, gx0_64, _ = orc.rc_vjp(th, x0, u, 3600.0, None, gy[..., None])
    # yardstick for d loss/d x0: the same adjoint in fp32 reference order (NumPy oracle)
    y32 = orc.rc_rollout(th, x0, u, 3600.0, np.float32)[1][..., 0]
    _, gx0_32, _ = orc.rc_vjp(th, x0, u, 3600.0, None, ((2.0 / Tn) * (y32 - target))[..., None].astype(np.float32), np.float32)
    e_ours, e_ref = rel_err(gx0.cpu().numpy(), gx0_64, floor=1.0), rel_err(gx0_32, gx0_64, floor=1.0)
    assert e_ours <= max(GRAD_RTOL, 1.25 * e_ref), (e_ours, e_ref)


@pytest.mark.parametrize("cfg", ["0", "1"])
def test_adj_tile_configs(T, ops, monkeypatch, cfg):
    th, x0, u, target = make_case(300, 777, 32)
    monkeypatch.setenv("DYNAX_GRAD_ALGO", "0")
    monkeypatch.setenv("DYNAX_ADJ_CFG", cfg)
    loss, grad, _ = ops.rc_loss_grad(cu(T, th), cu(T, x0), cu(T, u), cu(T, target), 3600.0)
    l64, g64 = orc.rc_loss_grad(th, x0, u, target, 3600.0, None, None, np.float64)
    lbud, gbud = orc.rc_output_sensitivity_budget(th, x0, u, target, 3600.0)
    assert_loss_parity(loss.cpu().numpy(), l64, lbud)
    assert_grad_parity(grad.cpu().numpy(), g64, gbud)


@pytest.mark.parametrize("cfg,split", [("0", "0"), ("2", "0"), ("3", "0"), ("0", "1"), ("0", "2")])
def test_tangent_tile_configs(T, ops, monkeypatch, cfg, split):
    th, x0, u, target = make_case(300, 777, 32)
    monkeypatch.setenv("DYNAX_GRAD_ALGO", "1")
    monkeypatch.setenv("DYNAX_FS_CFG", cfg)
    monkeypatch.setenv("DYNAX_FS_SPLIT", split)
    loss, grad, _ = ops.rc_loss_grad(cu(T, th), cu(T, x0), cu(T, u), cu(T, target), 3600.0, time_chunks=0)
    l64, g64 = orc.rc_loss_grad(th, x0, u, target, 3600.0, None, None, np.float64)
    lbud, gbud = orc.rc_output_sensitivity_budget(th, x0, u, target, 3600.0)
    assert_loss_parity(loss.cpu().numpy(), l64, lbud)
    assert_grad_parity(grad.cpu().numpy(), g64, gbud)


def test_tangent_split_roles_bit_identical(T, ops, monkeypatch):
    """1 or 2 threads per system (SPLIT, rc_fwdsens.cuh): every role steps x with the same IEEE operations and owns
    whole sensitivity columns, so loss and gradient do not depend on the split -- bit for bit."""
    th, x0, u, target = make_case(1000, 1237, 40)
    a = [cu(T, v) for v in (th, x0, u, target)]
    monkeypatch.setenv("DYNAX_GRAD_ALGO", "1")
    monkeypatch.setenv("DYNAX_FS_SPLIT", "1")
    l0, g0, _ = ops.rc_loss_grad(*a, 3600.0, orc.RC_LB, orc.RC_UB, time_chunks=0)
    for split in ("2",):
        monkeypatch.setenv("DYNAX_FS_SPLIT", split)
        l1, g1, _ = ops.rc_loss_grad(*a, 3600.0, orc.RC_LB, orc.RC_UB, time_chunks=0)
        assert T.equal(l0, l1) and T.equal(g0, g1), split
    monkeypatch.delenv("DYNAX_FS_SPLIT")
    # ... and the trajectory inside the loss kernels is the forward kernel's, bit for bit (RCModel::rhs fixes the
    # order of operations), so the loss equals the MSE of the emitted ysol up to the rounding of the sum
    _, ys, _ = ops.rc_forward(a[0], a[1], a[2], 3600.0, time_chunks=0)
    mse = ((ys[..., 0].double() - a[3].double()) ** 2).mean(0)
    pen = T.as_tensor(orc.bound_penalty(th, orc.RC_LB, orc.RC_UB), device="cuda")
    assert T.allclose(l0.double(), mse + pen, rtol=2e-6, atol=0)   # fp32 partial sums over 32 steps, fp32 output
    monkeypatch.setenv("DYNAX_GRAD_ALGO", "0")
    l2, _, _ = ops.rc_loss_grad(*a, 3600.0, orc.RC_LB, orc.RC_UB, time_chunks=0)
    assert T.allclose(l2.double(), mse + pen, rtol=2e-6, atol=0)   # the adjoint kernel steps the same trajectory


def test_tangent_and_adjoint_agree(T, ops, monkeypatch):
    """The two gradient kernels are different fp32 evaluations of one quantity: both sit inside the
    parity band around the fp64 oracle, so they differ by at most twice that band."""
    th, x0, u, target = make_case(1024, 2000, 39)
    th[3, 2] = 5e6              # clipped by ub: penalty path in both
    a = [cu(T, v) for v in (th, x0, u, target)]
    monkeypatch.setenv("DYNAX_GRAD_ALGO", "1")
    l1, g1, _ = ops.rc_loss_grad(*a, 3600.0, orc.RC_LB, orc.RC_UB, time_chunks=0)
    monkeypatch.setenv("DYNAX_GRAD_ALGO", "0")
    l2, g2, _ = ops.rc_loss_grad(*a, 3600.0, orc.RC_LB, orc.RC_UB, time_chunks=0)
    assert not T.equal(g1, g2)          # really two code paths
    l64, g64 = orc.rc_loss_grad(th, x0, u, target, 3600.0, orc.RC_LB, orc.RC_UB, np.float64)
    lbud, gbud = orc.rc_output_sensitivity_budget(th, x0, u, target, 3600.0)
    dl = np.abs((l1 - l2).double().cpu().numpy())
    assert (dl <= 2 * (1e-5 * np.abs(l64) + lbud)).all(), (dl / (1e-5 * np.abs(l64) + lbud)).max()
    dg = np.abs((g1 - g2).double().cpu().numpy())
    assert (dg <= 2 * (GRAD_RTOL * np.abs(g64) + gbud)).all(), (dg / (GRAD_RTOL * np.abs(g64) + gbud)).max()


@pytest.mark.parametrize("algo", ["tangent", "adjoint"])
def test_no_tma_path_and_ragged(T, ops, monkeypatch, algo):
    monkeypatch.setenv("DYNAX_GRAD_ALGO", ALGO_ENV[algo])
    th, x0, u, target = make_case(261, 300, 33)          # N % 4 != 0 -> plain ld path
    a = [cu(T, v) for v in (th, x0, u, target)]
    loss, grad, _ = ops.rc_loss_grad(*a, 3600.0)
    l64, g64 = orc.rc_loss_grad(th, x0, u, target, 3600.0, None, None, np.float64)
    lbud, gbud = orc.rc_output_sensitivity_budget(th, x0, u, target, 3600.0)
    assert_loss_parity(loss.cpu().numpy(), l64, lbud)
    assert_grad_parity(grad.cpu().numpy(), g64, gbud)
    th, x0, u, target = make_case(256, 300, 34)
    a = [cu(T, v) for v in (th, x0, u, target)]
    l1, g1, _ = ops.rc_loss_grad(*a, 3600.0)
    monkeypatch.setenv("DYNAX_NO_BULK", "1")
    l2, g2, _ = ops.rc_loss_grad(*a, 3600.0)
    assert T.equal(l1, l2) and T.equal(g1, g2)


@pytest.mark.parametrize("algo", ["tangent", "adjoint"])
def test_shared_modes(T, ops, monkeypatch, algo):
    monkeypatch.setenv("DYNAX_GRAD_ALGO", ALGO_ENV[algo])
    N, Tn = 384, 501            # ragged last chunk
    th, x0, u, target = make_case(N, Tn, 35, None)
    # shared u + shared target, per-system theta  (config 3 variant ii)
    loss, grad, _ = ops.rc_loss_grad(cu(T, th), cu(T, x0), cu(T, u[:, :1]), cu(T, target[:, 0]), 3600.0)
    l64, g64 = orc.rc_loss_grad(th, x0, u[:, :1], target[:, :1], 3600.0, None, None, np.float64)
    lbud, gbud = orc.rc_output_sensitivity_budget(th, x0, u[:, :1], target[:, :1], 3600.0)
    assert_loss_parity(loss.cpu().numpy(), l64, lbud)
    assert_grad_parity(grad.cpu().numpy(), g64, gbud)
    # shared theta: gradient summed over systems, penalty counted once
    th1 = th[:1].copy(); th1[0, 6] = 60.0
    loss, grad, _ = ops.rc_loss_grad(cu(T, th1[0]), cu(T, x0), cu(T, u), cu(T, target), 3600.0, orc.RC_LB, orc.RC_UB)
    l64, g64 = orc.rc_loss_grad(np.repeat(th1, N, 0), x0, u, target, 3600.0, None, None, np.float64)
    pen = orc.bound_penalty(th1, orc.RC_LB, orc.RC_UB)[0]
    peng = orc.bound_penalty_grad(th1, orc.RC_LB, orc.RC_UB)[0]
    lbud, gbud = orc.rc_output_sensitivity_budget(np.repeat(th1, N, 0), x0, u, target, 3600.0)
    assert_loss_parity(loss.cpu().numpy()[0], l64.sum() + pen, lbud.sum())
    assert_grad_parity(grad.cpu().numpy(), g64.sum(0) + peng, gbud.sum(0))
    # per-system u, shared target
    loss, grad, _ = ops.rc_loss_grad(cu(T, th), cu(T, x0), cu(T, u), cu(T, target[:, 0]), 3600.0)
    l64, g64 = orc.rc_loss_grad(th, x0, u, target[:, :1], 3600.0, None, None, np.float64)
    lbud, gbud = orc.rc_output_sensitivity_budget(th, x0, u, target[:, :1], 3600.0)
    assert_loss_parity(loss.cpu().numpy(), l64, lbud)
    assert_grad_parity(grad.cpu().numpy(), g64, gbud)
    # shared u, per-system target, box penalty on
    th2 = th.copy(); th2[5, 3] = 50.0
    loss, grad, _ = ops.rc_loss_grad(cu(T, th2), cu(T, x0), cu(T, u[:, :1]), cu(T, target), 3600.0, orc.RC_LB, orc.RC_UB)
    l64, g64 = orc.rc_loss_grad(th2, x0, u[:, :1], target, 3600.0, orc.RC_LB, orc.RC_UB, np.float64)
    lbud, gbud = orc.rc_output_sensitivity_budget(th2, x0, u[:, :1], target, 3600.0)
    assert_loss_parity(loss.cpu().numpy(), l64, lbud)
    assert_grad_parity(grad.cpu().numpy(), g64, gbud)


@pytest.mark.parametrize("N,Tn,use_gx,use_gy", [(200, 400, True, True), (33, 129, True, False), (64, 50, False, True)])
def test_general_vjp(T, ops, N, Tn, use_gx, use_gy):
    rng = np.random.default_rng(36)
    th, x0, u = orc.synth_theta(N, 36), orc.synth_x0(N, 36), orc.synth_u(Tn, N, 36)
    gx = rng.normal(size=(Tn, N, 3)).astype(np.float32) if use_gx else None
    gy = rng.normal(size=(Tn, N, 1)).astype(np.float32) if use_gy else None
    gth, gx0, gu = ops.rc_vjp(cu(T, th), cu(T, x0), cu(T, u), 3600.0,
                              None if gx is None else cu(T, gx), None if gy is None else cu(T, gy))
    rth, rx0, ru = orc.rc_vjp(th, x0, u, 3600.0, gx, gy, np.float64)
    # random zero-mean cotangents cancel heavily; yardstick = the same adjoint evaluated in fp32
    # reference order (NumPy oracle): we must be at least as close to fp64 truth as it is, or 1e-4
    fth, fx0, fu = orc.rc_vjp(th, x0, u, 3600.0, gx, gy, np.float32)
    no, nr = normwise(gth.cpu().numpy(), rth), normwise(fth, rth)
    assert no.max() <= max(GRAD_RTOL, 1.25 * nr.max()) and np.median(no) <= max(GRAD_RTOL, np.median(nr)), (no.max(), nr.max())
    assert rel_err(gx0.cpu().numpy(), rx0, floor=1.0) < max(GRAD_RTOL, 1.25 * rel_err(fx0, rx0, floor=1.0))
    assert rel_err(gu.cpu().numpy(), ru, floor=1.0) < max(GRAD_RTOL, 1.25 * rel_err(fu, ru, floor=1.0))


def test_autograd_function_matches_fused(T, ops):
    """jax.grad-style use: loss built by the caller on top of simulator outputs."""
    N, Tn = 128, 300
    th, x0, u, target = make_case(N, Tn, 37, None)
    tth = cu(T, th).requires_grad_(True)
    tx0 = cu(T, x0).requires_grad_(True)
    tu = cu(T, u).requires_grad_(True)
    xs, ys = ops.rc_rollout(tth, tx0, tu, 3600.0)
    loss = ((ys[..., 0] - cu(T, target)) ** 2).mean(0)
    loss.sum().backward()
    l64, g64 = orc.rc_loss_grad(th, x0, u, target, 3600.0, None, None, np.float64)
    _, gbud = orc.rc_output_sensitivity_budget(th, x0, u, target, 3600.0)
    assert_grad_parity(tth.grad.cpu().numpy(), g64, gbud)
    lf, gf, gx0f = ops.rc_loss_grad(cu(T, th), cu(T, x0), cu(T, u), cu(T, target), 3600.0, want_gx0=True)
    assert_grad_parity(gf.cpu().numpy(), g64, gbud)
    # fused kernel and rollout+autograd are two fp32 evaluations of the same gradient
    assert (np.abs(gf.cpu().numpy() - tth.grad.cpu().numpy()) <= 2 * (GRAD_RTOL * np.abs(g64) + gbud)).all()
    assert rel_err(gx0f.cpu().numpy(), tx0.grad.cpu().numpy(), floor=1.0) < 2 * GRAD_RTOL
    assert tu.grad is not None and tu.grad.shape == tu.shape


def test_grad_property_large(T, ops):
    """Full-machine tile: shared-theta gradient == sum of per-system gradients (fp64-folded), and
    64 random systems vs the fp64 oracle."""
    N, Tn = 148 * 2 * 128, 200
    g = T.Generator(device="cuda").manual_seed(6)
    th1 = cu(T, orc.RC_NOMINAL[None])
    th = th1.expand(N, 7).contiguous()
    x0 = cu(T, [[20.0, 30.0, 26.0]]) + T.empty((N, 3), device="cuda").uniform_(-1, 1, generator=g)
    u = T.empty((Tn, N, 5), device="cuda").uniform_(0, 1, generator=g)
    u[..., 0] = 15 + 10 * u[..., 0]
    target = 20 + 2 * T.empty((Tn, N), device="cuda").uniform_(-1, 1, generator=g)
    l_i, g_i, _ = ops.rc_loss_grad(th, x0, u, target, 3600.0)
    l_s, g_s, _ = ops.rc_loss_grad(th1[0], x0, u, target, 3600.0)
    assert np.allclose(l_s.item(), l_i.double().sum().item(), rtol=1e-5)
    gs = g_i.double().sum(0).cpu().numpy()
    assert np.linalg.norm(g_s.cpu().numpy() - gs) / np.linalg.norm(gs) < 1e-5
    idx = np.random.default_rng(1).choice(N, 64, replace=False)
    l64, g64 = orc.rc_loss_grad(th[idx].cpu().numpy(), x0[idx].cpu().numpy(), u[:, idx].cpu().numpy(),
                                target[:, idx].cpu().numpy(), 3600.0, None, None, np.float64)
    lbud, gbud = orc.rc_output_sensitivity_budget(th[idx].cpu().numpy(), x0[idx].cpu().numpy(), u[:, idx].cpu().numpy(),
                                                  target[:, idx].cpu().numpy(), 3600.0)
    assert_loss_parity(l_i[idx].cpu().numpy(), l64, lbud)
    assert_grad_parity(g_i[idx].cpu().numpy(), g64, gbud)


@pytest.mark.parametrize("algo", ["1", "0"])
def test_shared_theta_sum_is_reproducible_to_fp32_rounding(T, ops, monkeypatch, algo):
    """One parameter vector for all systems: the per-system fp64 gradients are summed with one fp64 atomic per warp, so the
    ORDER of the sum varies from run to run.  The contract: the fp32 results of two identical calls agree to 1e-6 relative
    (fp64 partial sums of ~1e4 systems differ at 1e-13; what is left is the final rounding to fp32) -- stated here because
    a caller comparing runs bit for bit would otherwise be surprised."""
    monkeypatch.setenv("DYNAX_GRAD_ALGO", algo)
    N, Tn = 20000, 96
    th, x0, u, target = make_case(N, Tn, 77)
    a = [cu(T, v) for v in (th[0], x0, u, target)]
    runs = [ops.rc_loss_grad(*a, 3600.0, time_chunks=0) for _ in range(4)]
    l0, g0 = runs[0][0].double(), runs[0][1].double()
    for l, g, _ in runs[1:]:
        assert float(((l.double() - l0).abs() / l0.abs()).max()) <= 1e-6
        assert float(((g.double() - g0).abs() / g0.abs().clamp_min(1e-30)).max()) <= 1e-6


def test_workspace_and_errors(T, ops):
    from dynax_b200 import _lib
    L = _lib.lib()
    th, x0, u, target = (cu(T, v) for v in make_case(8, 20, 38))
    loss = T.empty(8, device="cuda"); grad = T.empty(8, 7, device="cuda")
    # the default (forward-mode) kernel needs no scratch for per-system parameters, 1 KB for shared ones; the
    # checkpointed adjoint (DISCRETE flag here) needs its checkpoints -- dynax_workspace_bytes says how much
    assert L.dynax_workspace_bytes(_lib.OP_RC_LOSS_GRAD, 8, 20, 3, 5, 1, 0) == 0
    assert L.dynax_workspace_bytes(_lib.OP_RC_LOSS_GRAD, 8, 20, 3, 5, 1, _lib.SHARED_THETA) == 1024
    assert L.dynax_workspace_bytes(_lib.OP_RC_LOSS_GRAD, 8, 20, 3, 5, 1, _lib.DISCRETE) > 1024
    rc = L.dynax_rc_loss_grad_f32(th.data_ptr(), x0.data_ptr(), u.data_ptr(), target.data_ptr(), None, None,
                                  loss.data_ptr(), grad.data_ptr(), None, 8, 20, 3600.0, 0, None, 0, None)
    assert rc == 0
    T.cuda.synchronize()
    rc = L.dynax_rc_loss_grad_f32(th.data_ptr(), x0.data_ptr(), u.data_ptr(), target.data_ptr(), None, None,
                                  loss.data_ptr(), None, None, 8, 20, 3600.0, 0, None, 0, None)
    assert rc == 0                      # g_theta may be NULL: loss only
    T.cuda.synchronize()
    rc = L.dynax_rc_loss_grad_f32(th.data_ptr(), x0.data_ptr(), u.data_ptr(), target.data_ptr(), None, None,
                                  loss.data_ptr(), grad.data_ptr(), None, 8, 20, 3600.0, _lib.DISCRETE, None, 0, None)
    assert rc == -3 and b"workspace" in L.dynax_last_error()
    rc = L.dynax_rc_loss_grad_f32(th[0].data_ptr(), x0.data_ptr(), u.data_ptr(), target.data_ptr(), None, None,
                                  loss.data_ptr(), grad.data_ptr(), None, 8, 20, 3600.0, _lib.SHARED_THETA, None, 0, None)
    assert rc == -3 and b"workspace" in L.dynax_last_error()
    rc = L.dynax_rc_loss_grad_f32(th.data_ptr(), x0.data_ptr(), u.data_ptr(), None, None, None,
                                  loss.data_ptr(), grad.data_ptr(), None, 8, 20, 3600.0, 0, None, 0, None)
    assert rc == -1
    rc = L.dynax_rc_loss_grad_f32(th.data_ptr(), x0.data_ptr(), u.data_ptr(), target.data_ptr(), None, None,
                                  loss.data_ptr(), grad.data_ptr(), None, 8, 0, 3600.0, 0, None, 0, None)
    assert rc == -1      # T must be >= 1 for a mean over time


@pytest.mark.parametrize("N,Tn", [(300, 190), (1, 40), (129, 13)])
def test_vjp_shared_input_series_sums_g_u_in_kernel(T, ops, N, Tn):
    """A shared input series u[T,1,5] has ONE cotangent row per step = the sum over systems of the per-system rows; the
    kernel reduces it (per-CTA sums + a fixed-order reduction), no [T,N,5] array is materialised.  Against the fp64
    oracle's per-system g_u summed over systems; bit-reproducible from run to run; and through autograd."""
    rng = np.random.default_rng(46)
    th, x0, u = orc.synth_theta(N, 46), orc.synth_x0(N, 46), orc.synth_u(Tn, 1, 46)
    gx = rng.normal(size=(Tn, N, 3)).astype(np.float32)
    gy = rng.normal(size=(Tn, N, 1)).astype(np.float32)
    a = (cu(T, th), cu(T, x0), cu(T, u))
    gth, gx0, gu = ops.rc_vjp(*a, 3600.0, cu(T, gx), cu(T, gy), time_chunks=0)
    assert gu.shape == (Tn, 1, 5)
    uN = np.repeat(u, N, 1)
    rth, rx0, ru = orc.rc_vjp(th, x0, uN, 3600.0, gx, gy, np.float64)
    fth, fx0, fu = orc.rc_vjp(th, x0, uN, 3600.0, gx, gy, np.float32)
    e, e32 = rel_err(gu.cpu().numpy()[:, 0], ru.sum(1), floor=1.0), rel_err(fu.astype(np.float64).sum(1), ru.sum(1), floor=1.0)
    assert e < max(GRAD_RTOL, 1.25 * e32), (e, e32)
    assert normwise(gth.cpu().numpy(), rth).max() <= max(GRAD_RTOL, 1.25 * normwise(fth, rth).max())
    gth2, _, gu2 = ops.rc_vjp(*a, 3600.0, cu(T, gx), cu(T, gy), time_chunks=0)
    assert T.equal(gu, gu2)                                        # fixed-order reduction
    tu = cu(T, u[:, 0]).requires_grad_(True)                       # u[T,5] through autograd
    xs, ys = ops.rc_rollout(a[0], a[1], tu, 3600.0)
    ((xs * cu(T, gx)).sum() + (ys * cu(T, gy)).sum()).backward()
    assert tu.grad.shape == (Tn, 5) and T.allclose(tu.grad, gu[:, 0], rtol=1e-5, atol=1e-6 * float(gu.abs().max()))
