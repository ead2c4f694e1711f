"""GPU: parity at BASELINE.json's full horizon (T=8760) and full-machine tiles, through size-independent
properties plus oracle spot checks on randomly drawn systems of the big launch."""
import numpy as np
import pytest

from oracle import c_oracle, dynax_oracle as orc
from tests.util import TRAJ_RTOL, assert_grad_parity, assert_loss_parity, rel_err

pytestmark = pytest.mark.gpu
T_YEAR = 8760


def _inputs(torch, N, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    nom = torch.tensor(orc.RC_NOMINAL, dtype=torch.float32, device="cuda")
    th = nom[None] * torch.exp(torch.empty((N, 7), device="cuda").uniform_(-0.2, 0.2, generator=g))
    x0 = torch.tensor([[20.0, 30.0, 26.0]], device="cuda") + torch.empty((N, 3), device="cuda").uniform_(-1, 1, generator=g)
    u = torch.empty((T_YEAR, N, 5), device="cuda").uniform_(0, 1, generator=g)
    u[..., 0] = 15 + 10 * u[..., 0]
    u[..., 2] = -5 * u[..., 2]
    tg = 21 + 2 * torch.empty((T_YEAR, N), device="cuda").uniform_(-1, 1, generator=g)
    return th, x0, u, tg


def test_year_long_forward_one_wave():
    """One full wave of the forward kernel over a year: two half-year launches chained through x_final
    reproduce the single launch bit for bit; 48 random systems match the C oracle to 1e-5."""
    import torch
    from dynax_b200 import ops
    N = 148 * 2 * 256
    th, x0, u, _ = _inputs(torch, N, 161)
    xs, ys, xf = ops.rc_forward(th, x0, u, 3600.0, emit_final=True)
    h = T_YEAR // 2
    _, ya, xfa = ops.rc_forward(th, x0, u[:h], 3600.0, emit_x=False, emit_final=True)
    _, yb, xfb = ops.rc_forward(th, xfa, u[h:], 3600.0, emit_x=False, emit_final=True)
    assert torch.equal(torch.cat([ya, yb]), ys) and torch.equal(xfb, xf) and torch.equal(xf, xs[-1])
    idx = np.random.default_rng(0).choice(N, 48, replace=False)
    xr, yr, _ = c_oracle.rc_forward(th[idx].cpu().numpy(), x0[idx].cpu().numpy(), u[:, idx].cpu().numpy(), 3600.0)
    assert rel_err(xs[:, idx].cpu().numpy(), xr) < TRAJ_RTOL and rel_err(ys[:, idx].cpu().numpy(), yr) < TRAJ_RTOL
    assert bool(torch.isfinite(xs).all())


@pytest.mark.parametrize("algo,N", [("1", 148 * 2 * 256), ("0", 148 * 3 * 128)])
def test_year_long_loss_grad_one_wave(monkeypatch, algo, N):
    """One full wave of the tangent (DYNAX_GRAD_ALGO=1) / adjoint (=0) kernel over a year: permuting the systems permutes the results bit
    for bit (no cross-talk between threads / tiles); a launch split in two halves of the system axis gives
    the same bits; 32 random systems match the fp64 oracle within the conditioning-aware budget."""
    import torch
    from dynax_b200 import ops
    monkeypatch.setenv("DYNAX_GRAD_ALGO", algo)
    th, x0, u, tg = _inputs(torch, N, 162)
    loss, grad, _ = ops.rc_loss_grad(th, x0, u, tg, 3600.0)
    perm = torch.randperm(N, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    lp, gp, _ = ops.rc_loss_grad(th[perm], x0[perm], u[:, perm].contiguous(), tg[:, perm].contiguous(), 3600.0)
    assert torch.equal(lp, loss[perm]) and torch.equal(gp, grad[perm])
    h = N // 2
    l1, g1, _ = ops.rc_loss_grad(th[:h], x0[:h], u[:, :h].contiguous(), tg[:, :h].contiguous(), 3600.0)
    assert torch.equal(l1, loss[:h]) and torch.equal(g1, grad[:h])
    idx = np.random.default_rng(1).choice(N, 32, replace=False)
    a = [t[idx].cpu().numpy() for t in (th, x0)] + [u[:, idx].cpu().numpy(), tg[:, idx].cpu().numpy()]
    l64, g64 = orc.rc_loss_grad(*a, 3600.0)
    lbud, gbud = orc.rc_output_sensitivity_budget(*a, 3600.0)
    assert_loss_parity(loss[idx].cpu().numpy(), l64, lbud)
    assert_grad_parity(grad[idx].cpu().numpy(), g64, gbud)
    assert bool(torch.isfinite(grad).all())
