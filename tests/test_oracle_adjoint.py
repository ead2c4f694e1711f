"""Pin the oracle's hand adjoint (A5) against torch autograd over the restated scan (fp64).

The reference holds no gradient test, so this is what anchors the
gradient oracle: an independent reverse-mode derivation of the same loop.
"""
import numpy as np
import torch

from oracle import dynax_oracle as orc


def _torch_rc_abcd(th):
    Cai, Cwe, Cwi, Re, Ri, Rw, Rg = th.unbind(-1)
    z = torch.zeros_like(Cai)
    A = torch.stack([torch.stack([-1 / Cai * (1 / Rg + 1 / Ri), z, 1 / (Cai * Ri)], -1),
                     torch.stack([z, -1 / Cwe * (1 / Re + 1 / Rw), 1 / (Cwe * Rw)], -1),
                     torch.stack([1 / (Cwi * Ri), 1 / (Cwi * Rw), -1 / Cwi * (1 / Rw + 1 / Ri)], -1)], -2)
    B = torch.stack([torch.stack([1 / (Cai * Rg), 1 / Cai, 1 / Cai, z, z], -1),
                     torch.stack([1 / (Cwe * Re), z, z, 1 / Cwe, z], -1),
                     torch.stack([z, z, z, z, 1 / Cwi], -1)], -2)
    return A, B


def _torch_rollout(A, B, C, D, x0, u, dt):
    x = x0
    xs, ys = [], []
    for k in range(u.shape[0]):
        rhs = torch.einsum("nrc,nc->nr", A, x) + torch.einsum("nrc,nc->nr", B, u[k])
        y = torch.einsum("nrc,nc->nr", C, x) + torch.einsum("nrc,nc->nr", D, u[k])
        x = x + dt * rhs
        xs.append(x)
        ys.append(y)
    return torch.stack(xs), torch.stack(ys)


def test_rc_loss_grad_matches_autograd():
    T, N, dt = 300, 3, 3600.0
    th = orc.synth_theta(N, 11).astype(np.float64)
    th[0, 0] = 4.0e4            # above ub -> exercises the penalty branch
    th[1, 4] = 0.4              # below lb
    x0 = orc.synth_x0(N, 11).astype(np.float64)
    u = orc.synth_u(T, N, 11).astype(np.float64)
    target = 22.0 + np.random.default_rng(0).normal(size=(T, N))
    loss, g = orc.rc_loss_grad(th, x0, u, target, dt, orc.RC_LB, orc.RC_UB, np.float64)

    tth = torch.tensor(th, requires_grad=True)
    A, B = _torch_rc_abcd(tth)
    C = torch.zeros(N, 1, 3, dtype=torch.float64); C[:, 0, 0] = 1
    D = torch.zeros(N, 1, 5, dtype=torch.float64)
    _, ys = _torch_rollout(A, B, C, D, torch.tensor(x0), torch.tensor(u), dt)
    lb, ub = torch.tensor(orc.RC_LB), torch.tensor(orc.RC_UB)
    tl = ((ys[..., 0] - torch.tensor(target)) ** 2).mean(0) \
        + (torch.relu(tth - ub) / ub).sum(-1) + (torch.relu(lb - tth) / ub).sum(-1)
    tl.sum().backward()
    assert np.allclose(loss, tl.detach().numpy(), rtol=1e-12)
    assert np.allclose(g, tth.grad.numpy(), rtol=1e-9, atol=1e-18)


def test_general_lti_vjp_matches_autograd():
    rng = np.random.default_rng(3)
    T, N, n, m, p, dt = 120, 2, 4, 3, 4, 0.05
    A = -np.eye(n)[None] + 0.3 * rng.normal(size=(N, n, n))
    B = rng.normal(size=(N, n, m)); C = rng.normal(size=(N, p, n)); D = rng.normal(size=(N, p, m))
    x0 = rng.normal(size=(N, n)); u = rng.normal(size=(T, N, m))
    gx = rng.normal(size=(T, N, n)); gy = rng.normal(size=(T, N, p))
    for discrete in (False, True):
        r = orc.lti_vjp(A, B, C, D, x0, u, dt, gx, gy, np.float64, discrete=discrete)
        t = {k: torch.tensor(v, requires_grad=True) for k, v in dict(A=A, B=B, C=C, D=D, x0=x0, u=u).items()}
        x = t["x0"]; L = 0
        for k in range(T):
            rhs = torch.einsum("nrc,nc->nr", t["A"], x) + torch.einsum("nrc,nc->nr", t["B"], t["u"][k])
            y = torch.einsum("nrc,nc->nr", t["C"], x) + torch.einsum("nrc,nc->nr", t["D"], t["u"][k])
            x = rhs if discrete else x + dt * rhs
            L = L + (x * torch.tensor(gx[k])).sum() + (y * torch.tensor(gy[k])).sum()
        L.backward()
        for k_, name in (("gA", "A"), ("gB", "B"), ("gC", "C"), ("gD", "D"), ("gx0", "x0"), ("gu", "u")):
            ref = t[name].grad.numpy()
            assert np.allclose(r[k_], ref, rtol=1e-9, atol=1e-9 * np.abs(ref).max()), (k_, discrete)
        xs, ys = orc.lti_rollout(A, B, C, D, x0, u, dt, np.float64, discrete=discrete)
        assert np.allclose(xs[-1], x.detach().numpy(), rtol=1e-12)


def test_finite_difference_spot_check():
    T, N, dt = 200, 1, 3600.0
    th = orc.synth_theta(N, 5).astype(np.float64)
    x0 = orc.synth_x0(N, 5).astype(np.float64)
    u = orc.synth_u(T, N, 5).astype(np.float64)
    target = 21.0 + np.zeros((T, N))
    _, g = orc.rc_loss_grad(th, x0, u, target, dt, None, None, np.float64)
    for j in range(7):
        h = 1e-6 * th[0, j]
        tp, tm = th.copy(), th.copy()
        tp[0, j] += h; tm[0, j] -= h
        fd = (orc.rc_loss(tp, x0, u, target, dt, dtype=np.float64) - orc.rc_loss(tm, x0, u, target, dt, dtype=np.float64)) / (2 * h)
        assert abs(fd[0] - g[0, j]) <= 1e-5 * abs(g[0, j]) + 1e-16, (j, fd, g[0, j])
