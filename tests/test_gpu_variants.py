"""GPU parity per KERNEL VARIANT: every rollout-kernel launch site the library instantiates (tile widths, chunk
lengths, shared-input modes, control channel, d loss/d x0, threads per system, time-chunked launches, the dense
(n,m,p) shapes) is selected by a case below -- proven through the library's launch registry
(dynax_debug_kernel_*, include/dynax_b200.h), not assumed from the dispatch code -- and its results are compared with
the fp64 oracle on a sample of systems.  The last test fails if any registered launch site was never selected in this
process, listing the missing template arguments (a new instantiation needs a new case here).

Sizes: the wide tiles need N > 148 x 128 systems; T stays short so a case costs a few hundred MB and milliseconds.
BASELINE.json config 3's own launch (65 536 candidates against ONE measured series, T=8760) is the last explicit case.
"""
import re

import numpy as np
import pytest

from oracle import dynax_oracle as orc
from tests.util import assert_grad_parity, assert_loss_parity, normwise, rel_err, GRAD_RTOL, TRAJ_RTOL

pytestmark = pytest.mark.gpu

DT = 3600.0
SM = 148                      # B200
N_SMALL = 1000                # <= 2 x 32-system CTAs per SM
N_MID = 148 * 2 * 64 - 40     # one thread per system would leave SMs idle: 2 threads per system
N_128 = 148 * 128 - 24        # one 128-system CTA per SM
N_224 = 65536                 # 293 CTAs of 224 fill the 296 slots once
N_256 = 148 * 2 * 256         # exactly one wave of 256-system CTAs
SAMPLE = 32


@pytest.fixture(scope="module")
def T():
    import torch
    return torch


@pytest.fixture(scope="module")
def ops():
    from dynax_b200 import ops
    return ops


@pytest.fixture(scope="module")
def L():
    from dynax_b200 import _lib
    return _lib.lib()


def cu(T, a):
    return T.as_tensor(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def registry(L):
    """tag -> launches so far (sites registered from two translation units share a tag: summed)."""
    out = {}
    for i in range(L.dynax_debug_kernel_count()):
        t = L.dynax_debug_kernel_tag(i).decode()
        out[t] = out.get(t, 0) + L.dynax_debug_kernel_launches(i)
    return out


def short(tag):
    f = re.search(r"dynax::(launch_\w+)\(", tag).group(1)
    w = tag[tag.index("[with ") + 6:-1].replace("dynax::", "").replace("int ", "").replace("bool ", "")
    return f"{f}<{w}>"


def launched_by(L, fn):
    """Runs fn() and returns (its result, the set of launch sites it used)."""
    before = registry(L)
    out = fn()
    after = registry(L)
    return out, {short(t) for t in after if after[t] > before.get(t, 0)}


def expect(sites, *needles):
    """One of the launch sites used contains every needle (template arguments as printed by the registry)."""
    ok = any(all(n in s for n in needles) for s in sites)
    assert ok, (needles, sorted(sites))


_DATA = {}


def rc_data(N, Tn, seed=7):
    """Synthetic inputs, cached per (N, Tn): theta, x0, u, target (target = 22 + N(0,1): large residual)."""
    key = (N, Tn, seed)
    if key not in _DATA:
        _DATA.clear()                                     # one big case at a time
        rng = np.random.default_rng(seed)
        th = (orc.RC_NOMINAL[None] * np.exp(rng.uniform(-0.2, 0.2, size=(N, 7)))).astype(np.float32)
        x0 = (np.array([20.0, 30.0, 26.0])[None] + rng.uniform(-1, 1, size=(N, 3))).astype(np.float32)
        u = rng.random((Tn, N, 5), dtype=np.float32)
        u[..., 0] = 15 + 10 * u[..., 0]
        u[..., 2] *= -5
        target = (22 + rng.normal(size=(Tn, N))).astype(np.float32)
        _DATA[key] = (th, x0, u, target)
    return _DATA[key]


def sample_idx(N, seed=3):
    return np.sort(np.random.default_rng(seed).choice(N, min(SAMPLE, N), replace=False))


# ----------------------------------------------------------------------------------------------
# tangent kernel (fused loss + gradient): widths x input modes x g_x0, threads per system, chunked
# ----------------------------------------------------------------------------------------------
MODES = ["per", "su", "stg", "su_stg", "ctrl", "ctrl_stg"]


def _loss_grad_case(T, ops, N, Tn, mode, gx0, shared_theta=False, **kw):
    th, x0, u, target = rc_data(N, Tn)
    idx = sample_idx(N)
    su, stg, ctrl = mode in ("su", "su_stg"), mode.endswith("stg"), mode.startswith("ctrl")
    u_k = u[:, :1] if (su or ctrl) else u
    tg_k = target[:, 0] if stg else target
    th_k = th[0] if shared_theta else th
    if ctrl:
        cch = np.ascontiguousarray(u[:, :, 2])
        out = ops.rc_loss_grad_ctrl(cu(T, th_k), cu(T, x0), cu(T, u_k[:, 0]), cu(T, cch), cu(T, tg_k), DT, ctrl_col=2,
                                    lb=orc.RC_LB, ub=orc.RC_UB, want_gx0=gx0)
    else:
        out = ops.rc_loss_grad(cu(T, th_k), cu(T, x0), cu(T, u_k), cu(T, tg_k), DT, orc.RC_LB, orc.RC_UB, want_gx0=gx0, **kw)
    # oracle on the sampled systems with the inputs each of them really saw
    us = (np.repeat(u_k, len(idx), axis=1) if (su or ctrl) else u[:, idx]).copy()
    if ctrl:
        us[:, :, 2] = u[:, idx, 2]
    tgs = np.repeat(target[:, :1], len(idx), axis=1) if stg else target[:, idx]
    ths = np.repeat(th[:1], len(idx), 0) if shared_theta else th[idx]
    return out, idx, (ths, x0[idx], us, tgs)


def _check_loss_grad(out, idx, ref_in, gx0):
    loss, grad, g_x0 = out
    ths, x0s, us, tgs = ref_in
    l64, g64 = orc.rc_loss_grad(ths, x0s, us, tgs, DT, orc.RC_LB, orc.RC_UB, np.float64)
    lbud, gbud = orc.rc_output_sensitivity_budget(ths, x0s, us, tgs, DT)
    assert_loss_parity(loss.cpu().numpy()[idx], l64, lbud)
    assert_grad_parity(grad.cpu().numpy()[idx], g64, gbud)
    if gx0:
        gy = (2.0 / us.shape[0]) * (orc.rc_rollout(ths, x0s, us, DT, np.float64)[1][..., 0] - tgs)
        _, gx64, _ = orc.rc_vjp(ths, x0s, us, DT, None, gy[..., None])
        assert rel_err(g_x0.cpu().numpy()[idx], gx64, floor=1.0) < 10 * GRAD_RTOL


@pytest.mark.parametrize("gx0", [False, True])
@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("N,tn", [(N_128, 128), (N_224, 224), (N_256, 256)])
def test_tangent_wide_variants(T, ops, L, monkeypatch, N, tn, mode, gx0):
    Tn = 67                                            # 8 full 8-row chunks + a ragged one
    monkeypatch.setenv("DYNAX_GRAD_ALGO", "1")
    monkeypatch.setenv("DYNAX_FS_SPLIT", "1")          # one thread per system: the wide tiles
    (out, idx, ref_in), sites = launched_by(L, lambda: _loss_grad_case(T, ops, N, Tn, mode, gx0, time_chunks=0))
    want_tn = 224 if (gx0 and tn == 256) else tn       # d loss/d x0: 11 fp64 totals per system -> 224-wide at most
    su, stg, ctrl = mode in ("su", "su_stg") or mode.startswith("ctrl"), mode.endswith("stg"), mode.startswith("ctrl")
    expect(sites, "launch_fwdsens", f"TN = {want_tn};", f"SU = {str(su).lower()};", f"STG = {str(stg).lower()};",
           f"CTRL = {str(ctrl).lower()};", f"GX0 = {str(gx0).lower()};", "SPLIT = 1;", "CHUNK = false")
    _check_loss_grad(out, idx, ref_in, gx0)


@pytest.mark.parametrize("mode", ["per", "su", "stg", "su_stg"])
@pytest.mark.parametrize("N,split,tn,force", [(N_SMALL, 2, 64, None), (8192, 2, 64, None), (N_MID, 2, 64, "2")])
def test_tangent_split_variants(T, ops, L, monkeypatch, N, split, tn, force, mode):
    """Launches that cannot fill the GPU with one thread per system: 2 threads per system (automatic up to one
    64-system CTA per SM -- 8192 systems is BASELINE.json config 3 sharded over 8 GPUs), or forced through
    DYNAX_FS_SPLIT.  T = 150: two full 64-row chunks and a ragged one."""
    Tn = 150
    monkeypatch.setenv("DYNAX_GRAD_ALGO", "1")
    if force:
        monkeypatch.setenv("DYNAX_FS_SPLIT", force)
    (out, idx, ref_in), sites = launched_by(L, lambda: _loss_grad_case(T, ops, N, Tn, mode, False, time_chunks=0))
    su, stg = mode in ("su", "su_stg"), mode.endswith("stg")
    expect(sites, "launch_fwdsens", f"TN = {tn};", f"SPLIT = {split};", f"SU = {str(su).lower()};",
           f"STG = {str(stg).lower()};")
    _check_loss_grad(out, idx, ref_in, False)


@pytest.mark.parametrize("split", [1, 2])
def test_tangent_shared_theta_all_splits(T, ops, L, monkeypatch, split):
    """ONE parameter vector for all systems (BASELINE.json config 3's reduction): gradient summed over systems."""
    N, Tn = 3000, 333
    monkeypatch.setenv("DYNAX_GRAD_ALGO", "1")
    monkeypatch.setenv("DYNAX_FS_SPLIT", str(split))
    th, x0, u, target = rc_data(N, Tn)
    (loss, grad, _), sites = launched_by(L, lambda: ops.rc_loss_grad(cu(T, th[0]), cu(T, x0), cu(T, u), cu(T, target), DT,
                                                                      orc.RC_LB, orc.RC_UB, time_chunks=0))
    expect(sites, "launch_fwdsens", f"SPLIT = {split};")
    thN = np.repeat(th[:1], N, 0)
    l64, g64 = orc.rc_loss_grad(thN, x0, u, target, DT, None, None, np.float64)
    lbud, gbud = orc.rc_output_sensitivity_budget(thN, x0, u, target, DT)
    pen, peng = orc.bound_penalty(th[:1], orc.RC_LB, orc.RC_UB)[0], orc.bound_penalty_grad(th[:1], orc.RC_LB, orc.RC_UB)[0]
    assert_loss_parity(loss.cpu().numpy()[0], l64.sum() + pen, lbud.sum())
    assert_grad_parity(grad.cpu().numpy(), g64.sum(0) + peng, gbud.sum(0))


@pytest.mark.parametrize("mode", ["per", "su", "stg", "su_stg"])
@pytest.mark.parametrize("algo", ["1", "0"])
def test_chunked_loss_grad_variants(T, ops, L, monkeypatch, mode, algo):
    """Time-parallel loss+gradient: tangent CHUNK kernels (default) and the adjoint-based scheme (DYNAX_GRAD_ALGO=0)."""
    N, Tn = 40, 1200
    monkeypatch.setenv("DYNAX_GRAD_ALGO", algo)
    (out, idx, ref_in), sites = launched_by(L, lambda: _loss_grad_case(T, ops, N, Tn, mode, False, time_chunks=9))
    su, stg = mode in ("su", "su_stg"), mode.endswith("stg")
    if algo == "1":
        expect(sites, "launch_fwdsens", "CHUNK = true", f"SU = {str(su).lower()};", f"STG = {str(stg).lower()};")
    else:
        expect(sites, "launch_adj", "RCModel", "MODE = 0", f"SHARED_U = {str(su).lower()};", f"SHARED_TGT = {str(stg).lower()};")
    expect(sites, "launch_fwd<", "RCModel", f"SHARED_U = {str(su).lower()};")      # the zero-state pass
    _check_loss_grad(out, idx, ref_in, False)


# ----------------------------------------------------------------------------------------------
# adjoint kernel, fused MSE (DYNAX_GRAD_ALGO=0) and general cotangents
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", ["per", "su", "stg", "su_stg"])
@pytest.mark.parametrize("N", [N_SMALL, N_224])
def test_adjoint_mse_variants(T, ops, L, monkeypatch, N, mode):
    Tn = 61
    monkeypatch.setenv("DYNAX_GRAD_ALGO", "0")
    (out, idx, ref_in), sites = launched_by(L, lambda: _loss_grad_case(T, ops, N, Tn, mode, True, time_chunks=0))
    su, stg = mode in ("su", "su_stg"), mode.endswith("stg")
    expect(sites, "launch_adj", "RCModel", "MODE = 0", f"SHARED_U = {str(su).lower()};", f"SHARED_TGT = {str(stg).lower()};")
    if mode == "per":   # wave quantisation picks 12-step segments x3 CTAs/SM or 16-step x2 (65536: the latter)
        expect(sites, "launch_adj", "KC = 16;" if N == N_224 else "KC = 12;", "MODE = 0")
    _check_loss_grad(out, idx, ref_in, True)


@pytest.mark.parametrize("N,gx,need_u,su,kc", [(N_SMALL, True, True, False, 16), (N_224, True, True, False, 12),
                                               (N_224, False, False, False, 8), (N_SMALL, True, False, True, 8)])
def test_adjoint_vjp_variants(T, ops, L, N, gx, need_u, su, kc):
    Tn = 45
    th, x0, u, _ = rc_data(N, Tn)
    rng = np.random.default_rng(5)
    g_x = rng.normal(size=(Tn, N, 3)).astype(np.float32) if gx else None
    g_y = rng.normal(size=(Tn, N, 1)).astype(np.float32)
    u_k = u[:, :1] if su else u
    (gth, gx0, gu), sites = launched_by(L, lambda: ops.rc_vjp(cu(T, th), cu(T, x0), cu(T, u_k), DT,
                                                              None if g_x is None else cu(T, g_x), cu(T, g_y),
                                                              need_u=need_u, time_chunks=0))
    expect(sites, "launch_adj", "RCModel", f"KC = {kc};", "MODE = 1", f"SHARED_U = {str(su).lower()};")
    idx = sample_idx(N)
    us = np.repeat(u_k, len(idx), 1) if su else u[:, idx]
    gxs = None if g_x is None else g_x[:, idx]
    rth, rx0, ru = orc.rc_vjp(th[idx], x0[idx], us, DT, gxs, g_y[:, idx], np.float64)
    fth, fx0, fu = orc.rc_vjp(th[idx], x0[idx], us, DT, gxs, g_y[:, idx], np.float32)
    no, nr = normwise(gth.cpu().numpy()[idx], rth), normwise(fth, rth)
    assert no.max() <= max(GRAD_RTOL, 1.25 * nr.max()), (no.max(), nr.max())
    assert rel_err(gx0.cpu().numpy()[idx], rx0, floor=1.0) < max(GRAD_RTOL, 1.25 * rel_err(fx0, rx0, floor=1.0))
    if need_u:
        assert rel_err(gu.cpu().numpy()[:, idx], ru, floor=1.0) < max(GRAD_RTOL, 1.25 * rel_err(fu, ru, floor=1.0))


def test_chunked_vjp_variants(T, ops, L):
    N, Tn = 24, 1500
    th, x0, u, _ = rc_data(N, Tn)
    rng = np.random.default_rng(5)
    g_x = rng.normal(size=(Tn, N, 3)).astype(np.float32)
    g_y = rng.normal(size=(Tn, N, 1)).astype(np.float32)
    (gth, gx0, gu), sites = launched_by(L, lambda: ops.rc_vjp(cu(T, th), cu(T, x0), cu(T, u), DT, cu(T, g_x), cu(T, g_y),
                                                              time_chunks=8))
    expect(sites, "launch_adj", "RCModel", "KC = 8;", "MODE = 1")
    expect(sites, "launch_fwd<", "RCModel", "KF = 8;", "MINB = 2")
    rth, rx0, ru = orc.rc_vjp(th, x0, u, DT, g_x, g_y, np.float64)
    fth, fx0, fu = orc.rc_vjp(th, x0, u, DT, g_x, g_y, np.float32)
    assert normwise(gth.cpu().numpy(), rth).max() <= max(GRAD_RTOL, 1.25 * normwise(fth, rth).max())
    assert rel_err(gu.cpu().numpy(), ru, floor=1.0) < max(GRAD_RTOL, 1.25 * rel_err(fu, ru, floor=1.0))


# ----------------------------------------------------------------------------------------------
# forward kernel
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,su,tn,kf,stages", [(N_SMALL, False, 128, 16, 2), (N_128, False, 128, 16, 4), (148 * 200, False, 128, 8, 2),
                                               (N_256, False, 256, 4, 4), (N_SMALL, True, 128, 16, 4), (N_224, True, 128, 4, 4)])
def test_forward_variants(T, ops, L, N, su, tn, kf, stages):
    Tn = 50
    th, x0, u, _ = rc_data(N, Tn)
    u_k = u[:, :1] if su else u
    (xs, ys, xf), sites = launched_by(L, lambda: ops.rc_forward(cu(T, th), cu(T, x0), cu(T, u_k), DT, emit_final=True,
                                                                time_chunks=0))
    expect(sites, "launch_fwd<", "RCModel", f"TN = {tn};", f"KF = {kf};", f"S = {stages};", f"SHARED_U = {str(su).lower()};")
    idx = sample_idx(N)
    us = np.repeat(u_k, len(idx), 1) if su else u[:, idx]
    x64, y64 = orc.rc_rollout(th[idx], x0[idx], us, DT, np.float64)
    assert rel_err(xs.cpu().numpy()[:, idx], x64) < TRAJ_RTOL and rel_err(ys.cpu().numpy()[:, idx], y64) < TRAJ_RTOL
    assert T.equal(xf, xs[-1])


@pytest.mark.parametrize("su", [False, True])
def test_chunked_forward_variants(T, ops, L, su):
    N, Tn = 30, 2000
    th, x0, u, _ = rc_data(N, Tn)
    u_k = u[:, :1] if su else u
    (xs, ys, _), sites = launched_by(L, lambda: ops.rc_forward(cu(T, th), cu(T, x0), cu(T, u_k), DT, time_chunks=10))
    expect(sites, "launch_fwd<", "RCModel", "TN = 128;", f"SHARED_U = {str(su).lower()};")
    x64, y64 = orc.rc_rollout(th, x0, np.repeat(u_k, N, 1) if su else u, DT, np.float64)
    assert rel_err(xs.cpu().numpy(), x64) < TRAJ_RTOL and rel_err(ys.cpu().numpy(), y64) < TRAJ_RTOL


# ----------------------------------------------------------------------------------------------
# rows TMA cannot move (N % 4 != 0) on launches of more than one wave: the shifted-row instantiations
# (rows staged at their global offset modulo 16 bytes, one bulk copy per row interior)
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N", [148 * 2 * 256 + 1, 75777, 75778, 75779])   # N % 4 = 1, 1, 2, 3
def test_shifted_row_variants(T, ops, L, monkeypatch, N):
    Tn = 43
    th, x0, u, tg = rc_data(N, Tn)
    idx = sample_idx(N)
    thd, x0d, ud, tgd = cu(T, th), cu(T, x0), cu(T, u), cu(T, tg)
    (xs, ys, _), sites = launched_by(L, lambda: ops.rc_forward(thd, x0d, ud, DT, time_chunks=0))
    expect(sites, "launch_fwd<", "RCModel", "TN = 256;", "SHIFT = true")
    x64, y64 = orc.rc_rollout(th[idx], x0[idx], u[:, idx], DT, np.float64)
    assert rel_err(xs.cpu().numpy()[:, idx], x64) < TRAJ_RTOL and rel_err(ys.cpu().numpy()[:, idx], y64) < TRAJ_RTOL
    lb, ub = orc.RC_LB, orc.RC_UB
    outs = {}
    for algo in ("1", "0"):                        # tangent kernel, adjoint kernel
        monkeypatch.setenv("DYNAX_GRAD_ALGO", algo)
        (out, sites) = launched_by(L, lambda: ops.rc_loss_grad(thd, x0d, ud, tgd, DT, lb, ub, time_chunks=0))
        expect(sites, "launch_fwdsens" if algo == "1" else "launch_adj", "SHIFT = true")
        outs[algo] = out
        _check_loss_grad(out, idx, (th[idx], x0[idx], u[:, idx], tg[:, idx]), False)
    monkeypatch.delenv("DYNAX_GRAD_ALGO")
    # the unshifted paths (4-byte copies) give the same numbers bit for bit: same arithmetic on the same inputs
    monkeypatch.setenv("DYNAX_NO_SHIFT", "1")
    xs2, ys2, _ = ops.rc_forward(thd, x0d, ud, DT, time_chunks=0)
    l2, g2, _ = ops.rc_loss_grad(thd, x0d, ud, tgd, DT, lb, ub, time_chunks=0)
    monkeypatch.delenv("DYNAX_NO_SHIFT")
    assert T.equal(xs, xs2) and T.equal(ys, ys2)
    assert T.equal(outs["1"][0], l2) and T.equal(outs["1"][1], g2)
    rng = np.random.default_rng(6)
    g_x = rng.normal(size=(Tn, N, 3)).astype(np.float32)
    g_y = rng.normal(size=(Tn, N, 1)).astype(np.float32)
    (gth, gx0, gu), sites = launched_by(L, lambda: ops.rc_vjp(thd, x0d, ud, DT, cu(T, g_x), cu(T, g_y), time_chunks=0))
    expect(sites, "launch_adj", "MODE = 1", "SHIFT = true")
    rth, rx0, ru = orc.rc_vjp(th[idx], x0[idx], u[:, idx], DT, g_x[:, idx], g_y[:, idx], np.float64)
    fth, fx0, fu = orc.rc_vjp(th[idx], x0[idx], u[:, idx], DT, g_x[:, idx], g_y[:, idx], np.float32)
    assert normwise(gth.cpu().numpy()[idx], rth).max() <= max(GRAD_RTOL, 1.25 * normwise(fth, rth).max())
    assert rel_err(gu.cpu().numpy()[:, idx], ru, floor=1.0) < max(GRAD_RTOL, 1.25 * rel_err(fu, ru, floor=1.0))


# ----------------------------------------------------------------------------------------------
# dense LTI shapes: forward (shared u / small / wide), VJP (shared u / narrow / wide tile), chunked
# ----------------------------------------------------------------------------------------------
DIMS = [(1, 1, 1), (2, 1, 1), (2, 2, 1), (2, 2, 2), (3, 1, 1), (3, 3, 1), (3, 3, 3), (3, 5, 1), (4, 1, 1), (4, 2, 2),
        (4, 3, 4), (4, 4, 4)]


def lti_data(N, Tn, n, m, p, seed=11):
    rng = np.random.default_rng(seed)
    A = (-np.eye(n)[None] + 0.3 * rng.standard_normal((N, n, n), dtype=np.float32)).astype(np.float32)
    B = rng.standard_normal((N, n, m), dtype=np.float32)
    C = rng.standard_normal((N, p, n), dtype=np.float32)
    D = rng.standard_normal((N, p, m), dtype=np.float32)
    x0 = rng.standard_normal((N, n), dtype=np.float32)
    u = rng.standard_normal((Tn, N, m), dtype=np.float32)
    return A, B, C, D, x0, u


@pytest.mark.parametrize("n,m,p", DIMS)
@pytest.mark.parametrize("N,su", [(300, False), (300, True), (148 * 256 + 300, True), (148 * 128 + 4000, False)])
def test_lti_forward_variants(T, ops, L, n, m, p, N, su):
    Tn, dt = 40, 0.05
    A, B, C, D, x0, u = lti_data(N, Tn, n, m, p)
    u_k = u[:, :1] if su else u
    (xs, ys, _), sites = launched_by(L, lambda: ops.lti_forward(*(cu(T, v) for v in (A, B, C, D, x0, u_k)), dt, time_chunks=0))
    expect(sites, "launch_fwd<", f"DenseModel<{n}, {m}, {p}>", f"SHARED_U = {str(su).lower()};",
           "TN = 256;" if (N > 148 * 128 and not su) else "TN = 128;",
           *([("KF = 16;" if N <= 148 * 256 else "KF = 4;")] if su else []))
    idx = sample_idx(N)
    us = np.repeat(u_k, len(idx), 1) if su else u[:, idx]
    xr, yr = orc.lti_rollout(A[idx], B[idx], C[idx], D[idx], x0[idx], us, dt, np.float64)
    assert rel_err(xs.cpu().numpy()[:, idx], xr, floor=1.0) < TRAJ_RTOL
    assert rel_err(ys.cpu().numpy()[:, idx], yr, floor=1.0) < TRAJ_RTOL


@pytest.mark.parametrize("n,m,p", DIMS)
@pytest.mark.parametrize("N,su", [(300, False), (300, True), (148 * 128 + 4000, False)])
def test_lti_vjp_variants(T, ops, L, n, m, p, N, su):
    Tn, dt = 33, 0.05
    A, B, C, D, x0, u = lti_data(N, Tn, n, m, p, 12)
    rng = np.random.default_rng(13)
    gx = rng.standard_normal((Tn, N, n), dtype=np.float32)
    gy = rng.standard_normal((Tn, N, p), dtype=np.float32)
    u_k = u[:, :1] if su else u
    out, sites = launched_by(L, lambda: ops.lti_vjp(*(cu(T, v) for v in (A, B, C, D, x0, u_k)), dt, cu(T, gx), cu(T, gy),
                                                    need_u=not su, time_chunks=0))
    expect(sites, "launch_adj", f"DenseModel<{n}, {m}, {p}>", "MODE = 1", f"SHARED_U = {str(su).lower()};")
    idx = sample_idx(N)
    us = np.repeat(u_k, len(idx), 1) if su else u[:, idx]
    r = orc.lti_vjp(A[idx], B[idx], C[idx], D[idx], x0[idx], us, dt, gx[:, idx], gy[:, idx], np.float64)
    for got, key in zip(out[:5], ("gA", "gB", "gC", "gD", "gx0")):
        assert rel_err(got.cpu().numpy()[idx], r[key], floor=1.0) < GRAD_RTOL, key
    if not su:
        assert rel_err(out[5].cpu().numpy()[:, idx], r["gu"], floor=1.0) < GRAD_RTOL


@pytest.mark.parametrize("n,m,p", DIMS)
def test_lti_chunked_variants(T, ops, L, n, m, p):
    N, Tn, dt = 20, 900, 0.02
    A, B, C, D, x0, u = lti_data(N, Tn, n, m, p, 14)
    A *= 0.5
    rng = np.random.default_rng(15)
    gx = rng.standard_normal((Tn, N, n), dtype=np.float32)
    gy = rng.standard_normal((Tn, N, p), dtype=np.float32)
    d = [cu(T, v) for v in (A, B, C, D, x0, u)]
    (xs, ys, _), s1 = launched_by(L, lambda: ops.lti_forward(*d, dt, time_chunks=6))
    (_, _, _), s1s = launched_by(L, lambda: ops.lti_forward(*d[:5], d[5][:, :1].contiguous(), dt, time_chunks=6))
    out, s2 = launched_by(L, lambda: ops.lti_vjp(*d, dt, cu(T, gx), cu(T, gy), time_chunks=6))
    expect(s1, "launch_fwd<", f"DenseModel<{n}, {m}, {p}>", "KF = 8;", "SHARED_U = false")
    expect(s1s, "launch_fwd<", f"DenseModel<{n}, {m}, {p}>", "SHARED_U = true")
    expect(s2, "launch_adj", f"DenseModel<{n}, {m}, {p}>", "MODE = 1")
    xr, yr = orc.lti_rollout(A, B, C, D, x0, u, dt, np.float64)
    assert rel_err(xs.cpu().numpy(), xr, floor=1.0) < TRAJ_RTOL and rel_err(ys.cpu().numpy(), yr, floor=1.0) < TRAJ_RTOL
    r = orc.lti_vjp(A, B, C, D, x0, u, dt, gx, gy, np.float64)
    for got, key in zip(out, ("gA", "gB", "gC", "gD", "gx0", "gu")):
        assert rel_err(got.cpu().numpy(), r[key], floor=1.0) < GRAD_RTOL, key


def test_node_cuda_core_forward_variants(T, ops, L, monkeypatch):
    """The CUDA-core MLP rollout (DYNAX_NODE_IMPL=0, and every shared-input launch) on the shared forward pipeline."""
    N, Tn, dt = 200, 40, 0.25
    W = orc.synth_mlp(21)
    rng = np.random.default_rng(22)
    x0 = rng.normal(size=(N, 3)).astype(np.float32)
    u = rng.normal(size=(Tn, N, 5)).astype(np.float32)
    monkeypatch.setenv("DYNAX_NODE_IMPL", "0")
    for su in (False, True):
        u_k = u[:, :1] if su else u
        (xs, ys, _), sites = launched_by(L, lambda: ops.node_rk4_forward(*(cu(T, w) for w in W), cu(T, x0), cu(T, u_k), dt))
        expect(sites, "launch_fwd<", "MlpModel", f"SHARED_U = {str(su).lower()};")
        xr, _ = orc.node_rk4_rollout(*W, x0, np.repeat(u_k, N, 1) if su else u, dt, np.float64)
        assert rel_err(xs.cpu().numpy(), xr, floor=1.0) < TRAJ_RTOL


# ----------------------------------------------------------------------------------------------
# BASELINE.json config 3 as it is launched: 65 536 candidates against ONE measured series, T = 8760
# ----------------------------------------------------------------------------------------------
def test_config3_full_size(T, ops, L, monkeypatch):
    N, Tn = 65536, 8760
    monkeypatch.setenv("DYNAX_GRAD_ALGO", "1")
    monkeypatch.delenv("DYNAX_FS_SPLIT", raising=False)
    th = orc.synth_theta(N, 2345)
    x0 = orc.synth_x0(N, 2345)
    u1 = orc.synth_u(Tn, 1, 2345)
    hidden = orc.RC_NOMINAL[None].astype(np.float32)
    _, y = orc.rc_rollout(hidden, np.array([[20.0, 30.0, 26.0]], np.float32), u1, DT, np.float64)
    target = (y[:, 0, 0] + 0.1 * np.random.default_rng(2345).normal(size=Tn)).astype(np.float32)
    (loss, grad, _), sites = launched_by(L, lambda: ops.rc_loss_grad(cu(T, th), cu(T, x0), cu(T, u1), cu(T, target), DT,
                                                                      orc.RC_LB, orc.RC_UB))
    expect(sites, "launch_fwdsens", "TN = 224;", "SU = true;", "STG = true;", "SPLIT = 1;")
    idx = sample_idx(N)
    us, tgs = np.repeat(u1, len(idx), 1), np.repeat(target[:, None], len(idx), 1)
    l64, g64 = orc.rc_loss_grad(th[idx], x0[idx], us, tgs, DT, orc.RC_LB, orc.RC_UB, np.float64)
    lbud, gbud = orc.rc_output_sensitivity_budget(th[idx], x0[idx], us, tgs, DT)
    assert_loss_parity(loss.cpu().numpy()[idx], l64, lbud)
    assert_grad_parity(grad.cpu().numpy()[idx], g64, gbud)
    assert bool(T.isfinite(loss).all()) and bool(T.isfinite(grad).all())


# ----------------------------------------------------------------------------------------------
def test_zz_every_instantiated_variant_was_selected(L):
    """Runs last in this file: every launch site registered by the library has been used by a case above."""
    reg = registry(L)
    missing = sorted(short(t) for t, n in reg.items() if n == 0)
    assert not missing, f"{len(missing)} of {len(reg)} launch sites never selected:\n" + "\n".join(missing)
