"""GPU parity: MLP-dynamics RK4 rollout (config 5, row A7) vs the restated oracle.
PARITY UNPINNED BY THE REFERENCE (it has no runnable neural ODE): the oracle is our restatement."""
import numpy as np
import pytest

from oracle import dynax_oracle as orc
from tests.util import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def T():
    import torch
    return torch


def cu(T, a):
    return T.as_tensor(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def inputs(N, Tn, seed):
    rng = np.random.default_rng(seed)
    x0 = rng.normal(size=(N, 3)).astype(np.float32)
    u = rng.normal(size=(Tn, N, 5)).astype(np.float32)        # normalised channels (cf. examples/4-pinn.py:63-69)
    return x0, u


@pytest.mark.parametrize("impl", ["1", "0"])        # 1: tcgen05/TMEM kernel (default), 0: CUDA-core kernel
@pytest.mark.parametrize("N,Tn,euler", [(300, 64, False), (129, 33, True), (4, 7, False), (1024, 256, False)])
def test_matches_oracle(T, monkeypatch, N, Tn, euler, impl):
    from dynax_b200 import ops
    monkeypatch.setenv("DYNAX_NODE_IMPL", impl)
    W = orc.synth_mlp(101)
    x0, u = inputs(N, Tn, 102)
    xs, ys, xf = ops.node_rk4_forward(*(cu(T, w) for w in W), cu(T, x0), cu(T, u), 0.25, emit_final=True, euler=euler)
    xr, yr = orc.node_rk4_rollout(*W, x0, u, 0.25, np.float64, euler=euler)
    assert rel_err(xs.cpu().numpy(), xr, floor=1.0) < 1e-5
    assert rel_err(ys.cpu().numpy(), yr, floor=1.0) < 1e-5
    assert np.array_equal(xf.cpu().numpy(), xs[-1].cpu().numpy())


@pytest.mark.parametrize("hidden", [32, 64, 128])
@pytest.mark.parametrize("N,Tn,euler,scale,shared", [(300, 40, False, 0.1, False), (131, 19, True, 1.0, False),
                                                      (517, 24, False, 1.0, True), (5, 9, False, 3.0, False),
                                                      (260, 700, False, 0.1, False)])   # (long horizon: barrier phases)
def test_hidden_widths_match_oracle(T, hidden, N, Tn, euler, scale, shared):
    """The tcgen05 kernel is instantiated for hidden widths 32, 64 and 128 (TMEM 128 / 256 / 512 columns); per-system
    and shared input series; O(1)..O(3) weights drive tanh into saturation (the quad reciprocal's clamp)."""
    from dynax_b200 import ops
    W = orc.synth_mlp(300 + hidden, hidden=hidden, scale=scale)
    x0, u = inputs(N, Tn, 301 + hidden)
    ud = cu(T, u[:, 0]) if shared else cu(T, u)
    xs, ys, xf = ops.node_rk4_forward(*(cu(T, w) for w in W), cu(T, x0), ud, 0.1, emit_final=True, euler=euler)
    xr, yr = orc.node_rk4_rollout(*W, x0, u[:, :1] if shared else u, 0.1, np.float64, euler=euler)
    assert rel_err(xs.cpu().numpy(), xr, floor=1.0) < 1e-5
    assert rel_err(ys.cpu().numpy(), yr, floor=1.0) < 1e-5
    assert np.array_equal(xf.cpu().numpy(), xs[-1].cpu().numpy())


def test_tanh_saturation_is_exact(T):
    """Pre-activations far beyond the clamp of the shared reciprocal (|x| ~ 1e3) give tanh = +-1 exactly, never NaN."""
    from dynax_b200 import ops
    W = [w.copy() for w in orc.synth_mlp(311, scale=1.0)]
    W[0] *= 300.0                                              # layer-1 pre-activations of order 1e3, both signs
    x0, u = inputs(256, 6, 312)
    xs, _, _ = ops.node_rk4_forward(*(cu(T, w) for w in W), cu(T, x0), cu(T, u), 0.05)
    xr, _ = orc.node_rk4_rollout(*W, x0, u, 0.05, np.float64)
    assert np.isfinite(xs.cpu().numpy()).all()
    assert rel_err(xs.cpu().numpy(), xr, floor=1.0) < 1e-5


def test_tensor_core_kernel_saturating_tanh_and_chaining(T):
    """tcgen05 kernel with O(1) weights (saturating tanh), ragged N, and T steps == T1 + (T-T1) steps."""
    from dynax_b200 import ops
    W = [cu(T, w) for w in orc.synth_mlp(107, scale=1.0)]
    x0, u = inputs(1000, 48, 108)
    xs, ys, xf = ops.node_rk4_forward(*W, cu(T, x0), cu(T, u), 0.1, emit_final=True)
    xr, yr = orc.node_rk4_rollout(*[w.cpu().numpy() for w in W], x0, u, 0.1, np.float64)
    assert rel_err(xs.cpu().numpy(), xr, floor=1.0) < 1e-5 and rel_err(ys.cpu().numpy(), yr, floor=1.0) < 1e-5
    xa, _, xfa = ops.node_rk4_forward(*W, cu(T, x0), cu(T, u[:20]), 0.1, emit_final=True)
    xb, _, xfb = ops.node_rk4_forward(*W, xfa, cu(T, u[20:]), 0.1, emit_final=True)
    assert T.equal(T.cat([xa, xb]), xs) and T.equal(xfb, xf)
    x3, _, _ = ops.node_rk4_forward(*W, cu(T, x0[:999]), cu(T, u[:, :999]), 0.1)       # N % 4 != 0: no-TMA producer
    assert T.equal(x3, xs[:, :999])


def test_strong_nonlinearity_and_shared_u(T):
    from dynax_b200 import ops
    W = orc.synth_mlp(103, scale=1.0)                          # saturating tanh
    x0, u = inputs(200, 40, 104)
    xs, _, _ = ops.node_rk4_forward(*(cu(T, w) for w in W), cu(T, x0), cu(T, u[:, 0]), 0.1)
    xr, _ = orc.node_rk4_rollout(*W, x0, u[:, :1], 0.1, np.float64)
    assert rel_err(xs.cpu().numpy(), xr, floor=1.0) < 1e-5


def test_unsupported_dims(T):
    from dynax_b200 import ops
    from dynax_b200._lib import DynaxError
    W = orc.synth_mlp(105, hidden=48)                          # instantiated: 32, 64, 128
    x0, u = inputs(8, 4, 106)
    with pytest.raises(DynaxError) as e:
        ops.node_rk4_forward(*(cu(T, w) for w in W), cu(T, x0), cu(T, u), 0.1)
    assert e.value.code == -2
    xs = T.zeros((4, 8, 3), device="cuda")
    with pytest.raises(DynaxError) as e:                       # the gradient kernels: same widths
        ops.node_rk4_vjp(*(cu(T, w) for w in W), cu(T, x0), cu(T, u), xs, 0.1, T.ones_like(xs), None)
    assert e.value.code == -2


def _torch_node_rollout(T, W, x0, u, dt, euler):
    W1, b1, W2, b2, W3, b3 = W
    f = lambda x, uk: T.tanh(T.tanh(T.cat([x, uk], -1) @ W1.T + b1) @ W2.T + b2) @ W3.T + b3
    x = x0; xs, ys = [], []
    for k in range(u.shape[0]):
        ys.append(x[:, :1])
        k1 = f(x, u[k])
        if euler:
            x = x + dt * k1
        else:
            k2 = f(x + 0.5 * dt * k1, u[k]); k3 = f(x + 0.5 * dt * k2, u[k]); k4 = f(x + dt * k3, u[k])
            x = x + dt / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
        xs.append(x)
    return T.stack(xs), T.stack(ys)


@pytest.mark.parametrize("impl,hidden", [("tc", 64), ("cuda_core", 64), ("cuda_core", 32), ("wide", 128)])
@pytest.mark.parametrize("N,Tn,euler,scale", [(200, 24, False, 0.5), (130, 17, True, 1.0), (3, 5, False, 0.3), (256, 40, False, 0.1)])
def test_vjp_matches_torch_autograd_fp64(T, monkeypatch, N, Tn, euler, scale, impl, hidden):
    """Oracle = torch autograd (fp64, CPU) over the restated RK4-MLP rollout (the reference has neither).
    The VJP kernels: tcgen05 (default for hidden = 64), CUDA cores (DYNAX_NODE_VJP_IMPL=0; hidden = 32 always), and the
    wide CUDA-core kernel (hidden = 128: 32-trajectory CTAs, eight threads per trajectory)."""
    from dynax_b200 import ops
    if impl == "cuda_core":
        monkeypatch.setenv("DYNAX_NODE_VJP_IMPL", "0")
    Wnp = orc.synth_mlp(109, hidden=hidden, scale=scale)
    x0, u = inputs(N, Tn, 110)
    rng = np.random.default_rng(111)
    gx = rng.normal(size=(Tn, N, 3)).astype(np.float32); gy = rng.normal(size=(Tn, N, 1)).astype(np.float32)
    Wd = [T.tensor(w, dtype=T.float64, requires_grad=True) for w in Wnp]
    x0d = T.tensor(x0, dtype=T.float64, requires_grad=True); ud = T.tensor(u, dtype=T.float64, requires_grad=True)
    xs, ys = _torch_node_rollout(T, Wd, x0d, ud, 0.25, euler)
    ((xs * T.tensor(gx, dtype=T.float64)).sum() + (ys * T.tensor(gy, dtype=T.float64)).sum()).backward()
    Wc = [cu(T, w).requires_grad_(True) for w in Wnp]
    x0c = cu(T, x0).requires_grad_(True); uc = cu(T, u).requires_grad_(True)
    xs_c, ys_c = ops.node_rollout(*Wc, x0c, uc, 0.25, euler=euler)
    ((xs_c * cu(T, gx)).sum() + (ys_c * cu(T, gy)).sum()).backward()
    for got, ref, nm in zip(Wc + [x0c, uc], Wd + [x0d, ud], ["W1", "b1", "W2", "b2", "W3", "b3", "x0", "u"]):
        assert rel_err(got.grad.cpu().numpy(), ref.grad.numpy(), floor=1.0) < 1e-4, nm


@pytest.mark.parametrize("impl,hidden", [("tc", 64), ("cuda_core", 64), ("cuda_core", 32), ("wide", 128)])
def test_vjp_with_saved_stage_points(T, monkeypatch, impl, hidden):
    """The forward kernel can save the RK4 stage points (stage_x[T,N,9]); the VJP kernels then skip the three
    forward evaluations per step that recompute them.  Same gradient within rounding (the recomputation sums the
    64 -> 3 output layer in a different order)."""
    from dynax_b200 import ops
    if impl == "cuda_core":
        monkeypatch.setenv("DYNAX_NODE_VJP_IMPL", "0")
    Wnp = orc.synth_mlp(120, hidden=hidden, scale=0.5)
    x0, u = inputs(300, 33, 121)
    W = [cu(T, w) for w in Wnp]
    xs, ys, _, st = ops.node_rk4_forward(*W, cu(T, x0), cu(T, u), 0.25, emit_stages=True)
    assert st is not None and st.shape == (33, 300, 9)
    # stage points are what the restated RK4 visits: z2 = x + dt/2 k1 ... checked through the oracle's rollout of ONE step
    rng = np.random.default_rng(122)
    gx = cu(T, rng.normal(size=(33, 300, 3))); gy = cu(T, rng.normal(size=(33, 300, 1)))
    a = ops.node_rk4_vjp(*W, cu(T, x0), cu(T, u), xs, 0.25, gx, gy)
    b = ops.node_rk4_vjp(*W, cu(T, x0), cu(T, u), xs, 0.25, gx, gy, stage_x=st)
    for p_, q_, nm in zip(a, b, ["W1", "b1", "W2", "b2", "W3", "b3", "x0", "u"]):
        assert rel_err(q_.cpu().numpy(), p_.cpu().numpy(), floor=1.0) < 2e-5, nm
    # Euler has no stage points; a shared input series emits them like a per-system one (they are per trajectory)
    assert ops.node_rk4_forward(*W, cu(T, x0), cu(T, u), 0.25, euler=True, emit_stages=True)[3] is None
    st1 = ops.node_rk4_forward(*W, cu(T, x0), cu(T, u[:, 0]), 0.25, emit_stages=True)[3]
    st2 = ops.node_rk4_forward(*W, cu(T, x0), cu(T, np.repeat(u[:, :1], 300, 1)), 0.25, emit_stages=True)[3]
    assert st1 is not None and T.equal(st1, st2)


@pytest.mark.parametrize("impl,hidden", [("tc", 64), ("cuda_core", 64), ("cuda_core", 32), ("wide", 128)])
@pytest.mark.parametrize("shape", ["T1m", "Tm"])
def test_vjp_shared_input_series(T, monkeypatch, shape, impl, hidden):
    """A shared input series (u[T,1,5] or u[T,5] with N > 1 systems) through autograd: forward and VJP kernels take it as
    ONE series (DYNAX_SHARED_U); g_u comes back in the caller's shape, summed over the systems inside the kernel
    (round-1 advisor finding: u[T,1,5] reached the kernel as a T*5 buffer)."""
    from dynax_b200 import ops
    if impl == "cuda_core":
        monkeypatch.setenv("DYNAX_NODE_VJP_IMPL", "0")
    N, Tn = 70, 19
    Wnp = orc.synth_mlp(130, hidden=hidden, scale=0.5)
    x0, u = inputs(N, Tn, 131)
    u1 = u[:, :1] if shape == "T1m" else u[:, 0]
    rng = np.random.default_rng(132)
    gx = rng.normal(size=(Tn, N, 3)).astype(np.float32); gy = rng.normal(size=(Tn, N, 1)).astype(np.float32)
    Wd = [T.tensor(w, dtype=T.float64, requires_grad=True) for w in Wnp]
    x0d = T.tensor(x0, dtype=T.float64, requires_grad=True); ud = T.tensor(u1, dtype=T.float64, requires_grad=True)
    ub = ud.expand(Tn, N, 5) if shape == "T1m" else ud[:, None, :].expand(Tn, N, 5)
    xs, ys = _torch_node_rollout(T, Wd, x0d, ub, 0.25, False)
    ((xs * T.tensor(gx, dtype=T.float64)).sum() + (ys * T.tensor(gy, dtype=T.float64)).sum()).backward()
    Wc = [cu(T, w).requires_grad_(True) for w in Wnp]
    x0c = cu(T, x0).requires_grad_(True); uc = cu(T, u1).requires_grad_(True)
    xs_c, ys_c = ops.node_rollout(*Wc, x0c, uc, 0.25)
    ((xs_c * cu(T, gx)).sum() + (ys_c * cu(T, gy)).sum()).backward()
    assert uc.grad.shape == uc.shape
    for got, ref, nm in zip(Wc + [x0c, uc], Wd + [x0d, ud], ["W1", "b1", "W2", "b2", "W3", "b3", "x0", "u"]):
        assert rel_err(got.grad.cpu().numpy(), ref.grad.numpy(), floor=1.0) < 1e-4, nm


def test_golden_node_vector_on_gpu(T, golden_dir):
    """The committed golden vector (written by the torch fp64 implementation, tests/golden/make_golden_node.py)."""
    import json, os
    from dynax_b200 import ops
    g = json.load(open(os.path.join(golden_dir, "golden_node.json")))
    W = [cu(T, np.array(g[k], dtype=np.float32)) for k in ("W1", "b1", "W2", "b2", "W3", "b3")]
    x0, u = cu(T, np.array(g["x0"])), cu(T, np.array(g["u"]))
    for key, euler in (("xsol_rk4", False), ("xsol_euler", True)):
        xs, _, _ = ops.node_rk4_forward(*W, x0, u, g["dt"], euler=euler)
        assert rel_err(xs.cpu().numpy(), np.array(g[key]), floor=1.0) < 1e-5, key


@pytest.mark.parametrize("dt", [1.0, 900.0])
def test_config5_horizon_T2048(T, dt):
    """BASELINE.json config 5's horizon (T = 2048, normalised inputs, LeCun x 0.1 weights) against the fp64 oracle.
    dt = 1: every trajectory within the tolerance 1e-5.
    dt = 900: dt*|J| reaches ~1, the map amplifies perturbations, and ANY fp32 arithmetic
    loses a tail of the trajectories -- two orders of evaluation of the NumPy fp32 restatement differ from fp64 by
    9e-7 (median) / 6e-6 (90 %) / 2e-4 (99 %) / 6e-3 (worst) per trajectory (measured, N = 512).  The tcgen05 kernel's
    per-operation error (3xTF32 products at 2^-21, tanh at 5e-7 absolute) is a few times fp32's and shows as the same
    factor at every quantile (measured: 2.8e-6 / 4.6e-5 / 3e-3 / 1).  So the bar there is per quantile: median within
    1e-5, 90th percentile no worse than 10x the fp32 restatement's."""
    from dynax_b200 import ops
    N, Tn = 512, 2048
    W = orc.synth_mlp(4567)
    x0, u = inputs(N, Tn, 4568)
    xs, ys, _ = ops.node_rk4_forward(*(cu(T, w) for w in W), cu(T, x0), cu(T, u), dt)
    x64, y64 = orc.node_rk4_rollout(*W, x0, u, dt, np.float64)
    den = np.abs(x64).max(axis=(0, 2))
    e = np.abs(xs.cpu().numpy() - x64).max(axis=(0, 2)) / den            # per trajectory, inf-norm relative
    if dt == 1.0:
        assert e.max() < 1e-5, e.max()
        assert rel_err(ys.cpu().numpy(), y64, floor=1.0) < 1e-5
    else:
        x32, _ = orc.node_rk4_rollout(*W, x0, u, dt, np.float32)
        f = np.abs(x32 - x64).max(axis=(0, 2)) / den
        q = lambda v: np.quantile(v, [0.5, 0.9, 0.99, 1.0])
        print("config5 dt=900 per-trajectory error quantiles (50/90/99/100 %): ours", q(e), " fp32 restatement", q(f))
        assert np.median(e) < 1e-5, q(e)
        assert np.quantile(e, 0.9) < max(1e-5, 10 * np.quantile(f, 0.9)), (q(e), q(f))


def test_config5_full_size_properties(T):
    """N = 262 144 trajectories x T = 2048 (config 5 as stated): size-independent properties -- T steps == T1 steps
    then T - T1 steps from x_final (bit for bit), a permutation of the systems permutes the outputs (bit for bit:
    a trajectory does not depend on its position in the tile) -- and 64 sampled trajectories against the fp64 oracle."""
    from dynax_b200 import ops
    N, Tn, T1, dt = 262144, 2048, 777, 1.0
    Wnp = orc.synth_mlp(4567)
    W = [cu(T, w) for w in Wnp]
    g = T.Generator(device="cuda").manual_seed(4569)
    x0 = T.randn((N, 3), device="cuda", generator=g)
    u = T.randn((Tn, N, 5), device="cuda", generator=g)
    xs, ys, xf = ops.node_rk4_forward(*W, x0, u, dt, emit_final=True)
    assert bool(T.isfinite(xf).all())
    idx = np.sort(np.random.default_rng(7).choice(N, 64, replace=False))
    sel = T.as_tensor(idx, device="cuda")
    x64, y64 = orc.node_rk4_rollout(*Wnp, x0[sel].cpu().numpy(), u[:, sel].cpu().numpy(), dt, np.float64)
    assert rel_err(xs[:, sel].cpu().numpy(), x64, floor=1.0) < 1e-5
    assert rel_err(ys[:, sel].cpu().numpy(), y64, floor=1.0) < 1e-5
    xa, _, xfa = ops.node_rk4_forward(*W, x0, u[:T1], dt, emit_x=False, emit_y=False, emit_final=True)
    xb, _, xfb = ops.node_rk4_forward(*W, xfa, u[T1:], dt, emit_x=False, emit_y=False, emit_final=True)
    assert T.equal(xfb, xf)
    del xs, ys
    perm = T.randperm(4096, device="cuda", generator=g)                       # permute the first 4096 systems
    xp, _, _ = ops.node_rk4_forward(*W, x0[:4096][perm].contiguous(), u[:512, :4096][:, perm].contiguous(), dt)
    xq, _, _ = ops.node_rk4_forward(*W, x0[:4096].contiguous(), u[:512, :4096].contiguous(), dt)
    assert T.equal(xp, xq[:, perm])
