"""Pins the neural-ODE oracle (oracle/dynax_oracle.py::node_rk4_rollout, a NumPy restatement) with a SECOND,
independently written implementation: torch fp64 ``nn.Linear`` modules + a hand-written classical RK4 / Euler loop,
differentiated by torch autograd.  The reference has only the dispatch for general fx/fy models
(dynax/core/base_block_state_space.py:62-63,72-73; its example examples/4-neural-ode.py does not run), so parity of
this path cannot be pinned by reference outputs; two independent restatements agreeing to 1e-12, plus the committed
golden vector below, are what the GPU tests lean on (tests/test_gpu_node.py)."""
import json
import os

import numpy as np
import torch

from oracle import dynax_oracle as orc


class TorchNode(torch.nn.Module):
    """rhs = Dense(Dense(Dense([x;u]).tanh()).tanh()) as three nn.Linear (weights [out,in]); y = x[:, :1]."""

    def __init__(self, W):
        super().__init__()
        W1, b1, W2, b2, W3, b3 = (torch.tensor(w, dtype=torch.float64) for w in W)
        self.l1 = torch.nn.Linear(W1.shape[1], W1.shape[0]).double()
        self.l2 = torch.nn.Linear(W2.shape[1], W2.shape[0]).double()
        self.l3 = torch.nn.Linear(W3.shape[1], W3.shape[0]).double()
        with torch.no_grad():
            for lin, w, b in ((self.l1, W1, b1), (self.l2, W2, b2), (self.l3, W3, b3)):
                lin.weight.copy_(w); lin.bias.copy_(b)

    def forward(self, x, u):
        return self.l3(torch.tanh(self.l2(torch.tanh(self.l1(torch.cat([x, u], -1))))))


def torch_rollout(net, x0, u, dt, euler=False):
    x, xs, ys = x0, [], []
    for k in range(u.shape[0]):
        uk = u[k].expand(x.shape[0], -1)
        ys.append(x[:, :1])                                   # output of the PRE-update state (simulator.py:38)
        k1 = net(x, uk)
        if euler:
            x = x + dt * k1                                   # simulator.py:40
        else:                                                 # classical RK4, u held over the step
            k2 = net(x + 0.5 * dt * k1, uk)
            k3 = net(x + 0.5 * dt * k2, uk)
            k4 = net(x + dt * k3, uk)
            x = x + dt / 6.0 * (k1 + 2.0 * k2 + 2.0 * k3 + k4)
        xs.append(x)
    return torch.stack(xs), torch.stack(ys)


def _case(seed, N, T, scale):
    W = orc.synth_mlp(seed, scale=scale)
    rng = np.random.default_rng(seed + 1)
    return W, rng.normal(size=(N, 3)).astype(np.float32), rng.normal(size=(T, N, 5)).astype(np.float32)


def test_numpy_restatement_equals_torch_modules():
    for seed, N, T, dt, euler, scale in ((1, 17, 60, 0.25, False, 0.1), (2, 5, 40, 0.1, False, 1.0), (3, 9, 33, 0.25, True, 0.5),
                                         (4, 4, 300, 900.0, False, 0.1)):
        W, x0, u = _case(seed, N, T, scale)
        xr, yr = orc.node_rk4_rollout(*W, x0, u, dt, np.float64, euler=euler)
        xs, ys = torch_rollout(TorchNode(W), torch.tensor(x0, dtype=torch.float64), torch.tensor(u, dtype=torch.float64), dt, euler)
        assert np.allclose(xs.detach().numpy(), xr, rtol=1e-11, atol=1e-12 * np.abs(xr).max())
        assert np.allclose(ys.detach().numpy(), yr, rtol=1e-11, atol=1e-12 * np.abs(yr).max())
    # a shared input series broadcasts over systems in both
    W, x0, u = _case(5, 6, 20, 0.3)
    xr, _ = orc.node_rk4_rollout(*W, x0, u[:, :1], 0.25, np.float64)
    xs, _ = torch_rollout(TorchNode(W), torch.tensor(x0, dtype=torch.float64), torch.tensor(u[:, :1], dtype=torch.float64), 0.25)
    assert np.allclose(xs.detach().numpy(), xr, rtol=1e-11)


def test_golden_node_vector(golden_dir):
    """tests/golden/golden_node.json was written by tests/golden/make_golden_node.py from the TORCH implementation;
    the NumPy oracle must reproduce it (and so must the CUDA kernels, tests/test_gpu_node.py)."""
    g = json.load(open(os.path.join(golden_dir, "golden_node.json")))
    W = [np.array(g[k], dtype=np.float32) for k in ("W1", "b1", "W2", "b2", "W3", "b3")]
    x0, u = np.array(g["x0"], np.float32), np.array(g["u"], np.float32)
    for key, euler in (("xsol_rk4", False), ("xsol_euler", True)):
        xr, _ = orc.node_rk4_rollout(*W, x0, u, g["dt"], np.float64, euler=euler)
        assert np.allclose(xr, np.array(g[key]), rtol=1e-12, atol=1e-14)
