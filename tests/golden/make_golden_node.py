"""Writes tests/golden/golden_node.json: a small MLP-dynamics rollout (8 -> 16 -> 16 -> 3 tanh MLP would not match the
instantiated kernels, so the real 8 -> 64 -> 64 -> 3 shape is used with N=3 systems, T=6 steps) computed by the TORCH
fp64 implementation of tests/test_oracle_node.py -- independent of oracle/dynax_oracle.py, which is then checked
against it.  The reference cannot provide this vector: examples/4-neural-ode.py does not run.

    python tests/golden/make_golden_node.py
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.test_oracle_node import TorchNode, torch_rollout  # noqa: E402

rng = np.random.default_rng(20261020)
mk = lambda o, i, s: (s * rng.normal(size=(o, i)) / np.sqrt(i)).astype(np.float32)
W = (mk(64, 8, 0.5), (0.01 * rng.normal(size=64)).astype(np.float32), mk(64, 64, 0.5), (0.01 * rng.normal(size=64)).astype(np.float32),
     mk(3, 64, 0.5), (0.01 * rng.normal(size=3)).astype(np.float32))
x0 = rng.normal(size=(3, 3)).astype(np.float32)
u = rng.normal(size=(6, 3, 5)).astype(np.float32)
dt = 0.25
net = TorchNode(W)
t64 = lambda a: torch.tensor(a, dtype=torch.float64)
out = {"dt": dt, "x0": x0.tolist(), "u": u.tolist()}
for k, w in zip(("W1", "b1", "W2", "b2", "W3", "b3"), W):
    out[k] = w.tolist()
out["xsol_rk4"] = torch_rollout(net, t64(x0), t64(u), dt, False)[0].detach().numpy().tolist()
out["xsol_euler"] = torch_rollout(net, t64(x0), t64(u), dt, True)[0].detach().numpy().tolist()
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_node.json"), "w"))
print("wrote golden_node.json")
