"""Synthetic inputs and reference constants for the timing scripts under tools/.

The timing scripts measure the product path only, so they take their inputs from here and never from `oracle/`
(the oracle is the checker of `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU arm).
Constants restate the reference: nominal zone parameters `examples/1-forward.py:41-43` (== `dynax/wrapper/envs/rc.py:25-27`),
parameter bounds `examples/3-inverse.py:66-85`, comfort band and tariff `dynax/wrapper/envs/rc.py:32-42`.
"""
import numpy as np

from dynax_b200.wrapper.envs.rc import PARAMS, PRICE, T_HIGH, T_LOW

RC_NOMINAL = np.array(PARAMS, dtype=np.float64)  # order Cai, Cwe, Cwi, Re, Ri, Rw, Rg (RC.py:176-182)
RC_LB = np.array([3.0e3, 5.0e4, 5.0e5, 1.0, 0.5, 1.0, 5.0])
RC_UB = np.array([3.0e4, 5.0e5, 5.0e6, 1.0e1, 5.0, 1.0e1, 5.0e1])
RC_PRICE, RC_T_LOW, RC_T_HIGH = np.array(PRICE), np.array(T_LOW), np.array(T_HIGH)


def synth_theta(N, seed, spread=0.2):
    """Per-building parameters: nominal * exp(U(-spread, spread)) (Euler-stable at dt = 3600 s for spread <= 0.2)."""
    rng = np.random.default_rng(seed)
    return (RC_NOMINAL[None] * np.exp(rng.uniform(-spread, spread, size=(N, 7)))).astype(np.float32)


def synth_x0(N, seed):
    rng = np.random.default_rng(seed + 1)
    return (np.array([20.0, 30.0, 26.0])[None] + rng.uniform(-1, 1, size=(N, 3))).astype(np.float32)


def synth_u(T, N, seed):
    """u[T,N,5] fp32 = [Tao, qCon_i, qHVAC, qRad_e, qRad_i]: a yearly + daily outdoor wave and bounded random gains."""
    rng = np.random.default_rng(seed + 2)
    hour = np.arange(T, dtype=np.float64)[:, None]
    day = 2 * np.pi * hour / 24
    u = np.empty((T, N, 5), dtype=np.float32)
    u[..., 0] = (15 + 10 * np.sin(2 * np.pi * hour / 8760) + 6 * np.sin(day + rng.uniform(0, 2 * np.pi, size=(1, N)))
                 + rng.normal(size=(T, N)))
    u[..., 1] = rng.uniform(0, 1, size=(T, N))
    u[..., 2] = -rng.uniform(0, 5, size=(T, N))
    u[..., 3] = rng.uniform(0, 2, size=(T, N)) * (np.sin(day) > 0)
    u[..., 4] = rng.uniform(0, 1, size=(T, N))
    return u


def synth_mlp(seed, n=3, m=5, hidden=64, scale=0.1):
    """Weights of the 8 -> 64 -> 64 -> 3 dynamics network as row-major [out, in] matrices, small enough to stay stable."""
    rng = np.random.default_rng(seed)

    def dense(o, i):
        return (scale * rng.normal(size=(o, i)) / np.sqrt(i)).astype(np.float32)

    def bias(o):
        return (0.01 * rng.normal(size=o)).astype(np.float32)

    W1, b1, W2, b2 = dense(hidden, n + m), bias(hidden), dense(hidden, hidden), bias(hidden)
    return W1, b1, W2, b2, dense(n, hidden), np.zeros(n, np.float32)
