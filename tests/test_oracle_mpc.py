"""CPU: the MPC-cost oracle (config 4) is pinned for a one-step horizon by the reference's env
golden (tests/test_gymwrapper_basics.py:8 REWARD) and is linear in the price / comfort weights."""
import json
import os

import numpy as np

from oracle import dynax_oracle as orc


def test_one_step_cost_is_minus_golden_reward(golden_dir):
    g = json.load(open(os.path.join(golden_dir, "golden2.json")))
    r0 = g["disturbance_row0"]
    d = np.array([[r0["out_temp"], r0["qint_lump"], r0["qwin_lump"], r0["qradin_lump"]]], np.float32)
    u = np.array([[-3.0]], np.float32)      # action 7 of 11 -> -3 kW (rc.py:166)
    cost = orc.mpc_cost(g["theta"], g["x0"], d, u, g["dt"], g["price"], orc.RC_T_LOW, orc.RC_T_HIGH,
                        g["COP"], g["weights"][0], g["weights"][1])
    assert abs(float(cost[0]) + g["REWARD"]) < 1e-7
    assert np.allclose(orc.RC_PRICE, g["price"])


def test_cost_terms():
    rng = np.random.default_rng(0)
    H, N = 96, 50
    d = np.stack([15 + 5 * rng.random(H), rng.random(H), rng.random(H), rng.random(H)], 1)
    u = -10 * rng.random((H, N))
    e = orc.mpc_cost(orc.RC_NOMINAL, [20, 35.8, 26], d, u, 900.0, w_comfort=0.0, dtype=np.float64)
    c = orc.mpc_cost(orc.RC_NOMINAL, [20, 35.8, 26], d, u, 900.0, w_energy=0.0, dtype=np.float64)
    both = orc.mpc_cost(orc.RC_NOMINAL, [20, 35.8, 26], d, u, 900.0, dtype=np.float64)
    assert np.allclose(e + c, both) and (e > 0).all() and (c >= 0).all()
