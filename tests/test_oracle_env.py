"""CPU: the env-step oracle reproduces the reference's known-answer test (tests/test_gymwrapper_basics.py:7-43)."""
import json
import os

import numpy as np

from oracle import dynax_oracle as orc


def golden2_table(g):
    r0 = g["disturbance_row0"]
    row = [r0["out_temp"], r0["qint_lump"], r0["qwin_lump"], r0["qradin_lump"]]
    return np.array([0.0, 60.0, 120.0]), np.array([row, row, row])       # 1-min table; only row 0 is pinned


def test_env_step_golden2(golden_dir):
    g = json.load(open(os.path.join(golden_dir, "golden2.json")))
    ts, dist = golden2_table(g)
    r = orc.rc_env_step(np.array(g["theta"]), np.array([g["x0"]]), np.array([0.0]), np.array([g["action"]]), ts, dist,
                        g["dt"], g["price"], orc.RC_T_LOW, orc.RC_T_HIGH, g["COP"], g["weights"], g["u_low"], g["u_high"],
                        g["num_actions"])
    assert np.allclose(r["obs"][0], g["OBS_NEXT"], rtol=1e-6)                     # test_gymwrapper_basics.py:38
    assert abs(float(r["reward"][0]) - g["REWARD"]) < 1e-7                       # :39
    assert r["terminated"][0] == 0.0 and r["t_next"][0] == g["TIME_NEXT"]         # :40,:42
    assert np.allclose(r["x_next"][0], g["STATES_NEXT"], rtol=0, atol=2e-6)       # :43
