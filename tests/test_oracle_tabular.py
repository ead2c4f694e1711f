"""CPU: the interpolation oracle reproduces the reference's own table (tests/test_agents.py:13-25)."""
import json
import os

import numpy as np

from oracle import dynax_oracle as orc


def test_tabular_golden(golden_dir):
    g = json.load(open(os.path.join(golden_dir, "golden_tabular.json")))
    assert np.allclose(orc.tabular_interp(g["ts"], g["xs"], g["at"], "constant"), g["constant"])
    assert np.allclose(orc.tabular_interp(g["ts"], g["xs"], g["at"], "linear"), g["linear"])
    # clamping outside the table
    assert np.allclose(orc.tabular_interp(g["ts"], g["xs"], [-1.0, 9.0], "linear"), [g["xs"][0], g["xs"][-1]])
