"""Regenerate tests/golden/*.json|npz from the read-only reference checkout.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

Nothing is computed here: the fixtures are the literal known-answer constants held by
the reference's own tests plus the input rows they were produced from.

* golden1.json  <- tests/test_simulator.py:21-36 (x0, RESULTS table), setup at :15-20,:40-44
* golden2.json  <- tests/test_gymwrapper_basics.py:7-12 (OBS_NEXT, REWARD, STATES_NEXT),
                   theta/price/T bounds from dynax/wrapper/envs/rc.py:25-42, inputs = row 0 of
                   resources/jax/data/{out_temp,qint_lump,qwin_lump,qradin_lump}.csv (q's /1000,
                   resources/jax/data/combine-data.py:4-8)
* golden_tabular.json <- tests/test_agents.py:13-25 (interpolation tables)
* c1_hourly.npz <- the examples/1-forward.py:19-29 input pipeline (1-min CSVs -> hourly means)
                   rebuilt from resources/jax/data/*.csv because examples/data/eplus_1min.csv
                   is absent from the checkout (.MISSING_LARGE_BLOBS).
"""
import ast
import json
import os

import numpy as np
import pandas as pd

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def _literal_assignments(path, names):
    """Pull `NAME = jnp.array(<literal>)` / `NAME = <literal>` constants out of a test file."""
    tree = ast.parse(open(path).read())
    found = {}
    for node in tree.body:
        if isinstance(node, ast.Assign) and isinstance(node.targets[0], ast.Name):
            name = node.targets[0].id
            if name not in names or name in found:  # first (literal) assignment wins
                continue
            val = node.value
            if isinstance(val, ast.Call):           # jnp.array([...])
                val = val.args[0]
            try:
                found[name] = ast.literal_eval(val)
            except ValueError:                      # e.g. PRICE = jnp.array(PRICE) re-binding
                pass
    return found


def main():
    g1 = _literal_assignments(os.path.join(REF, "tests/test_simulator.py"), {"RESULTS", "x0", "dt", "ts", "te"})
    g1["theta"] = [1.0] * 7           # RC.py:176-182 nn.initializers.ones
    g1["u"] = [[1.0] * 5] * 10        # test_simulator.py:40-44 (ones, T=10)
    g1["source"] = "tests/test_simulator.py:15-36"
    json.dump(g1, open(os.path.join(OUT, "golden1.json"), "w"), indent=1)

    g2 = _literal_assignments(os.path.join(REF, "tests/test_gymwrapper_basics.py"),
                              {"OBS_NEXT", "REWARD", "STATES_NEXT", "TIME_NEXT", "dt"})
    rc = _literal_assignments(os.path.join(REF, "dynax/wrapper/envs/rc.py"), {"params", "PRICE", "WEIGHTS"})
    row0 = {}
    for col in ("out_temp", "qint_lump", "qwin_lump", "qradin_lump"):
        df = pd.read_csv(os.path.join(REF, "resources/jax/data", col + ".csv"), index_col=[0])
        row0[col] = float(df.iloc[0, 0]) / (1.0 if col == "out_temp" else 1000.0)
    g2.update(theta=rc["params"], price=rc["PRICE"], weights=rc["WEIGHTS"], disturbance_row0=row0,
              x0=[20.0, 30.0, 26.0], action=7, num_actions=11, u_low=-10.0, u_high=0.0, COP=2.5,
              T_low_h0=12.0, T_high_h0=30.0,
              source="tests/test_gymwrapper_basics.py:7-43; dynax/wrapper/envs/rc.py:25-42,125-132,227-252")
    json.dump(g2, open(os.path.join(OUT, "golden2.json"), "w"), indent=1)

    gt = dict(ts=[0.0, 2.0, 3.0, 4.0], xs=[[0.0, 0.0], [1.0, 2.0], [4.0, 6.0], [9.0, 12.0]],
              at=[0.5, 1.5, 2.5, 3.5],
              constant=[[0.0, 0.0], [1.0, 2.0], [4.0, 6.0], [9.0, 12.0]],
              linear=[[0.25, 0.5], [0.75, 1.5], [2.5, 4.0], [6.5, 9.0]],
              source="tests/test_agents.py:13-25")
    json.dump(gt, open(os.path.join(OUT, "golden_tabular.json"), "w"), indent=1)

    # config 1 inputs: combine-data.py:3-11 then examples/1-forward.py:19-29
    d = os.path.join(REF, "resources/jax/data")
    cols = [("out_temp", 1.0), ("qint_lump", 1e-3), ("qhvac_lump", 1e-3), ("qwin_lump", 1e-3),
            ("qradin_lump", 1e-3), ("zone_temp", 1.0)]
    # positional concat (the files share one 1-minute time base; their Date/Time labels are not unique
    # so label alignment would inject NaNs)
    data = pd.concat([pd.read_csv(os.path.join(d, c + ".csv"), index_col=[0]).reset_index(drop=True) * s
                      for c, s in cols], axis=1)
    data.index = range(0, len(data) * 60, 60)
    hourly = data.groupby([data.index // 3600]).mean()
    np.savez_compressed(os.path.join(OUT, "c1_hourly.npz"),
                        u=hourly.values[:, :5].astype(np.float32), y=hourly.values[:, 5].astype(np.float32))
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
