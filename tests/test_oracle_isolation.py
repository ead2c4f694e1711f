"""The oracle is test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's CPU arm may use it."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PAT = re.compile(r"^\s*(from\s+oracle\b|import\s+oracle\b|from\s+\.+oracle\b)|libdynax_oracle", re.M)


def _sources(*dirs):
    for d in dirs:
        for base, _, files in os.walk(os.path.join(ROOT, d)):
            if "__pycache__" in base or os.sep + "_build" in base:
                continue
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".cpp")):
                    yield os.path.join(base, f)


def test_product_examples_and_tools_never_touch_the_oracle():
    bad = [p for p in _sources("dynax_b200", "examples", "tools", "ffi", "include") if PAT.search(open(p).read())]
    assert not bad, f"oracle used outside tests/bench/smoke: {bad}"


def test_bench_uses_the_oracle_only_in_the_cpu_arm():
    src = open(os.path.join(ROOT, "bench.py")).read()
    hits = [m.start() for m in re.finditer(r"from oracle import", src)]
    assert hits, "bench.py's CPU arm is expected to time the oracle port"
    for h in hits:  # every import sits inside the function that times the CPU port, or inside an in-run parity
        fn = re.findall(r"^def (\w+)\(", src[:h], re.M)[-1]     # sample (parity_*: the oracle as the CHECKER of a timed run)
        assert "cpu" in fn or "reference" in fn or fn.startswith("parity_"), fn


def test_tools_synthetic_inputs_match_the_test_generators():
    """tools/_synth.py restates the generators the parity tests use, so timings and tests see the same inputs."""
    import sys

    import numpy as np
    sys.path.insert(0, ROOT)
    from oracle import dynax_oracle as orc
    from tools import _synth as syn
    assert np.array_equal(syn.synth_theta(9, 3), orc.synth_theta(9, 3))
    assert np.array_equal(syn.synth_x0(9, 3), orc.synth_x0(9, 3))
    assert np.array_equal(syn.synth_u(50, 6, 3), orc.synth_u(50, 6, 3))
    assert all(np.array_equal(a, b) for a, b in zip(syn.synth_mlp(7), orc.synth_mlp(7)))
    for name in ("RC_NOMINAL", "RC_LB", "RC_UB", "RC_PRICE", "RC_T_LOW", "RC_T_HIGH"):
        assert np.array_equal(getattr(syn, name), getattr(orc, name)), name
