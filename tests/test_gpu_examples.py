"""GPU: the mirrored example scripts (examples/1-forward.py, 2-simulator-as-layer.py, 3-inverse.py, 4-neural-ode.py)
run end to end."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    r = subprocess.run([sys.executable, *args], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def test_example_forward():
    out = _run("examples/1-forward.py")
    assert "torch.Size([672, 3]) torch.Size([672, 1])" in out and "system-steps/s" in out


def test_example_inverse_decreases_loss():
    out = _run("examples/3-inverse.py", "150", "64")
    line = [l for l in out.splitlines() if l.startswith("initial loss")][0]
    l0, l1 = float(line.split()[4]), float(line.split()[6])
    assert l1 < l0
    os.remove(os.path.join(ROOT, "examples", "zone_coefficients.json"))


def test_example_simulator_as_layer():
    out = _run("examples/2-simulator-as-layer.py")
    assert "torch.Size([672, 3]) torch.Size([672, 1])" in out
    line = [l for l in out.splitlines() if l.startswith("layer training")][0]
    l0, l1 = float(line.split()[4]), float(line.split()[8])
    assert l1 < l0


def test_example_neural_ode_trains():
    out = _run("examples/4-neural-ode.py", "60", "1024")
    line = [l for l in out.splitlines() if l.startswith("initial loss")][0]
    l0, l1 = float(line.split()[2]), float(line.split()[6])
    assert l1 < 0.5 * l0
