"""Shared helpers for the GPU parity tests: the written-down tolerances.

north_star: trajectories within rel 1e-5, gradients within rel 1e-4 (fp32).
"""
import os

import numpy as np

TRAJ_RTOL = 1e-5
GRAD_RTOL = 1e-4
_MARGINS = os.environ.get("DYNAX_TEST_MARGINS")   # print how much of each tolerance a check used (pytest -s)


def _margin(name, frac):
    if _MARGINS:
        print(f"[margin] {name}: {frac:.3f} of tolerance  ({os.environ.get('PYTEST_CURRENT_TEST', '')})", flush=True)


def rel_err(a, ref, floor=1e-3):
    """max |a-ref| / max(|ref|, floor*max|ref|)   (SURVEY.md section 8(d) uses floor=1e-3 for
    temperatures; floor=1.0 is the plain inf-norm relative error, used for zero-mean signals)."""
    a = np.asarray(a, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    if ref.size == 0:
        return 0.0
    den = np.maximum(np.abs(ref), floor * np.abs(ref).max() + 1e-300)
    e = float((np.abs(a - ref) / den).max())
    _margin("rel_err/1e-5", e / TRAJ_RTOL)
    return e


def normwise(g, g64):
    g = np.asarray(g, np.float64); g64 = np.asarray(g64, np.float64)
    return np.linalg.norm(g - g64, axis=-1) / (np.linalg.norm(g64, axis=-1) + 1e-300)


def assert_loss_parity(loss, l64, loss_budget, what=""):
    """|loss - loss64| <= 1e-5*loss64 + budget, budget = what a 1e-5 trajectory perturbation can do
    to mean((y-target)^2) (oracle.rc_output_sensitivity_budget)."""
    loss = np.asarray(loss, np.float64); l64 = np.asarray(l64, np.float64)
    tol = TRAJ_RTOL * np.abs(l64) + loss_budget
    bad = np.abs(loss - l64) > tol
    _margin("loss", float((np.abs(loss - l64) / tol).max()))
    assert not bad.any(), (what, float((np.abs(loss - l64) / tol).max()))


def assert_grad_parity(g, g64, grad_budget, g32=None, what=""):
    """Gradient acceptance (DESIGN.md 'Tolerances'):
      (1) per element |g - g64| <= 1e-4*|g64| + budget_ij, where budget is the first-order change of
          the gradient under a 1e-5 relative perturbation of the trajectory -- the accuracy to which
          the north_star's own trajectory tolerance determines the gradient;
      (2) if the all-fp32 reference-order BPTT (oracle C port) is given: in aggregate we are at least
          as close to fp64 truth as it is: median of the norm-wise error no larger, and the worst system
          within 2x of ITS worst (the worst system is the most ill-conditioned one, where every fp32
          evaluation -- the reference's included -- lands a few 1e-4 from truth and which of them is
          closest is one draw of rounding noise; criterion (1) is what bounds it)."""
    g = np.asarray(g, np.float64); g64 = np.asarray(g64, np.float64)
    tol = GRAD_RTOL * np.abs(g64) + grad_budget
    ratio = np.abs(g - g64) / (tol + 1e-300)
    _margin("grad", float(ratio.max()))
    assert ratio.max() <= 1.0, (what, "worst element at %.2f of its tolerance" % ratio.max())
    if g32 is not None:
        no, nr = normwise(g, g64), normwise(g32, g64)
        _margin("grad median vs fp32 reference", float(np.median(no) / max(GRAD_RTOL, np.median(nr))))
        _margin("grad worst vs fp32 reference", float(no.max() / max(GRAD_RTOL, 2.0 * nr.max())))
        assert np.median(no) <= max(GRAD_RTOL, np.median(nr)), (what, np.median(no), np.median(nr))
        assert no.max() <= max(GRAD_RTOL, 2.0 * nr.max()), (what, no.max(), nr.max())
    return float(ratio.max())
