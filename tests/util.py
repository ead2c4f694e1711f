"""Shared helpers for the GPU parity tests."""
import numpy as np


def rel_err(a, ref):
    """|a-ref| / max(|ref|, 1e-3*max|ref|), max over elements."""
    a = np.asarray(a, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    den = np.maximum(np.abs(ref), 1e-3 * np.abs(ref).max() + 1e-300)
    return float((np.abs(a - ref) / den).max())


def grad_ok(g, g64, g32_ref=None, rtol=1e-4):
    """SURVEY.md section 7 H2 acceptance: norm-wise rel <= rtol per system AND per element
    |g-g64| <= max(rtol*|g64|, |g32_ref-g64|): at least as close to fp64 truth as the
    all-fp32 reference-order BPTT is."""
    g = np.asarray(g, np.float64); g64 = np.asarray(g64, np.float64)
    nrm = np.linalg.norm(g - g64, axis=-1) / np.linalg.norm(g64, axis=-1)
    tol = rtol * np.abs(g64)
    if g32_ref is not None:
        tol = np.maximum(tol, np.abs(np.asarray(g32_ref, np.float64) - g64))
    elem = np.abs(g - g64) <= tol + 1e-30
    return float(nrm.max()), bool(elem.all()), float((np.abs(g - g64) / (np.abs(g64) + 1e-300)).max())
