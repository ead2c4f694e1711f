"""dynax_b200 -- B200-native batched rollout engine behind dynax's module API.

Mirrors the reference package layout for the hot path only:

    dynax_b200.core        BaseBlockSSM, Discrete/ContinuousLinearSSM   (dynax/core/*)
    dynax_b200.models      Continuous4R3C, Discrete4R3C                 (dynax/models/RC.py)
    dynax_b200.simulators  DifferentiableSimulator                      (dynax/simulators/simulator.py)
    dynax_b200.estimators  fused MSE loss + adjoint gradient            (examples/3-inverse.py:96-112)
    dynax_b200.ops         tensor-level calls into libdynax_b200.so

Arrays are torch CUDA tensors (JAX is not installable in the build image; INTEGRATION.md holds the
jax.ffi binding).  All compute happens in libdynax_b200.so; importing without it raises.
"""
from . import _lib

_lib.lib()  # fail loudly at import time if the CUDA library has not been built

from . import agents, core, estimators, models, ops, simulators, utils  # noqa: E402
from .simulators import DifferentiableSimulator  # noqa: E402

__all__ = ["agents", "core", "models", "simulators", "estimators", "ops", "utils", "DifferentiableSimulator"]
