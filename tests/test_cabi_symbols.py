"""CPU-only: the C-ABI shared library loads, exports every symbol include/dynax_b200.h declares,
the ctypes table matches the header, and -- with no GPU here -- the product path FAILS LOUDLY
instead of falling back to anything."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dynax_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dynax_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def built():
    # load build.py by path: importing the package would (rightly) refuse a stale or missing library
    import importlib.util
    spec = importlib.util.spec_from_file_location("dynax_b200_build", os.path.join(ROOT, "dynax_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build()


def test_header_declares_expected_entry_points():
    fns = header_functions()
    for must in ("dynax_rc_forward_f32", "dynax_rc_loss_grad_f32", "dynax_rc_vjp_f32", "dynax_lti_forward_f32",
                 "dynax_lti_vjp_f32", "dynax_rc_forward_host_f32", "dynax_rc_loss_grad_host_f32"):
        assert must in fns


def test_library_exports_every_declared_symbol(built):
    out = subprocess.check_output(["nm", "-D", "--defined-only", built], text=True)
    exported = set(re.findall(r" T (dynax_[a-z0-9_]+)", out))
    missing = [f for f in header_functions() if f not in exported]
    assert not missing, missing


def test_ctypes_table_matches_header(built):
    from dynax_b200 import _lib
    assert sorted(_lib.SIGNATURES) == header_functions()
    L = _lib.lib()
    assert L.dynax_version() == 1


def test_sass_is_blackwell_native(built):
    """cuobjdump: the image is sm_100a and the streaming path uses TMA bulk copies + mbarriers."""
    arch = subprocess.check_output(["cuobjdump", "-lelf", built], text=True)
    assert "sm_100a" in arch
    sass = subprocess.run("cuobjdump -sass %s | grep -c -E 'UBLKCP|SYNCS'" % built, shell=True, capture_output=True, text=True)
    assert int(sass.stdout.strip() or 0) > 0


def _no_gpu():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


@pytest.mark.skipif(not _no_gpu(), reason="only meaningful without a GPU")
def test_no_device_fails_loudly(built):
    from dynax_b200 import _lib, host
    L = _lib.lib()
    assert L.dynax_device_check() == -5
    th = np.ones((4, 7), np.float32); x0 = np.ones((4, 3), np.float32); u = np.ones((6, 4, 5), np.float32)
    with pytest.raises(_lib.DynaxError) as e:
        host.rc_forward(th, x0, u, 1.0)
    assert e.value.code == -5 and "no CPU fallback" in str(e.value)
    with pytest.raises(_lib.DynaxError):
        host.rc_loss_grad(th, x0, u, np.ones((6, 4), np.float32), 1.0)
    # argument validation happens before any device work
    assert L.dynax_rc_forward_f32(None, None, None, None, None, None, -1, 5, 1.0, 0, None) == -1
    assert L.dynax_lti_supported(4, 3, 4) == 1 and L.dynax_lti_supported(9, 9, 9) == 0
    # algorithm-aware: the default (forward-mode) loss+gradient needs no scratch, the checkpointed adjoint
    # (general cotangents, or discrete stepping) 1 KB of fp64 sums + ceil(T/8) checkpoints of 3 floats per system
    assert L.dynax_workspace_bytes(1, 1024, 8760, 3, 5, 1, 0) == 0
    assert L.dynax_workspace_bytes(1, 1024, 8760, 3, 5, 1, _lib.SHARED_THETA) == 1024
    assert L.dynax_workspace_bytes(2, 1024, 8760, 3, 5, 1, 0) == 1024 + 4 * 1095 * 3 * 1024
    assert L.dynax_workspace_bytes(1, 1024, 8760, 3, 5, 1, _lib.DISCRETE) == 1024 + 4 * 1095 * 3 * 1024
