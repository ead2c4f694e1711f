"""ctypes binding of libdynax_b200.so -- the C-ABI declared in include/dynax_b200.h.

There is deliberately NO fallback: if the shared library is missing this module raises, and
if no sm_100 device is present every compute entry point returns DYNAX_ERR_NO_DEVICE, which
``check`` turns into a ``DynaxError``.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# DYNAX_B200_LIB: another build of the same library (A/B timing of kernel variants, tools/ab_tangent3.py)
LIB_PATH = os.environ.get("DYNAX_B200_LIB") or os.path.join(_HERE, "libdynax_b200.so")

# flags / ops (include/dynax_b200.h)
SHARED_THETA, SHARED_U, SHARED_TARGET, DISCRETE, SOLVER_EULER = 1, 2, 4, 8, 16
OP_RC_FORWARD, OP_RC_LOSS_GRAD, OP_RC_VJP, OP_LTI_FORWARD, OP_LTI_VJP = 0, 1, 2, 3, 4

_f = ctypes.c_void_p       # device or host float* passed as an integer address
_i64 = ctypes.c_int64
_u32 = ctypes.c_uint32
_int = ctypes.c_int
_flt = ctypes.c_float
_sz = ctypes.c_size_t

# name -> (restype, argtypes): every symbol include/dynax_b200.h declares
SIGNATURES = {
    "dynax_version": (_int, []),
    "dynax_last_error": (ctypes.c_char_p, []),
    "dynax_device_check": (_int, []),
    "dynax_debug_kernel_count": (_int, []),
    "dynax_debug_kernel_tag": (ctypes.c_char_p, [_int]),
    "dynax_debug_kernel_launches": (ctypes.c_longlong, [_int]),
    "dynax_debug_last_kernel": (ctypes.c_char_p, []),
    "dynax_workspace_bytes": (_sz, [_int, _i64, _i64, _int, _int, _int, _u32]),
    "dynax_lti_supported": (_int, [_int, _int, _int]),
    "dynax_rc_forward_f32": (_int, [_f] * 6 + [_i64, _i64, _flt, _u32, _f]),
    "dynax_rc_loss_grad_f32": (_int, [_f] * 9 + [_i64, _i64, _flt, _u32, _f, _sz, _f]),
    "dynax_rc_vjp_f32": (_int, [_f] * 8 + [_i64, _i64, _flt, _u32, _f, _sz, _f]),
    "dynax_lti_forward_f32": (_int, [_f] * 9 + [_i64, _i64, _int, _int, _int, _flt, _u32, _f]),
    "dynax_lti_forward_chunked_workspace": (_sz, [_i64, _i64, _int, _i64]),
    "dynax_lti_vjp_chunked_workspace": (_sz, [_i64, _i64, _int, _int, _int, _i64]),
    "dynax_lti_vjp_chunked_f32": (_int, [_f] * 14 + [_i64, _i64, _int, _int, _int, _flt, _u32, _i64, _f, _sz, _f]),
    "dynax_lti_forward_chunked_f32": (_int, [_f] * 9 + [_i64, _i64, _int, _int, _int, _flt, _u32, _i64, _f, _sz, _f]),
    "dynax_lti_vjp_f32": (_int, [_f] * 14 + [_i64, _i64, _int, _int, _int, _flt, _u32, _f, _sz, _f]),
    "dynax_rc_forward_host_f32": (_int, [_f] * 6 + [_i64, _i64, _flt, _u32, _i64]),
    "dynax_rc_loss_grad_host_f32": (_int, [_f] * 8 + [_i64, _i64, _flt, _u32, _i64]),
    "dynax_nonfinite_rows_f32": (_int, [_f, _i64, _int, _f, _f, _f]),
    "dynax_host_release": (_int, []),
    "dynax_rc_dataset_create_f32": (_int, [_f, _f, _i64, _i64, _u32, ctypes.POINTER(ctypes.c_void_p)]),
    "dynax_rc_dataset_loss_grad_f32": (_int, [_f, _f, _f, _f, _f, _flt, _u32, _f, _f]),
    "dynax_rc_dataset_destroy": (_int, [_f]),
    "dynax_host_register": (_int, [_f, _sz]),
    "dynax_host_unregister": (_int, [_f]),
    "dynax_node_rk4_vjp_workspace": (_sz, []),
    "dynax_node_rk4_vjp_f32": (_int, [_f] * 19 + [_i64, _i64, _int, _int, _int, _int, _flt, _u32, _f, _sz, _f]),
    "dynax_node_rk4_vjp_stages_f32": (_int, [_f] * 20 + [_i64, _i64, _int, _int, _int, _int, _flt, _u32, _f, _sz, _f]),
    "dynax_rc_loss_grad_chunked_workspace": (_sz, [_i64, _i64, _i64]),
    "dynax_rc_loss_grad_chunked_f32": (_int, [_f] * 9 + [_i64, _i64, _flt, _u32, _i64, _f, _sz, _f]),
    "dynax_lamb_update_f32": (_int, [_f] * 4 + [_i64, _i64, _flt, _flt, _flt, _flt, _f]),
    "dynax_resample_mean_f32": (_int, [_f, _f, _i64, _int, _i64, _f]),
    "dynax_lamb_update_dev_f32": (_int, [_f] * 4 + [_i64, _f, _flt, _flt, _flt, _flt, _f]),
    "dynax_rc_env_step_f32": (_int, [_f] * 6 + [_i64] + [_f] * 3 + [_flt] * 5 + [_int] + [_f] * 7 + [_i64, _i64, _flt, _u32, _f]),
    "dynax_rc_loss_grad_ctrl_host_f32": (_int, [_f] * 4 + [_int] + [_f] * 5 + [_i64, _i64, _flt, _u32, _i64]),
    "dynax_rc_forward_ctrl_f32": (_int, [_f] * 4 + [_int] + [_f] * 3 + [_i64, _i64, _flt, _u32, _f]),
    "dynax_rc_loss_grad_ctrl_f32": (_int, [_f] * 4 + [_int] + [_f] * 6 + [_i64, _i64, _flt, _u32, _f, _sz, _f]),
    "dynax_tabular_interp_f32": (_int, [_f, _f, _i64, _int, _f, _i64, _int, _f, _f]),
    "dynax_rc_forward_chunked_workspace": (_sz, [_i64, _i64, _i64]),
    "dynax_rc_vjp_chunked_workspace": (_sz, [_i64, _i64, _i64]),
    "dynax_rc_vjp_chunked_f32": (_int, [_f] * 8 + [_i64, _i64, _flt, _u32, _i64, _f, _sz, _f]),
    "dynax_rc_forward_chunked_f32": (_int, [_f] * 6 + [_i64, _i64, _flt, _u32, _i64, _f, _sz, _f]),
    "dynax_node_rk4_forward_f32": (_int, [_f] * 11 + [_i64, _i64, _int, _int, _int, _int, _flt, _u32, _f]),
    "dynax_node_rk4_forward_stages_f32": (_int, [_f] * 12 + [_i64, _i64, _int, _int, _int, _int, _flt, _u32, _f]),
    "dynax_rc_mpc_cost_f32": (_int, [_f] * 7 + [_flt] * 4 + [_f] * 3 + [_i64, _i64, _flt, _f, _sz, _f]),
}

_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python dynax_b200/build.py` "
                "(dynax_b200 has no CPU or PyTorch fallback)")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)       # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


class DynaxError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"dynax_b200 error {code}: {msg}")
        self.code = code


def check(rc):
    if rc != 0:
        raise DynaxError(rc, lib().dynax_last_error().decode(errors="replace"))
