"""The jax.ffi shim (ffi/dynax_ffi.cc) cannot be compiled here (no JAX headers), so this test keeps it
honest the cheap way: every C-ABI function it calls must be declared in include/dynax_b200.h with the
same number of arguments, and every XLA handler it defines must be registered by ffi/dynax_ffi_b200.py."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _args(text, start):
    """Number of top-level comma-separated arguments of the parenthesis opening at text[start]."""
    depth, n, i, seen = 0, 0, start, False
    while True:
        c = text[i]
        if c == "(":
            depth += 1
        elif c == ")":
            depth -= 1
            if depth == 0:
                return n + (1 if seen else 0)
        elif c == "," and depth == 1:
            n += 1
        elif not c.isspace() and depth == 1:
            seen = True
        i += 1


def _strip_comments(s):
    return re.sub(r"//[^\n]*", "", re.sub(r"/\*.*?\*/", "", s, flags=re.S))


def test_shim_calls_match_header():
    hdr = _strip_comments(open(os.path.join(ROOT, "include", "dynax_b200.h")).read())
    src = _strip_comments(open(os.path.join(ROOT, "ffi", "dynax_ffi.cc")).read())
    decl = {m.group(1): (0 if re.match(r"\(\s*void\s*\)", hdr[m.end() - 1:]) else _args(hdr, m.end() - 1))
            for m in re.finditer(r"\b(dynax_\w+)\s*\(", hdr)}
    calls = [(m.group(1), _args(src, m.end() - 1)) for m in re.finditer(r"\b(dynax_\w+_f32|dynax_last_error)\s*\(", src)]
    assert {c[0] for c in calls} >= {"dynax_rc_forward_f32", "dynax_rc_vjp_f32", "dynax_rc_loss_grad_f32",
                                     "dynax_rc_loss_grad_ctrl_f32", "dynax_lti_forward_f32", "dynax_lti_vjp_f32",
                                     "dynax_node_rk4_forward_stages_f32", "dynax_node_rk4_vjp_stages_f32",
                                     "dynax_rc_mpc_cost_f32"}
    for name, n in calls:
        assert name in decl, f"{name} is not declared in include/dynax_b200.h"
        assert decl[name] == n, f"{name}: shim passes {n} arguments, header declares {decl[name]}"


def test_python_front_end_registers_every_handler():
    src = open(os.path.join(ROOT, "ffi", "dynax_ffi.cc")).read()
    py = open(os.path.join(ROOT, "ffi", "dynax_ffi_b200.py")).read()
    handlers = re.findall(r"XLA_FFI_DEFINE_HANDLER_SYMBOL\(\s*(\w+)", src)
    assert len(handlers) == 9
    for h in handlers:
        assert f'"{h}"' in py, f"handler {h} is never registered / called"
    # op codes used for the workspace query agree with the header
    hdr = open(os.path.join(ROOT, "include", "dynax_b200.h")).read()
    assert re.search(r"#define DYNAX_OP_RC_LOSS_GRAD 1\b", hdr) and re.search(r"#define DYNAX_OP_RC_VJP 2\b", hdr)


def test_shim_type_checks_against_a_mock_of_the_xla_header():
    """`g++ -fsyntax-only` of the shim against tests/ffi_stub/xla/ffi/api/ffi.h -- a mock that models the binding
    builder, so every handler's parameter list must be what its Ffi::Bind() chain implies and every C-ABI call must
    pass the argument types include/dynax_b200.h declares.  (The real header needs JAX, absent from this image.)"""
    import shutil
    import subprocess
    gxx = shutil.which("g++")
    assert gxx, "g++ is part of the image"
    cuda_inc = "/usr/local/cuda/include"
    r = subprocess.run([gxx, "-std=c++17", "-fsyntax-only", "-I", os.path.join(ROOT, "tests", "ffi_stub"), "-I",
                        os.path.join(ROOT, "include"), "-I", cuda_inc, os.path.join(ROOT, "ffi", "dynax_ffi.cc")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_python_front_end_is_valid_python():
    import ast
    ast.parse(open(os.path.join(ROOT, "ffi", "dynax_ffi_b200.py")).read())
