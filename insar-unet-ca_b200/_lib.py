"""ctypes binding of libunetca_b200.so (C ABI declared in include/unetca_b200.h).

The header is the single source of truth: it is parsed here to set argtypes/restype of every exported
function, so a declaration that drifts from the library shows up as a missing symbol at load time.
There is no fallback of any kind: if the shared library is absent or an op fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
HEADER = os.path.join(ROOT, "include", "unetca_b200.h")
TUNING_HEADER = os.path.join(ROOT, "include", "unetca_b200_tuning.h")      # development knobs, not the drop-in ABI
LIB_PATH = os.path.join(_HERE, "libunetca_b200.so")

F32, BF16 = 0, 1

_DECL = re.compile(r"^(const char\*|int|void)\s+(unetca_\w+)\s*\(([^;]*?)\)\s*;", re.M | re.S)


def parse_header(path: str = HEADER):
    """-> {name: (restype, [argtypes])} for every function the header declares."""
    src = re.sub(r"/\*.*?\*/", "", open(path).read(), flags=re.S)
    out = {}
    for ret, name, args in _DECL.findall(src):
        restype = {"int": ctypes.c_int, "void": None, "const char*": ctypes.c_char_p}[ret]
        argtypes = []
        args = args.strip()
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                elif a.startswith("long long"):
                    argtypes.append(ctypes.c_longlong)
                elif a.startswith("long"):
                    argtypes.append(ctypes.c_long)
                elif a.startswith("float"):
                    argtypes.append(ctypes.c_float)
                elif a.startswith("double"):
                    argtypes.append(ctypes.c_double)
                elif a.startswith("int"):
                    argtypes.append(ctypes.c_int)
                else:
                    raise ValueError(f"unetca_b200.h: cannot parse argument {a!r} of {name}")
        out[name] = (restype, argtypes)
    return out


_lib = None


def load():
    """Load the library (once) and bind every declared symbol; raises if anything is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build the CUDA extension first (python -c 'import __graft_entry__ as g; g.build()'). "
            "unetca_b200 has no CPU or PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    decls = parse_header()
    decls.update(parse_header(TUNING_HEADER))
    for name, (restype, argtypes) in decls.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise RuntimeError(f"libunetca_b200.so does not export {name} declared in unetca_b200.h") from e
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def last_error() -> str:
    return load().unetca_last_error().decode()


# kernels launched per entry point (default 1) — used for the launch count bench.py reports
_LAUNCHES = {
    "unetca_conv3x3_wgrad": 2, "unetca_im2col_wgrad": 2, "unetca_convT2x2_wgrad": 2, "unetca_chan_sum": 2,
    "unetca_se_fc_bwd": 2, "unetca_se_fc_bwd_fused": 2, "unetca_first_pairs_wgrad": 2, "unetca_outc_bwd": 2, "unetca_cross_entropy": 2,
}
launch_count = 0
_hook = None


def set_hook(fn):
    """fn(name, args) -> context manager entered around the call (bench.py times kernels with CUDA events)."""
    global _hook
    _hook = fn


def call(name: str, *args):
    """Call an int-returning entry point; negative return -> RuntimeError with the library's message."""
    global launch_count
    fn = getattr(load(), name)
    launch_count += _LAUNCHES.get(name, 1)
    if _hook is not None:
        with _hook(name, args):
            rc = fn(*args)
    else:
        rc = fn(*args)
    if rc is not None and rc < 0:
        raise RuntimeError(f"{name} failed ({rc}): {last_error()}")
    return rc
