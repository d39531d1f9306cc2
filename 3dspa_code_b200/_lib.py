"""ctypes binding of lib3dspa_b200.so, generated from include/spa3d_b200.h.

The prototypes are parsed from the public header so the Python side can never drift from the
C ABI.  There is no fallback: if the shared library is missing the import fails loudly.
"""
from __future__ import annotations

import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(HERE, "..", "include", "spa3d_b200.h")
LIB_PATH = os.environ.get("SPA3D_LIB_PATH") or os.path.join(HERE, "lib3dspa_b200.so")   # override: A/B of development builds

_CTYPE = {
    "int": ctypes.c_int,
    "int64_t": ctypes.c_int64,
    "float": ctypes.c_float,
}


def parse_header(path=HEADER):
    """Return {function name: (restype, [(ctype, arg name), ...])} for every prototype."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    protos = {}
    for m in re.finditer(r"(int64_t|int|const char\*)\s+(spa3d_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        parsed = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                if "*" in a:
                    parsed.append((ctypes.c_void_p, a.split("*")[-1].strip()))
                else:
                    ty, nm = a.rsplit(" ", 1)
                    parsed.append((_CTYPE[ty.replace("const ", "").strip()], nm))
        protos[name] = (ctypes.c_char_p if "char" in ret else (ctypes.c_int64 if ret == "int64_t" else ctypes.c_int), parsed)
    return protos


PROTOTYPES = parse_header()


class Spa3dError(RuntimeError):
    pass


def load(path=LIB_PATH):
    if not os.path.isfile(path):
        raise ImportError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU or PyTorch fallback for the 3DSPA kernels)"
        )
    lib = ctypes.CDLL(path)
    for name, (restype, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = restype
        fn.argtypes = [t for t, _ in args]
    return lib


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = load()
    return _lib


def check(status, what):
    if status != 0:
        msg = lib().spa3d_last_error().decode()
        raise Spa3dError(f"{what} failed ({status}): {msg}")
