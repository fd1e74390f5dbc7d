"""ctypes binding of the C-ABI shared library (include/univl_b200.h).

The library is the product's only compute backend: there is no CPU or PyTorch fallback.  `load()` raises if
the library is missing, and every wrapper raises `RuntimeError(univl_last_error_string())` on a non-zero
return code (SURVEY.md §8b error convention).
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libunivl_b200.so")

c_int = ctypes.c_int
c_ll = ctypes.c_longlong
c_ull = ctypes.c_ulonglong
c_f = ctypes.c_float
c_p = ctypes.c_void_p

HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "univl_b200.h")


def _ctype_of(decl):
    """Map one C parameter declaration of include/univl_b200.h to a ctypes type."""
    decl = decl.strip()
    if "*" in decl:
        return c_p
    base = " ".join(decl.split()[:-1]) if len(decl.split()) > 1 else decl
    base = base.replace("const", "").strip()
    return {"int": c_int, "float": c_f, "long long": c_ll, "unsigned long long": c_ull}[base]


def parse_header(path=HEADER_PATH):
    """{name: (restype, [argtypes])} for every function the header declares — the single source of truth for the
    binding, so the ctypes signatures cannot drift from the C ABI."""
    import re
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    out = {}
    for m in re.finditer(r"(const char\*|int)\s+(univl_\w+)\s*\(([^)]*)\)\s*;", text):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        argtypes = [] if args in ("", "void") else [_ctype_of(a) for a in args.split(",")]
        out[name] = (ctypes.c_char_p if "char" in ret else c_int, argtypes)
    return out


_lib = None
_lock = threading.Lock()


def load(build_if_missing=False):
    """Load libunivl_b200.so and attach argtypes to every exported entry point that exists."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            if build_if_missing:
                from . import build as _build
                _build.build()
            else:
                raise RuntimeError(
                    "univl_b200: %s is missing — run `python -m univl_b200.build` (or __graft_entry__.build()); "
                    "there is no fallback compute path" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in parse_header().items():
            fn = getattr(lib, name, None)
            if fn is None:
                raise RuntimeError("univl_b200: %s does not export %s (stale build?)" % (LIB_PATH, name))
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def call(name, *args):
    """Invoke a C-ABI entry point; raise on a non-zero status."""
    lib = load()
    fn = getattr(lib, name, None)
    if fn is None:
        raise RuntimeError("univl_b200: entry point %s is not exported by %s" % (name, LIB_PATH))
    rc = fn(*args)
    if rc != 0:
        raise RuntimeError("%s failed (%d): %s" % (name, rc, lib.univl_last_error_string().decode()))
    return rc
