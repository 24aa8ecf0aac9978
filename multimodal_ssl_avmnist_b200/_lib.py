"""ctypes binding of libavmnist_b200.so (the C ABI declared in include/avmnist_b200.h).

There is deliberately no fallback: if the library is missing or a call fails, an exception is raised.
"""
import ctypes as C
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libavmnist_b200.so")
HEADER = os.path.join(os.path.dirname(HERE), "include", "avmnist_b200.h")
ABI_VERSION = 1

_lib = None

_CTYPES = {
    "int": C.c_int, "int64_t": C.c_int64, "uint64_t": C.c_uint64, "float": C.c_float, "double": C.c_double,
    "const char*": C.c_char_p,
}


class B200Error(RuntimeError):
    pass


def declared_functions(header=HEADER):
    """[(name, return_type, [arg types])] parsed from the C header (single source of truth for the signatures)."""
    src = open(header).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = []
    for m in re.finditer(r"^\s*(const char\*|int64_t|int)\s+(b200_\w+)\s*\(([^)]*)\)\s*;", src, flags=re.M):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        types = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                types.append("ptr" if "*" in a else a.rsplit(" ", 1)[0].replace("const ", ""))
        out.append((name, ret, types))
    return out


def load():
    """Load the shared library once and attach argtypes/restype from the header."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200Error(f"{LIB_PATH} not found: build it with `python -m multimodal_ssl_avmnist_b200.build` "
                        "(there is no CPU / PyTorch fallback for the hot path)")
    lib = C.CDLL(LIB_PATH)
    for name, ret, types in declared_functions():
        fn = getattr(lib, name)          # AttributeError here == header/library mismatch
        fn.restype = _CTYPES[ret]
        fn.argtypes = [C.c_void_p if t == "ptr" else _CTYPES[t] for t in types]
    if lib.b200_abi_version() != ABI_VERSION:
        raise B200Error("libavmnist_b200.so ABI version mismatch: rebuild the library")
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().b200_last_error().decode("utf-8", "replace")
        raise B200Error(f"{what} failed (rc={rc}): {msg}")
