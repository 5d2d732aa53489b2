"""ctypes binding of csrc/libltn_b200.so.  Signatures are read from include/latticenet_b200.h so the
header is the single source of truth for the C ABI.  There is no CPU fallback: a missing library or
a missing CUDA device raises."""
import ctypes
import os
import re

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(HERE), "include", "latticenet_b200.h")
LIB_PATH = os.path.join(HERE, "csrc", "libltn_b200.so")

_lib = None


def declared_functions(header=HEADER):
    """[(name, [ctypes arg types])] for every `int ltn_*(...)` prototype in the header."""
    with open(header) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    out = []
    for m in re.finditer(r"\bint\s+(ltn_\w+)\s*\(([^)]*)\)\s*;", text):
        name, args = m.group(1), m.group(2).strip()
        types = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    types.append(ctypes.c_void_p)
                elif a.startswith("float"):
                    types.append(ctypes.c_float)
                elif a.startswith("int"):
                    types.append(ctypes.c_int)
                else:
                    raise RuntimeError("unsupported parameter in header: %r" % a)
        out.append((name, types))
    return out


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("CUDA extension %s is missing: run `python -m temporal_latticenet_b200.build` "
                               "(there is no CPU fallback)" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, types in declared_functions():
            fn = getattr(lib, name)  # AttributeError if the header declares something the .so lacks
            fn.restype = ctypes.c_int
            fn.argtypes = types
        _lib = lib
    return _lib


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("temporal_latticenet_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("expected a CUDA tensor")
    if not t.is_contiguous():
        raise RuntimeError("expected a contiguous tensor")
    return ctypes.c_void_p(t.data_ptr())


def check(rc, what):
    if rc != 0:
        raise RuntimeError("%s failed with code %d" % (what, rc))
