"""ctypes binding of csrc/libltn_b200.so.  Signatures are read from include/latticenet_b200.h so the
header is the single source of truth for the C ABI.  There is no CPU fallback: a missing library or
a missing CUDA device raises."""
import ctypes
import os
import re
import threading

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
    """raw cudaStream_t of torch's current stream (the C accessor: torch.cuda.current_stream() builds a
    Python Stream object per call, ~15 us, which adds up over the ~250 launches of a frame)"""
    return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(torch.cuda.current_device()))


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


# ---------------------------------------------------------------------------------------------------
# Static-capacity mode (temporal_latticenet_b200/engine.py): every tensor's row count is a fixed
# CAPACITY and the true count lives in device memory, so a whole frame can be captured in a CUDA graph.
# The registry maps a capacity (= shape[0] of the tensors of that row class: points, rows, vertices
# of each lattice level) to the device int32 holding the live count; kernel wrappers look their row
# counts up here and pass the pointer as the `*_dev` companion of the host bound.
# ---------------------------------------------------------------------------------------------------
# Per THREAD: the lock-step engine runs one host thread per window in flight while it captures a frame (engine.py).
_static = threading.local()


def _rows():
    return getattr(_static, "rows", None) or {}


def set_static_rows(mapping):
    _static.rows = dict(mapping or {})


def static_mode():
    return bool(_rows())


def rows_tensor(nr_rows):
    return _rows().get(int(nr_rows))


def rows_dev(nr_rows):
    t = _rows().get(int(nr_rows))
    return None if t is None else ctypes.c_void_p(t.data_ptr())
