"""ctypes binding of libha2g_b200.so (the C-ABI drop-in boundary, include/ha2g_b200.h).

The prototypes are parsed from the public header so Python argtypes cannot drift from the C side.
There is NO fallback: if the shared library is missing or a launch fails, we raise.
"""
from __future__ import annotations

import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libha2g_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "ha2g_b200.h")

_CTYPE = {
    "int": ctypes.c_int, "float": ctypes.c_float, "double": ctypes.c_double, "int64_t": ctypes.c_int64,
    "uint64_t": ctypes.c_uint64, "uint32_t": ctypes.c_uint32, "cudaStream_t": ctypes.c_void_p, "size_t": ctypes.c_size_t,
}
_PROTO = re.compile(r"^int\s+(ha2g_\w+)\s*\(([^)]*)\)\s*;", re.M)


def parse_header(path: str = HEADER_PATH):
    """-> {name: [ctypes types]} for every `int ha2g_*(...)` declaration."""
    text = open(path).read()
    protos = {}
    for name, args in _PROTO.findall(text):
        types = []
        for a in [x.strip() for x in args.split(",") if x.strip()]:
            if "*" in a:
                types.append(ctypes.c_void_p)
            else:
                toks = a.replace("const", "").split()
                types.append(_CTYPE[toks[0]])
        protos[name] = types
    return protos


class Ha2gError(RuntimeError):
    pass


class _Lib:
    def __init__(self):
        self._dll = None
        self._fns = {}

    def load(self):
        if self._dll is not None:
            return self
        if not os.path.exists(LIB_PATH):
            raise Ha2gError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). ha2g_b200 has no CPU or PyTorch fallback path.")
        self._dll = ctypes.CDLL(LIB_PATH)
        for name, types in parse_header().items():
            fn = getattr(self._dll, name)  # AttributeError => header/library mismatch
            fn.argtypes = types
            fn.restype = ctypes.c_int
            self._fns[name] = fn
        return self

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        self.load()
        fn = self._fns[name]

        def call(*args):
            rc = fn(*args)
            if rc != 0:
                raise Ha2gError(f"{name} failed with cudaError {rc}")
        call.__name__ = name
        self.__dict__[name] = call
        return call


lib = _Lib()
