"""ctypes binding of liblshm_sm100.so (the C ABI declared in include/lshm.h).

The prototypes are parsed from the header itself, so the binding cannot drift from the
declared ABI.  There is NO fallback: if the shared library is missing or fails to load,
importing any compute entry point raises.
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List, Tuple

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(_HERE), "include", "lshm.h")
LIBRARY = os.environ.get("LSHM_LIBRARY") or os.path.join(_HERE, "liblshm_sm100.so")   # (override: A/B runs of two builds)

_CTYPES = {
    "int": ctypes.c_int,
    "int64_t": ctypes.c_int64,
    "float": ctypes.c_float,
    "double": ctypes.c_double,
    "lshm_stream_t": ctypes.c_void_p,
}


class LshmError(RuntimeError):
    """Raised when a library call returns a non-zero status (reference convention:
    Python exceptions / asserts, src/lofar_tools.py:69-70)."""


def parse_header(path: str = HEADER) -> Dict[str, Tuple[str, List[Tuple[str, str]]]]:
    """Return {name: (return type, [(ctype string, arg name), ...])} for every LSHM_API."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"LSHM_API\s+([\w\s\*]+?)\s*(lshm_\w+)\s*\(([^)]*)\)\s*;", text):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        parsed = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                mm = re.match(r"(.*?)(\w+)$", a)
                parsed.append((mm.group(1).strip(), mm.group(2)))
        protos[name] = (ret, parsed)
    return protos


def _to_ctype(t: str):
    if "*" in t:
        return ctypes.c_void_p
    t = t.replace("const", "").strip()
    return _CTYPES[t]


class _Library:
    def __init__(self):
        if not os.path.exists(LIBRARY):
            raise LshmError(
                f"{LIBRARY} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C lshm_b200/csrc`). There is no CPU fallback.")
        self.cdll = ctypes.CDLL(LIBRARY)
        self.protos = parse_header()
        self.launches = 0  # counted kernel-launching calls (bench.py's gpu_launches claim)
        self.cdll.lshm_last_error.restype = ctypes.c_char_p
        for name, (ret, args) in self.protos.items():
            fn = getattr(self.cdll, name)
            fn.argtypes = [_to_ctype(t) for t, _ in args]
            fn.restype = ctypes.c_char_p if "char" in ret else ctypes.c_int
            if name in ("lshm_last_error", "lshm_version", "lshm_device_info", "lshm_conv_image_bytes"):
                continue
            setattr(self, name[len("lshm_"):], self._wrap(name, fn))

    def _wrap(self, name, fn):
        host_only = name in ("lshm_conv_prep_record",)

        def call(*args):
            rc = fn(*args)
            if not host_only:
                self.launches += 1
            if rc != 0:
                raise LshmError(f"{name} failed ({rc}): {self.cdll.lshm_last_error().decode()}")
        call.__name__ = name
        return call

    def conv_image_bytes(self, dim: int, A: int, Bc: int, which: int) -> int:
        n = ctypes.c_int64()
        rc = self.cdll.lshm_conv_image_bytes(dim, A, Bc, which, ctypes.byref(n))
        if rc != 0:
            raise LshmError(f"lshm_conv_image_bytes failed: {self.cdll.lshm_last_error().decode()}")
        return int(n.value)

    def version(self) -> int:
        return int(self.cdll.lshm_version())

    def device_info(self):
        sm, a, b = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        rc = self.cdll.lshm_device_info(ctypes.byref(sm), ctypes.byref(a), ctypes.byref(b))
        if rc != 0:
            raise LshmError(f"lshm_device_info failed: {self.cdll.lshm_last_error().decode()}")
        return sm.value, a.value, b.value


_LIB = None


def lib() -> _Library:
    global _LIB
    if _LIB is None:
        _LIB = _Library()
    return _LIB
