"""In-memory stand-in for ``h5py`` so the UNMODIFIED reference loader can run here.

TEST INFRASTRUCTURE ONLY (used by oracle/gen_golden.py and tests that pin the oracle
against the live reference).  ``h5py`` is not installed in this image; the reference
only touches it inside function bodies (src/lofar_tools.py:76,227,360,415,442), always
as ``h5py.File(name,'r')[...]`` walks over groups, which nested dicts of numpy arrays
satisfy.
"""
import sys
import types

_REGISTRY = {}


def register(name: str, measurement: dict) -> None:
    _REGISTRY[name] = measurement


def _file(name, mode="r"):
    return _REGISTRY[name]


def install() -> None:
    mod = types.ModuleType("h5py")
    mod.File = _file
    sys.modules["h5py"] = mod
