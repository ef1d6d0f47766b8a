"""Loader of the CUDA C-ABI library (libworld_b200.so, built in-tree by build.py)."""
import ctypes
import os

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
# WORLD_B200_LIB: another build of the same library (tuning / test variants made with build.py --out)
SO_PATH = os.environ.get("WORLD_B200_LIB") or os.path.join(_HERE, "libworld_b200.so")
_lib = None


def load():
    """Load and declare the library.  Fails loudly when it is missing or is not a CUDA build."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise RuntimeError(
            "world_b200: %s not found -- build it with `python python-world_b200/build.py` "
            "(there is no CPU fallback)" % SO_PATH)
    lib = _abi.declare(ctypes.CDLL(SO_PATH))
    if lib.wb_is_cuda_build() != 1:
        raise RuntimeError("world_b200: %s is not a CUDA build" % SO_PATH)
    _lib = lib
    return lib
