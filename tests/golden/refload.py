"""Import the UNMODIFIED reference package from /root/reference (build container only).

Test infrastructure: used by make_golden.py and by the optional live-reference
checks in tests/ (skipped when /root/reference is absent, e.g. on the GPU box).
Three import-time shims are needed on numpy 2 / scipy 1.18:
np.int, scipy.signal.hanning, and a stub matplotlib.  Nothing in the reference
tree is edited or copied.
"""
import importlib
import os
import random
import sys
import types

REF_ROOT = os.environ.get("WORLD_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "world"))


def load():
    """Return the reference `world` package imported under the name `refworld`."""
    if "refworld" in sys.modules:
        return sys.modules["refworld"]
    import numpy as np
    import scipy.signal
    import scipy.signal.windows

    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache_refworld")
    sys.dont_write_bytecode = True
    if not hasattr(np, "int"):
        np.int = int  # noqa: NPY001  (shim for the 2018-era reference)
    if not hasattr(scipy.signal, "hanning"):
        scipy.signal.hanning = scipy.signal.windows.hann
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.mlab"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].mlab = sys.modules["matplotlib.mlab"]

    spec = importlib.util.spec_from_file_location(
        "refworld", os.path.join(REF_ROOT, "world", "__init__.py"),
        submodule_search_locations=[os.path.join(REF_ROOT, "world")])
    pkg = importlib.util.module_from_spec(spec)
    sys.modules["refworld"] = pkg
    spec.loader.exec_module(pkg)
    for sub in ("dio", "stonemask", "harvest", "cheaptrick", "d4c", "d4cRequiem",
                "get_seeds_signals", "synthesis", "synthesisRequiem"):
        importlib.import_module("refworld." + sub)
    return pkg


def reseed(seed: int = 0):
    """Reset every RNG / hidden state the reference consumes (SURVEY.md 8c caveats 3-4)."""
    import numpy as np
    np.random.seed(seed)
    random.seed(seed)
    if "refworld.synthesisRequiem" in sys.modules:
        sys.modules["refworld.synthesisRequiem"].generate_noise.current_index = None
