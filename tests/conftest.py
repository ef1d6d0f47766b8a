import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (ROOT, os.path.join(ROOT, "python-world_b200"), os.path.join(ROOT, "tests", "hostemu"),
          os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")
EPS = 2.220446049250313e-16


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def syn16k():
    return dict(np.load(os.path.join(GOLDEN, "syn16k_1s.npz")))


@pytest.fixture(scope="session")
def syn48k():
    return dict(np.load(os.path.join(GOLDEN, "syn48k_05s.npz")))


@pytest.fixture(scope="session")
def mwm():
    g = dict(np.load(os.path.join(GOLDEN, "mwm_full.npz")))
    g["x"] = g["x_int16"] / 32767.0  # test/speed.py:14
    return g


@pytest.fixture(scope="session")
def emu():
    import emu as emu_mod
    e = emu_mod.Emu()
    yield e
    e.close()


@pytest.fixture(scope="session")
def engine():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from world_b200 import engine as eng
    return eng.default_engine()


def legacy_dither(n_frames, bins, seed=0):
    """|rand|*eps in the order the reference draws it (cheaptrick.py:117) after np.random.seed(seed)."""
    np.random.seed(seed)
    return np.abs(np.random.rand(n_frames, bins)) * EPS


def spec_close(got, want, floor=1e-10):
    """SURVEY 8d tolerance: |d log10| p99 <= 1e-4 and max <= 1e-3 over bins with power > floor."""
    m = want > floor
    d = np.abs(np.log10(got[m]) - np.log10(want[m]))
    return float(np.percentile(d, 99)) if d.size else 0.0, float(d.max()) if d.size else 0.0
