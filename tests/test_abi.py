"""CPU tier: the C-ABI library loads and exports every symbol include/world_b200.h declares; the host-side
size helpers agree with NumPy."""
import ctypes
import os
import re

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _declared():
    src = open(os.path.join(ROOT, "include", "world_b200.h")).read()
    return sorted(set(re.findall(r"\b(wb_[a-z0-9_]+)\s*\(", src)) - {"wb_handle"})


def test_cuda_library_exports_every_declared_symbol():
    so = os.path.join(ROOT, "python-world_b200", "world_b200", "libworld_b200.so")
    assert os.path.exists(so), "build it first: python python-world_b200/build.py"
    lib = ctypes.CDLL(so)
    for name in _declared():
        assert hasattr(lib, name), name
    lib.wb_is_cuda_build.restype = ctypes.c_int
    assert lib.wb_is_cuda_build() == 1


def test_abi_table_matches_header():
    from world_b200 import _abi
    assert sorted(_abi.SIGNATURES) == _declared()


def test_size_helpers(emu):
    L = emu.L
    for n, fs, per in ((64000, 16000, 5.0), (102400, 22050, 5.0), (24000, 48000, 5.0), (16000, 16000, 1.0)):
        assert L.wb_frame_count(n, fs, per) == int(1000 * n / fs / per + 1)
    for fs, want in ((16000, 1024), (22050, 1024), (48000, 2048), (8000, 512)):
        assert L.wb_cheaptrick_fft_size(fs) == want
    for fs in (16000, 22050, 44100, 48000):
        for f in (2, 101, 801, 929):
            tp = np.arange(f) * 5 / 1000
            want = len(np.arange(tp[0], tp[-1] + 1 / fs, 1 / fs))
            assert L.wb_synthesis_length(float(tp[0]), float(tp[-1]), fs) == want
    assert L.wb_d4c_band_count(16000, 0) == 1 and L.wb_d4c_band_count(48000, 1) == 5


def test_package_refuses_without_gpu():
    """No CPU fallback: creating the engine without a CUDA device raises."""
    import pytest
    import torch
    from world_b200 import engine
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    with pytest.raises(RuntimeError):
        engine.default_engine()
