"""Helpers for the single-utterance (NumPy in / NumPy out) stage functions."""
import numpy as np
import torch

from . import engine as _engine


def eng():
    return _engine.default_engine()


def dev1(E, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    return E.f64(x[None]), E.i32([len(x)])


def frames1(E, *arrays):
    return [E.f64(np.ascontiguousarray(a, dtype=np.float64)[None]) for a in arrays]


def ref_matrix(t):
    """[F, bins] device tensor -> NumPy [bins, F], C-contiguous like the reference's arrays."""
    return np.ascontiguousarray(t.cpu().numpy().T)


def dev_matrix(E, m):
    """NumPy [bins, F] -> device [1, F, bins]."""
    return E.f64(np.ascontiguousarray(np.asarray(m, dtype=np.float64).T)[None])
