"""Deterministic "speech-shaped" synthetic utterances.

Host-side input generation only: used by bench.py, tests/ and
tests/golden/make_golden.py so that every arm (CUDA path, oracle, reference)
sees the same samples.  There is no network for datasets, hence synthetic.

  seed          = 1000 * config + utterance_index   (numpy.random.default_rng)
  F0 contour    = base * 2^(0.3 sin(2 pi 0.7 t + phi1) + 0.05 sin(2 pi 5 t)),
                  base log-uniform in [90, 250] Hz
  voiced mask   = sign of a 1.1 Hz sinusoid with an offset (about 65-70 % voiced)
  source        = sum_k cos(k phi)/k (band-limited) when voiced, 0.3 N(0,1) otherwise
  vocal tract   = three 2-pole resonators 700/1200/2600 Hz, bw 130/150/200 Hz
  level         = peak-normalised to 0.5 plus a 1e-3 N(0,1) floor, float64
"""
import numpy as np
from scipy.signal import lfilter

_FORMANTS = ((700.0, 130.0), (1200.0, 150.0), (2600.0, 200.0))


def utterance(fs: int, seconds: float, config: int, index: int) -> np.ndarray:
    rng = np.random.default_rng(1000 * int(config) + int(index))
    n = int(round(fs * seconds))
    t = np.arange(n) / fs
    base = 90.0 * (250.0 / 90.0) ** rng.random()
    phi1, phi2 = rng.random(2) * 2.0 * np.pi
    f0 = base * 2.0 ** (0.3 * np.sin(2 * np.pi * 0.7 * t + phi1) + 0.05 * np.sin(2 * np.pi * 5.0 * t))
    voiced = np.sin(2 * np.pi * 1.1 * t + phi2) > -0.5
    phase = 2.0 * np.pi * np.cumsum(f0) / fs
    n_harm = int(fs / 2.0 / (base * 2.0 ** 0.35))
    src = np.zeros(n)
    for k in range(1, max(1, n_harm) + 1):
        src += np.cos(k * phase) / k
    src = np.where(voiced, src, 0.3 * rng.standard_normal(n))
    y = np.zeros(n)
    for fc, bw in _FORMANTS:
        if fc >= fs / 2:
            continue
        r = np.exp(-np.pi * bw / fs)
        a = [1.0, -2.0 * r * np.cos(2 * np.pi * fc / fs), r * r]
        y += lfilter([1.0 - r], a, src)
    y *= 0.5 / np.max(np.abs(y))
    y += 1e-3 * rng.standard_normal(n)
    return y


def batch(fs: int, seconds: float, config: int, count: int, first: int = 0) -> np.ndarray:
    """[count, n] float64, utterance indices first .. first+count-1."""
    return np.stack([utterance(fs, seconds, config, first + i) for i in range(count)])
