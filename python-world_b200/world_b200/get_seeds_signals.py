"""Band-limited pulse and velvet-noise seeds for the requiem synthesiser (world/get_seeds_signals.py:8).

Host-side set-up (a few milliseconds, once per decode): the seeds are defined by draws from the legacy
`random` / `np.random` generators, made here in the reference's order so that a seeded decode reproduces
the reference's seeds; the synthesis itself runs on the GPU.
"""
import random

import numpy as np


def _round_up_half(v):
    """The reference's round_matlab: v + 0.5 for v > 0 (no truncation)."""
    v = np.asarray(v, dtype=np.float64)
    return np.where(v > 0, v + 0.5, v - 0.5)


def generate_short_velvet_noise(N):
    """get_seeds_signals.py:57-73."""
    cells = int(N // 4 + 0.5)
    signs = 2.0 * np.ones(cells)
    signs[int(cells // 2):] *= -1
    for i in range(cells):
        j = random.randint(0, cells - 1)
        signs[j], signs[i] = signs[i], signs[j]
    n = np.zeros(N)
    n[4 * np.arange(cells) + np.random.randint(4, size=cells)] = signs
    return n


def generate_modified_velvet_noise(N, fs):
    """get_seeds_signals.py:40-54."""
    periods = 8 * _round_up_half(np.array([8, 30, 60]) * fs / 48000)
    n = np.zeros(N + int(np.max(periods)) + 1)
    at = 0
    while True:
        length = int(periods[random.randint(0, len(periods) - 1)])
        n[at:at + length] = generate_short_velvet_noise(length)
        at += length
        if at >= N - 1:
            break
    return n[:N]


def get_seeds_signals(fs, fft_size=None, noise_length=None):
    if fft_size is None:
        fft_size = int(1024 * (2 ** np.ceil(np.log2(fs / 48000))))
    if noise_length is None:
        noise_length = int(2 ** np.ceil(np.log2(fs / 2)))
    w = np.arange(fft_size // 2 + 1) * fs / fft_size
    interval = 3000
    bands = int(2 + np.floor(min(15000, fs / 2 - interval) / interval))
    pulse = np.zeros((fft_size, bands))
    noise = np.zeros((noise_length, bands))
    spec_n = np.fft.fft(generate_modified_velvet_noise(noise_length, fs), noise_length)
    for i in range(bands):
        shape = 0.5 + 0.5 * np.cos(((w - interval * i) / (interval * 2)) * 2 * np.pi)
        shape[w > interval * (i + 1)] = 0
        shape[w < interval * (i - 1)] = 0
        if i == bands - 1:
            shape[w > interval * i] = 1
        pulse[:, i] = np.fft.fftshift(np.fft.ifft(np.r_[shape, shape[-2:0:-1]]).real)
        noise[:, i] = np.fft.ifft(spec_n * np.fft.fft(pulse[:, i], noise_length)).real
    n = np.arange(1, fft_size + 1)
    h = 0.5 - 0.5 * np.cos(2 * np.pi * n / (fft_size + 1))  # hanning(fft_size + 2)[1:-1]
    pulse[:, 0] = pulse[:, 0] - np.mean(pulse[:, 0]) * h / np.mean(h)
    return {'pulse': pulse, 'noise': noise}
