"""Waveform synthesis -- oracle restatement of world/synthesis.py, world/synthesisRequiem.py and
world/get_seeds_signals.py.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Pulses (synthesis) and frames (requiem) are processed as [item, bin] matrices; the reference loops.
All random draws are made in the reference's order from the same legacy generators (np.random,
random), so a seeded run reproduces the reference's noise stream.
"""
import random

import numpy as np
from scipy.signal.windows import hann

from . import common as C


def time_axis(tp, fs):
    """synthesis.py:39 / synthesisRequiem.py:39."""
    return np.arange(tp[0], tp[-1] + 1 / fs, 1 / fs)


def pulse_train(tp, f0, vuv, fs, t):
    """time_base_generation (synthesis.py:120-140, synthesisRequiem.py:104-118): sample-rate F0 (500 Hz
    where unvoiced), phase accumulation, one pulse per 2*pi wrap.  Returns pulse times, 1-based pulse
    sample indices, fractional delays (s) and the interpolated voicing flag."""
    f0_i = C.lerp_extrap(tp, f0, t)
    v_i = C.lerp_extrap(tp, vuv, t) > 0.5
    f0_i = f0_i * v_i
    f0_i[f0_i == 0] = f0_i[f0_i == 0] + 500
    wrap = np.remainder(np.cumsum(2 * np.pi * f0_i / fs), 2 * np.pi)
    loc = t[:-1][np.abs(np.diff(wrap)) > np.pi]
    idx = C.round_half_up(loc * fs) + 1
    y1 = wrap[idx - 1] - 2.0 * np.pi
    y2 = wrap[idx]
    return loc, idx, (-y1 / (y2 - y1)) / fs, v_i


def _anticausal_cepstrum(log_half_spectrum_rows, N):
    """Cepstrum of log|S|/2 folded onto index 0 and the upper half (synthesis.py:88-91, 106-109)."""
    cep = np.fft.fft(log_half_spectrum_rows, axis=1).real
    folded = np.zeros_like(cep)
    folded[:, N // 2:] = cep[:, N // 2:] * 2
    folded[:, 0] = cep[:, 0]
    return folded


def _sym(rows):
    return np.concatenate([rows, rows[:, -2:0:-1]], axis=1)


def synthesis(dat, noise=None):
    """synthesis.py:21-82.  `noise`: optional pre-drawn standard normals (consumed pulse by pulse, each
    pulse taking max(3, noise_size)); default draws np.random.randn in the reference's order."""
    f0, vuv, fs = np.asarray(dat["f0"], float), np.asarray(dat["vuv"], float), dat["fs"]
    spec, ap, tp = dat["spectrogram"], dat["aperiodicity"], np.asarray(dat["temporal_positions"], float)
    t = time_axis(tp, fs)
    y = np.zeros(len(t))
    loc, idx, shift, v_i = pulse_train(tp, f0, vuv, fs, t)
    assert len(loc) > 0
    N = (spec.shape[0] - 1) * 2
    base = np.arange(-N // 2 + 1, N // 2 + 1)
    F = len(tp)
    pos = np.clip(C.lerp_extrap(tp, np.arange(1, F + 1), loc), 1, F)
    lo = np.floor(pos).astype(int) - 1
    hi = np.ceil(pos).astype(int) - 1
    t1, t2 = tp[lo], tp[hi]
    xq = np.maximum(t1, np.minimum(t2, loc))
    same = t1 == t2
    b = np.where(same, 0.0, (xq - t1) / np.where(same, 1.0, t2 - t1))
    a = 1 - b
    amp_ap = ap ** 2
    amp_p = np.maximum(0.001, 1 - amp_ap)

    def slice_of(m):
        return np.where(same[:, None], m[:, lo].T, a[:, None] * m[:, lo].T + b[:, None] * m[:, hi].T)

    s_sl, p_sl, a_sl = slice_of(spec), slice_of(amp_p), slice_of(amp_ap)
    nxt = idx[np.minimum(len(idx) - 1, np.arange(len(idx)) + 1)]
    noise_size = nxt - idx
    voiced = (v_i[idx - 1] >= 0.5) & (a_sl[:, 0] <= 0.999)

    # periodic part: minimum-phase response with a fractional delay, DC removed (synthesis.py:100-116, 57-74)
    ps = s_sl * p_sl
    ps[ps == 0] = C.EPS
    half = np.exp(np.fft.ifft(_anticausal_cepstrum(np.log(np.abs(_sym(ps))) / 2, N), axis=1))[:, :N // 2 + 1]
    half = half * np.exp(-1j * (2.0 * np.pi * fs / N) * shift[:, None] * np.arange(N // 2 + 1)[None, :])
    full = np.concatenate([half, half[:, -2:0:-1].conj()], axis=1)
    per = np.fft.fftshift(np.fft.ifft(full, axis=1).real, axes=1)
    dc = hann(N + 2)[1:-1]
    dc = dc / np.sum(dc)
    per = (per + dc[None, :] * -np.sum(per, axis=1, keepdims=True)) * np.sqrt(np.maximum(1, noise_size))[:, None]

    # aperiodic part: minimum-phase response of S*A (S alone when unvoiced) driven by zero-mean noise
    aps = np.where(voiced[:, None], s_sl * a_sl, s_sl)
    aps[aps == 0] = C.EPS
    resp = np.exp(np.fft.ifft(_anticausal_cepstrum(np.log(np.abs(_sym(aps))) / 2, N), axis=1))
    resp = np.fft.fftshift(np.fft.ifft(resp, axis=1).real, axes=1)
    sizes = np.maximum(3, noise_size)
    if noise is None:
        noise = np.concatenate([np.random.randn(int(s)) for s in sizes])
    cuts = np.concatenate([[0], np.cumsum(sizes)]).astype(int)

    for i in range(len(idx)):
        out = np.clip(idx[i] + base, 1, len(y)) - 1
        if voiced[i]:
            y[out] += per[i]          # duplicate (clamped) indices: last write wins, as in the reference
        nz = noise[cuts[i]:cuts[i + 1]]
        y[out] += np.convolve(nz - np.mean(nz), resp[i])[:N]   # fftfilt(b, x) == conv truncated to len(x)
    return y


def draw_noise(dat):
    """The normals synthesis() consumes, in order (synthesis.py:93): for the product path, which takes the
    stream as an input."""
    tp = np.asarray(dat["temporal_positions"], float)
    t = time_axis(tp, dat["fs"])
    _, idx, _, _ = pulse_train(tp, np.asarray(dat["f0"], float), np.asarray(dat["vuv"], float), dat["fs"], t)
    nxt = idx[np.minimum(len(idx) - 1, np.arange(len(idx)) + 1)]
    return np.random.randn(int(np.sum(np.maximum(3, nxt - idx))))


# ------------------------------------------------------------------------------------------- requiem
def _short_velvet(n):
    """generate_short_velvet_noise (get_seeds_signals.py:57-73): +-2 impulses, one per 4-sample cell,
    signs shuffled with random.randint, positions from np.random.randint."""
    out = np.zeros(n)
    cells = int(n // 4 + 0.5)
    signs = np.ones(cells)
    signs[int(cells // 2):] *= -1
    signs *= 2
    for i in range(cells):
        j = random.randint(0, cells - 1)
        signs[j], signs[i] = signs[i], signs[j]
    out[4 * np.arange(cells) + np.random.randint(4, size=cells)] = signs
    return out


def _velvet(n, fs):
    """generate_modified_velvet_noise (get_seeds_signals.py:40-54)."""
    periods = (8 * C.half_away(np.array([8, 30, 60]) * fs / 48000)).astype(float)
    buf = np.zeros(n + int(np.max(periods)) + 1)
    at = 0
    while True:
        p = int(periods[random.randint(0, len(periods) - 1)])
        buf[at:at + p] = _short_velvet(p)
        at += p
        if at >= n - 1:
            break
    return buf[:n]


def seeds(fs, fft_size=None, noise_length=None):
    """get_seeds_signals (get_seeds_signals.py:8-38): raised-cosine band filters 3 kHz apart as
    zero-phase pulses, and velvet noise filtered by each band."""
    if fft_size is None:
        fft_size = int(1024 * (2 ** np.ceil(np.log2(fs / 48000))))
    if noise_length is None:
        noise_length = int(2 ** np.ceil(np.log2(fs / 2)))
    w = np.arange(fft_size // 2 + 1) * fs / fft_size
    step = 3000
    n_b = int(2 + np.floor(min(15000, fs / 2 - step) / step))
    pulse = np.zeros((fft_size, n_b))
    noise = np.zeros((noise_length, n_b))
    spec_n = np.fft.fft(_velvet(noise_length, fs), noise_length)
    for i in range(n_b):
        sp = 0.5 + 0.5 * np.cos(((w - step * i) / (step * 2)) * 2 * np.pi)
        sp[w > step * (i + 1)] = 0
        sp[w < step * (i - 1)] = 0
        if i == n_b - 1:
            sp[w > step * i] = 1
        pulse[:, i] = np.fft.fftshift(np.fft.ifft(np.r_[sp, sp[-2:0:-1]]).real)
        noise[:, i] = np.fft.ifft(spec_n * np.fft.fft(pulse[:, i], noise_length)).real
    h = hann(fft_size + 2)[1:-1]
    pulse[:, 0] = pulse[:, 0] - np.mean(pulse[:, 0]) * h / np.mean(h)
    return {"pulse": pulse, "noise": noise}


def excitation(dat, seed, cursor=None):
    """get_excitation_signal (synthesisRequiem.py:27-63).  `cursor` [bands] is the cyclic read position
    of each noise band (generate_noise.current_index, :131-141); returns (signal, new cursor)."""
    tp = np.asarray(dat["temporal_positions"], float)
    f0, vuv, fs = np.asarray(dat["f0"], float), np.asarray(dat["vuv"], float), dat["fs"]
    pulse_seed, noise_seed, band_ap = seed["pulse"], seed["noise"], dat["aperiodicity"]
    n_fft, n_b = pulse_seed.shape
    base = np.arange(-n_fft // 2 + 1, n_fft // 2 + 1)
    t = time_axis(tp, fs)
    _, idx, _, v_i = pulse_train(tp, f0, vuv, fs, t)
    ap_i = np.stack([C.lerp_extrap(tp, 10 ** (band_ap[b] / 10), t) for b in range(band_ap.shape[0])])
    cursor = np.zeros(n_b) if cursor is None else np.array(cursor, dtype=float)
    aper = np.zeros(len(t))
    n_len = noise_seed.shape[0]
    for b in range(n_b):
        k = np.remainder(np.arange(cursor[b], cursor[b] + len(t)), n_len).astype(int)
        aper += noise_seed[k, b] * ap_i[b, :len(t)]
        cursor[b] = k[-1]
    per = np.zeros(len(t))
    for i in range(len(idx)):
        if (v_i[idx[i] - 1] <= 0.5) or (ap_i[0, idx[i] - 1] > 0.999):
            continue
        gain = np.sqrt(max(1, idx[min(len(idx) - 1, i + 1)] - idx[i]))
        out = np.clip(idx[i] + base, 1, len(t)) - 1
        per[out] += (pulse_seed * (1 - ap_i[:, idx[i] - 1])[None, :]).sum(axis=1) * gain
    return per + aper, cursor


def requiem_waveform(exc, spec, tp, fs):
    """get_waveform (synthesisRequiem.py:74-101): every frame's Hann-windowed excitation is filtered with
    the minimum-phase response of that frame's envelope and overlap-added."""
    y = np.zeros(len(exc))
    N = (spec.shape[0] - 1) * 2
    hop = int((tp[1] - tp[0]) * fs)
    wl = hop * 2 - 1
    win = hann(wl + 2)[1:-1]
    frames = np.arange(2, spec.shape[1] - 1)
    if len(frames) == 0:
        return y
    origin = (frames - 1) * hop - (hop - 1)
    seg_idx = np.minimum(len(y), origin[:, None] + np.arange(wl)[None, :])
    tmp = exc[seg_idx - 1] * win[None, :]
    env = spec[:, frames - 1].T
    mp = np.exp(np.fft.ifft(_anticausal_cepstrum(np.log(np.abs(_sym(env))) / 2, N), axis=1))
    resp = np.fft.ifft(mp * np.fft.fft(tmp, N, axis=1), axis=1).real
    for k in range(len(frames)):
        out = np.minimum(len(y), np.arange(origin[k], origin[k] + N)) - 1
        y[out] += resp[k]
    return y


def synthesis_requiem(dat, seed, cursor=None):
    """synthesisRequiem.py:12-25.  Returns (y, new noise cursor)."""
    exc, cursor = excitation(dat, seed, cursor)
    tp = np.asarray(dat["temporal_positions"], float)
    return requiem_waveform(exc, dat["spectrogram"], tp, dat["fs"]), cursor


def decode(dat, cursor=None):
    """main.py:198-214 (without mutating dat): returns (out, new cursor)."""
    if dat["is_requiem"]:
        y, cursor = synthesis_requiem(dat, seeds(dat["fs"]), cursor)
    else:
        y = synthesis(dat)
    m = np.max(np.abs(y))
    if m > 1.0:
        y = y / m
    return y, cursor
