"""CheapTrick spectral envelope -- oracle restatement of world/cheaptrick.py.

TEST INFRASTRUCTURE (see oracle/__init__.py).  All frames are processed at once
as [frame, bin] matrices; the reference loops frame by frame.
"""
import numpy as np

from . import common as C


def default_fft_size(fs):
    """cheaptrick.py:20-22."""
    return int(2 ** np.ceil(np.log2(3 * fs / 71 + 1)))


def effective_f0(f0, vuv, fs, fft_size):
    """The F0 CheapTrick actually analyses with -- and leaves behind in
    source['f0'] (cheaptrick.py:24-33): unvoiced -> 500, below 3 fs/(N-3) -> 500."""
    limit = fs * 3.0 / (fft_size - 3.0)
    f = np.where(np.asarray(vuv) == 0, 500.0, np.asarray(f0, dtype=np.float64))
    return np.where(f < limit, 500.0, f)


def cheaptrick(x, fs, temporal_positions, f0, vuv, q1=-0.15, fft_size=None, dither="legacy"):
    """Returns dict(spectrogram [N/2+1, F], ps_spectrogram [N, F] complex, f0 [F] as
    left in source['f0']).  dither: 'legacy' draws |rand|*eps from np.random in the
    reference's order (cheaptrick.py:117); an ndarray [F, N/2+1] is added as is;
    None adds nothing."""
    x = np.asarray(x, dtype=np.float64)
    N = int(fft_size) if fft_size is not None else default_fft_size(fs)
    tp = np.asarray(temporal_positions, dtype=np.float64)
    f = effective_f0(f0, vuv, fs, N)
    F = len(f)

    # step 1: 3*T0 Hann window, unit energy, weighted mean removed (cheaptrick.py:79-99)
    seg, win, mask, half = C.pitch_windows(x, fs, f, tp, 1.5, "hann", subsample=False)
    win = win / np.sqrt(np.sum(win ** 2, axis=1, keepdims=True))
    wav = C.remove_weighted_mean(seg, win, half)
    ps = np.fft.fft(wav, N, axis=1)                         # cheaptrick.py:65
    power = C.mirror_low_band(np.abs(ps) ** 2, fs, f, "one_bin")   # :66-74

    # step 2: box smoothing of width 2 f0/3 (cheaptrick.py:103-118)
    smooth = C.box_integral(power, fs, f / 3) * 1.5 / f[:, None]
    if isinstance(dither, str) and dither == "legacy":
        smooth = smooth + np.abs(np.random.rand(F, N // 2 + 1)) * C.EPS
    elif dither is not None:
        smooth = smooth + dither

    # step 3: quefrency-domain lifter (cheaptrick.py:136-157)
    sym = np.concatenate([smooth, smooth[:, -2:0:-1]], axis=1)
    q = np.arange(N) / fs
    arg = np.pi * f[:, None] * q[None, 1:]
    lift_s = np.concatenate([np.ones((F, 1)), np.sin(arg) / arg], axis=1)
    lift_c = (1 - 2 * q1) + 2 * q1 * np.cos(2 * np.pi * q[None, :] * f[:, None])
    for lift in (lift_s, lift_c):
        lift[:, N // 2 + 1:] = lift[:, N // 2 - 1:0:-1]
    cep = np.fft.fft(np.log(sym), axis=1)
    env = np.exp(np.real(np.fft.ifft(cep * lift_s * lift_c, axis=1)))[:, :N // 2 + 1]
    return {"spectrogram": env.T.copy(), "ps_spectrogram": ps.T.copy(), "f0": f}
