"""D4C / D4C-Requiem band aperiodicity -- oracle restatement of world/d4c.py and
world/d4cRequiem.py.  TEST INFRASTRUCTURE (see oracle/__init__.py).

The two reference modules share one estimator (d4cRequiem.py:48-240 duplicates
d4c.py:68-260); here it exists once and is evaluated for all frames that pass
the love-train gate at the same time.
"""
import numpy as np

from . import common as C


def _band_count(fs, interval):
    """d4c.py:34, d4cRequiem.py:20."""
    return int(np.floor(min(15000, fs / 2 - interval) / interval))


def love_train(x, fs, f0, pos, threshold=0.85):
    """VUV gate on the <4 kHz / <7.9 kHz power ratio above 100 Hz (d4c.py:68-88).
    f0 == 0 frames fail outright.  Returns bool [F]."""
    f0 = np.asarray(f0, dtype=np.float64)
    ok = np.zeros(len(f0), dtype=bool)
    cand = np.nonzero(f0 != 0)[0]
    if len(cand) == 0:
        return ok
    N = int(2 ** np.ceil(np.log2(3 * fs / 40 + 1)))
    b0 = int(np.ceil(100 / (fs / N)) + 1)
    b1 = int(np.ceil(4000 / (fs / N)) + 1)
    b2 = int(np.ceil(7900 / (fs / N)) + 1)
    f = np.maximum(f0[cand], 40.0)
    seg, win, mask, half = C.pitch_windows(x, fs, f, pos[cand], 1.5, "blackman")
    wav = C.remove_weighted_mean(seg, win, half)
    p = np.abs(np.fft.fft(wav, N, axis=1)) ** 2
    p[:, :b0] = 0.0
    cum = np.cumsum(p, axis=1)
    ok[cand] = (cum[:, b1 - 1] / cum[:, b2 - 1]) > threshold
    return ok


def _centroid(x, fs, f0, pos, N):
    """Energy centroid (group-delay numerator) of one 4*T0 Blackman-windowed
    segment (d4c.py:132-153)."""
    seg, win, mask, half = C.pitch_windows(x, fs, f0, pos, 2, "blackman")
    wav = C.remove_weighted_mean(seg, win, half)
    wav = wav / np.sqrt(np.sum(wav ** 2, axis=1, keepdims=True))
    ramp = np.arange(1, wav.shape[1] + 1)[None, :]
    s = np.fft.fft(wav, N, axis=1)
    w = np.fft.fft(-wav * ramp * 1j, N, axis=1)
    return -w.imag * s.real + s.imag * w.real


def coarse_aperiodicity(x, fs, f0, pos, N, interval, n_bands, window):
    """d4c.py:114-128 for frames f0[F] > 0; returns [F, n_bands] (positive dB)."""
    f0 = np.asarray(f0, dtype=np.float64)
    F = len(f0)
    quarter = 1.0 / f0 / 4
    cen = _centroid(x, fs, f0, pos + quarter, N) + _centroid(x, fs, f0, pos - quarter, N)
    cen = C.mirror_low_band(cen, fs, f0, "wide")

    seg, win, mask, half = C.pitch_windows(x, fs, f0, pos, 2, "hann")
    wav = C.remove_weighted_mean(seg, win, half)
    power = C.mirror_low_band(np.abs(np.fft.fft(wav, N, axis=1)) ** 2, fs, f0, "wide")

    def sym(h):
        return np.concatenate([h, h[:, -2:0:-1]], axis=1)

    smooth_power = sym(C.box_integral(power, fs, f0 / 2) / f0[:, None])          # d4c.py:157-161
    gd = cen / smooth_power                                                      # d4c.py:169
    gd = sym(C.box_integral(gd, fs, f0 / 4) / (f0 / 2)[:, None])                 # :170-171
    gd_half = gd[:, :N // 2 + 1] - C.box_integral(gd, fs, f0 / 2) / f0[:, None]  # :172-173
    gd = sym(gd_half)

    boundary = int(N / len(window) * 8 + 0.5)                                    # d4c.py:197
    hw = int(np.floor(len(window) / 2))
    out = np.zeros((F, n_bands))
    for b in range(n_bands):
        centre = int(np.floor(interval * (b + 1) / (fs / N)))
        piece = gd[:, centre - hw:centre + hw + 1] * window[None, :]
        p = np.abs(np.fft.fft(piece, N, axis=1)) ** 2
        cum = np.cumsum(np.sort(p[:, :N // 2 + 1], axis=1), axis=1)
        out[:, b] = -10 * np.log10(cum[:, N // 2 - boundary - 1] / cum[:, -1])
    return out


def _setup(fs, N, interval):
    n_bands = _band_count(fs, interval)
    assert n_bands > 0
    wlen = int(np.floor(interval / (fs / N)) * 2 + 1)
    return n_bands, C.nuttall(wlen)


def d4c(x, fs, temporal_positions, f0, vuv, threshold=0.85, fft_size_for_spectrum=None):
    """d4c.py:10-64.  Returns dict(aperiodicity [Ns/2+1, F], coarse_ap [n_bands, F],
    f0 [F] as left in the shared dict: 0 where vuv == 0)."""
    x = np.asarray(x, dtype=np.float64)
    tp = np.asarray(temporal_positions, dtype=np.float64)
    N = int(2 ** np.ceil(np.log2(4 * fs / 47 + 1)))
    Ns = int(fft_size_for_spectrum) if fft_size_for_spectrum is not None \
        else int(2 ** np.ceil(np.log2(3 * fs / 71 + 1)))
    interval = 2000 if fs < 16000 else 3000
    f = np.where(np.asarray(vuv) == 0, 0.0, np.asarray(f0, dtype=np.float64))
    n_bands, window = _setup(fs, N, interval)
    F = len(f)
    ap = np.full((Ns // 2 + 1, F), 1 - 0.000000000001)
    dbg = np.zeros((n_bands, F))
    ok = np.nonzero(love_train(x, fs, f, tp, threshold))[0]
    if len(ok):
        cf = np.maximum(47.0, f[ok])
        coarse = coarse_aperiodicity(x, fs, cf, tp[ok], N, interval, n_bands, window)
        coarse = np.maximum(0, coarse - ((cf - 100) * 2 / 100)[:, None])        # d4c.py:56
        dbg[:, ok] = -coarse.T
        axis_c = np.r_[np.arange(n_bands + 1) * interval, fs / 2]
        axis_f = np.arange(Ns / 2 + 1) * fs / Ns
        for j, i in enumerate(ok):
            knots = np.r_[-60, -coarse[j], -0.000000000001]
            ap[:, i] = 10 ** (C.lerp_extrap(axis_c, knots, axis_f) / 20)        # d4c.py:58-59
    return {"aperiodicity": ap, "coarse_ap": dbg, "f0": f}


def d4c_requiem(x, fs, temporal_positions, f0, vuv, threshold=0.85, fft_size=None):
    """d4cRequiem.py:9-44.  Returns dict(aperiodicity [n_bands+2, F] in dB, f0 [F])."""
    x = np.asarray(x, dtype=np.float64)
    tp = np.asarray(temporal_positions, dtype=np.float64)
    N = int(fft_size) if fft_size is not None else int(2 ** np.ceil(np.log2(3 * fs / 47 + 1)))
    interval = 3000
    f = np.where(np.asarray(vuv) == 0, 0.0, np.asarray(f0, dtype=np.float64))
    n_bands, window = _setup(fs, N, interval)
    F = len(f)
    ap = np.full((n_bands + 2, F), -0.000000000001)
    ok = np.nonzero(love_train(x, fs, f, tp, threshold))[0]
    ap[0, ok] = -60
    if len(ok):
        cf = np.maximum(47.0, f[ok])
        coarse = coarse_aperiodicity(x, fs, cf, tp[ok], N, interval, n_bands, window)
        ap[1:-1, ok] = -np.maximum(0, coarse - ((cf - 100) * 2 / 100)[:, None]).T
    return {"aperiodicity": ap, "f0": f}
