"""DIO F0 estimator and StoneMask refinement -- oracle restatement of world/dio.py and
world/stonemask.py.  TEST INFRASTRUCTURE (see oracle/__init__.py).
"""
import math

import numpy as np
from scipy import signal

from . import common as C
from .harvest import crossing_intervals

# dio.py:359-436: 3rd-order low-pass per decimation ratio; y = B(z)/A(z) with
# A = 1 - a0 z^-1 - a1 z^-2 - a2 z^-3,  B = b0 + b1 z^-1 + b1 z^-2 + b0 z^-3
DECIMATOR = {
    2: ((0.041156734567757189, -0.42599112459189636, 0.041037215479961225), (0.16797464681802227, 0.50392394045406674)),
    3: ((0.95039378983237421, -0.67429146741526791, 0.15412211621346475), (0.071221945171178636, 0.21366583551353591)),
    4: ((1.4499664446880227, -0.98943497080950582, 0.24578252340690215), (0.036710750339322612, 0.11013225101796784)),
    5: ((1.7610939654280557, -1.2554914843859768, 0.3237186507788215), (0.021334858522387423, 0.06400457556716227)),
    6: ((1.9715352749512141, -1.4686795689225347, 0.3893908434965701), (0.013469181309343825, 0.040407543928031475)),
    7: ((2.1225239019534703, -1.6395144861046302, 0.44469707800587366), (0.0090366882681608418, 0.027110064804482525)),
    8: ((2.2357462340187593, -1.7780899984041358, 0.49152555365968692), (0.0063522763407111993, 0.019056829022133598)),
    9: ((2.3236003491759578, -1.8921545617463598, 0.53148928133729068), (0.0046331164041389372, 0.013899349212416812)),
    10: ((2.3936475118069387, -1.9873904075111861, 0.5658879979027055), (0.0034818622251927556, 0.010445586675578267)),
    11: ((2.450743295230728, -2.06794904601978, 0.59574774438332101), (0.0026822508007163792, 0.0080467524021491377)),
    12: ((2.4981398605924205, -2.1368928194784025, 0.62187513816221485), (0.0021097275904709001, 0.0063291827714127002)),
}


def _lowpass(x, r):
    """FilterForDecimate (dio.py:359-446); ratios outside 2..12 yield zeros, as in the reference."""
    if r not in DECIMATOR:
        return np.zeros_like(x)
    (a0, a1, a2), (b0, b1) = DECIMATOR[r]
    return signal.lfilter([b0, b1, b1, b0], [1.0, -a0, -a1, -a2], x)


def decimate(x, r):
    """dio.py:451-477: mirror-extend by 9 samples, filter forward and backward, keep every r-th sample."""
    x = np.asarray(x, dtype=np.float64)
    n = len(x)
    ext = np.concatenate([2 * x[0] - x[9:0:-1], x, 2 * x[-1] - x[-2:-11:-1]])
    z = _lowpass(_lowpass(ext, r)[::-1], r)[::-1]
    n_out = np.ceil(n / r + 1)
    first = int(r - r * n_out + n)
    return z[np.arange(first, n + 9, r) + 9 - 1]


def band_edges(f0_floor, f0_ceil, channels_in_octave):
    """dio.py:32-34."""
    k = np.arange(math.ceil(np.log2(f0_ceil / f0_floor) * channels_in_octave)) + 1
    return f0_floor * 2.0 ** (k / channels_in_octave)


def lowcut_spectrum(y, fs, f0_floor):
    """get_spectrum (dio.py:74-88): Hann-shaped DC/low-cut filter applied in the frequency domain."""
    n = 2 ** math.ceil(math.log(len(y) + int(fs / f0_floor / 2 + 0.5) * 4, 2))
    c = int(fs / 50 + 0.5)
    h = signal.windows.hann(2 * c + 3)[1:-1]
    h = -h / np.sum(h)
    h[c] += 1
    h = np.concatenate([h, np.zeros(n - len(h))])
    h = np.concatenate([h[c:], h[:c]])
    return np.fft.fft(y, n) * np.fft.fft(h, n)


def band_candidates(edge, fs, spec, y_len, times, f0_floor, f0_ceil):
    """get_raw_event + get_f0_candidates (dio.py:128-185)."""
    half = int(fs / edge / 2 + 0.5)
    lpf = C.nuttall(half * 4)
    bias = int(np.argmax(lpf))
    filt = np.real(np.fft.ifft(np.fft.fft(lpf, len(spec)) * spec))
    s = filt[bias + 1:bias + 1 + y_len]
    d = np.diff(s)
    streams = [crossing_intervals(v, fs) for v in (s, -s, d, -d)]
    usable = 1
    for loc, _ in streams:
        usable *= max(0, len(loc) - 2)
    if usable > 0:
        four = np.stack([C.lerp_extrap(loc, f, times) for loc, f in streams])
        est = np.mean(four, axis=0)
        dev = np.std(four, axis=0, ddof=1)
    else:
        est = times * 0
        dev = times * 0 + 1000
    est[est > edge] = 0
    est[est < edge / 2] = 0
    est[est > f0_ceil] = 0
    est[est < f0_floor] = 0
    dev[est == 0] = 100000
    return est, dev


def _round6(v):
    """float('{:.6f}'.format(v)) element-wise (dio.py:243)."""
    return np.array([float("{0:.6f}".format(e)) for e in v])


def _predict_pick(cur, past, column, tol):
    """select_best_f0 (dio.py:310-323): candidate nearest to the linear prediction, first minimum."""
    ref = (cur * 3 - past) / 2
    best = column[int(np.argmin(np.abs(ref - column)))]
    if abs(1 - best / (ref + C.EPS)) > tol:
        best = 0.0
    return best


def _sections(f0):
    """count_voiced_sections (dio.py:327-340), including its start-at-1 quirk."""
    v = (f0 != 0).astype(np.float64)
    d = np.diff(v)
    bl = np.concatenate([[0], np.nonzero(d != 0)[0], [len(v) - 2]]).astype(np.int64)
    first = int(np.ceil(-0.5 * d[bl[1]]))
    n = int(np.floor((len(bl) - (1 - first)) / 2))
    return [(1 + bl[2 * i + (1 - first)], bl[2 * i + 1 + (1 - first)]) for i in range(n)]


def fix_contour(cands, frame_period, f0_floor, tol):
    """fix_f0_contour (dio.py:216-326).  `cands` [7, F] sorted by stability; row 0 is modified in
    place exactly as the reference does (first/last frames zeroed) before steps 3 and 4 read it."""
    vrm = int(1 / (frame_period / 1000) / f0_floor + 0.5) * 2 + 1
    base = cands[0]
    base[:vrm] = 0
    base[-vrm:] = 0
    n = len(base)
    s1 = base.copy()
    r = _round6(base)
    for i in range(vrm - 1, n):
        if abs((r[i] - r[i - 1]) / (0.000001 + r[i])) > tol:
            s1[i] = 0
    hw = (vrm - 1) // 2
    s2 = s1.copy()
    for i in range(hw, n - hw):
        if np.any(s1[i - hw:i + hw + 1] == 0):
            s2[i] = 0
    secs = _sections(s2)
    s3 = s2.copy()
    for i, (st, ed) in enumerate(secs):
        limit = n - 1 if i == len(secs) - 1 else secs[i + 1][0] + 1
        for j in range(int(ed), int(limit)):
            s3[j + 1] = _predict_pick(s3[j], s3[j - 1], cands[:, j + 1], tol)
            if s3[j + 1] == 0:
                break
    s4 = s3.copy()
    for i in range(len(secs) - 1, -1, -1):
        limit = 1 if i == 0 else secs[i - 1][1]
        for j in range(int(secs[i][0]), int(limit) - 1, -1):
            s4[j - 1] = _predict_pick(s4[j], s4[j + 1], cands[:, j - 1], tol)
            if s4[j - 1] == 0:
                break
    return s4, (s4 != 0).astype(np.float64)


def dio(x, fs, f0_floor=71, f0_ceil=800, channels_in_octave=2, target_fs=4000, frame_period=5, allowed_range=0.1,
        stages=None):
    """dio.py:10-55.  Returns dict(f0, vuv, temporal_positions, f0_candidates, raw_f0_candidates)."""
    x = np.asarray(x, dtype=np.float64)
    n = C.frame_count(len(x), fs, frame_period)
    tp = np.arange(0, n) * frame_period / 1000
    edges = band_edges(f0_floor, f0_ceil, channels_in_octave)
    y = decimate(x, int(fs / target_fs))
    afs = target_fs                                   # dio.py:39: assumed, whatever fs / ratio really is
    spec = lowcut_spectrum(y, afs, f0_floor)
    raw = np.zeros((len(edges), n))
    stab = np.zeros((len(edges), n))
    for i, e in enumerate(edges):
        est, dev = band_candidates(e, afs, spec, len(y), tp, f0_floor, f0_ceil)
        stab[i] = np.exp(-(dev / np.maximum(est, 0.0000001)))
        raw[i] = est
    order = np.argsort(-stab, axis=0, kind="stable")
    cands = np.take_along_axis(raw, order, axis=0)
    kept = cands.copy()
    f0, vuv = fix_contour(cands, frame_period, f0_floor, allowed_range)
    if stages is not None:
        stages.update(y=y, raw=raw, stability=stab)
    return {"f0": f0, "f0_candidates": kept, "raw_f0_candidates": raw, "temporal_positions": tp, "vuv": vuv}


# ------------------------------------------------------------------------------------- StoneMask
def _round4_axis(half, fs):
    """float('{:.4f}'.format(k / fs)) for k = -half..half (stonemask.py:38)."""
    return np.array([float("{0:.4f}".format(e)) for e in np.arange(-half, half + 1) / fs])


def _refine_one(x, fs, t, f0):
    """get_refined_f0 (stonemask.py:30-76)."""
    half = int(np.ceil(3 * fs / f0 / 2))
    span = (2 * half + 1) / fs
    n_fft = 2 ** math.ceil(math.log((half * 2 + 1), 2) + 1)
    raw_idx = C.half_away((t + _round4_axis(half, fs)) * fs)
    wt = (raw_idx - 1) / fs - t
    main = 0.42 + 0.5 * np.cos(2 * math.pi * wt / span) + 0.08 * np.cos(4 * math.pi * wt / span)
    padded = np.concatenate([[0.0], main, [0.0]])
    dwin = -(padded[2:] - padded[:-2]) / 2
    seg = x[np.clip(raw_idx, 1, len(x)).astype(np.int64) - 1]
    S = np.fft.fft(seg * main, n_fft)
    D = np.fft.fft(seg * dwin, n_fft)
    power = np.abs(S) ** 2
    power[power == 0] = C.EPS
    inst = np.arange(n_fft) / n_fft * fs + (S.real * D.imag - S.imag * D.real) / power * fs / 2 / math.pi

    def harmonic_mean(f, count):
        h = np.arange(1, count + 1)
        idx = (C.half_away(f * n_fft / fs * h) + 1).astype(np.int64)
        amp = np.sqrt(power[idx - 1])
        return np.sum(amp * inst[idx - 1]) / np.sum(amp * h)

    f1 = harmonic_mean(f0, 2)
    if f1 < 0:
        return 0.0
    return harmonic_mean(f1, 6)


def stonemask(x, fs, temporal_positions, f0):
    """stonemask.py:8-27."""
    x = np.asarray(x, dtype=np.float64)
    out = np.array(f0, dtype=np.float64)
    for i, t in enumerate(temporal_positions):
        if f0[i] != 0:
            r = _refine_one(x, fs, t, f0[i])
            if abs(r - f0[i]) / f0[i] > 0.2:
                r = f0[i]
            out[i] = r
    return out
