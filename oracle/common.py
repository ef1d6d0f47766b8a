"""Shared numeric helpers of the CPU oracle.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ may be imported by the product
path (python-world_b200/); only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs use it, and only as the checker.

The oracle is a NumPy restatement (vectorised over frames, written from the
algorithm, not transcribed) of the reference's hot path.  It is pinned against
outputs of the unmodified reference run in the build container
(tests/golden/*.npz, produced by tests/golden/make_golden.py).
"""
import numpy as np

EPS = float(np.finfo(np.float64).eps)


def trunc_int(v):
    """Python int() on a float array: truncation toward zero."""
    return np.trunc(v).astype(np.int64)


def half_away(v):
    """The reference's `round_matlab` (cheaptrick.py:161-172 and five copies):
    it does NOT round, it returns v+0.5 for v>0 and v-0.5 otherwise; callers
    truncate afterwards."""
    v = np.asarray(v, dtype=np.float64)
    return np.where(v > 0, v + 0.5, v - 0.5)


def round_half_up(v):
    """Decimal(v).quantize(0, ROUND_HALF_UP) for v >= 0 on the exact binary value
    (harvest.py:253, synthesis.py:132)."""
    v = np.asarray(v, dtype=np.float64)
    fl = np.floor(v)
    return (fl + ((v - fl) >= 0.5)).astype(np.int64)


def nuttall(n: int) -> np.ndarray:
    """4-term Nuttall window, endpoints included (d4c.py:245-249, dio.py:208-212,
    harvest.py:563-567)."""
    n = int(n)
    ang = np.arange(n) * 2 * np.pi / (n - 1)
    c = np.array([0.355768, -0.487396, 0.144232, -0.012604])
    # a matrix product, like the reference: DIO takes argmax of this window (dio.py:131) and the two
    # central samples of an even-length window differ only in the last bit, so the summation order
    # of the BLAS kernel decides the index
    return c @ np.cos(np.arange(4.0)[:, None] * ang[None, :])


def lerp_extrap(xk, yk, xq):
    """scipy interp1d(kind='linear', fill_value='extrapolate') for 1-D knots.
    Knots are sorted first (interp1d does, assume_sorted=False); outside the knot
    range the first / last segment is extended."""
    xk = np.asarray(xk, dtype=np.float64)
    yk = np.asarray(yk, dtype=np.float64)
    order = np.argsort(xk, kind="mergesort")
    xk, yk = xk[order], yk[order]
    hi = np.clip(np.searchsorted(xk, xq), 1, len(xk) - 1)
    lo = hi - 1
    slope = (yk[hi] - yk[lo]) / (xk[hi] - xk[lo])
    return slope * (xq - xk[lo]) + yk[lo]


def pitch_windows(x, fs, f0, pos, span, kind, subsample=True):
    """Pitch-synchronous windowed segments for many frames at once.

    cheaptrick.py:79-99 (span=1.5, Hann, no sub-sample term) and d4c.py:92-110
    (span 1.5/2, Hann or Blackman, sub-sample offset of the frame time).
    Returns (seg*win [F, L], win [F, L], mask [F, L], half [F]); columns beyond a
    frame's own 2*half+1 samples are zero.
    """
    f0 = np.asarray(f0, dtype=np.float64)
    pos = np.asarray(pos, dtype=np.float64)
    half = trunc_int(span * fs / f0 + 0.5)
    width = int(2 * half.max() + 1)
    col = np.arange(width)[None, :]
    k = col - half[:, None]                      # -half .. +half, then padding
    mask = col <= 2 * half[:, None]
    centre = trunc_int(pos * fs + 0.501) + 1     # 1-based
    idx = np.clip(centre[:, None] + k, 1, len(x))
    seg = x[idx - 1]
    t = k / fs / span
    if subsample:
        t = t + ((pos * fs - trunc_int(pos * fs + 0.5)) / fs)[:, None]
    arg = np.pi * t * f0[:, None]
    if kind == "hann":
        win = 0.5 * np.cos(arg) + 0.5
    else:  # blackman
        win = 0.08 * np.cos(2 * arg) + 0.5 * np.cos(arg) + 0.42
    win = np.where(mask, win, 0.0)
    seg = np.where(mask, seg, 0.0)
    return seg, win, mask, half


def remove_weighted_mean(seg, win, half):
    """seg*win - win * mean(seg*win)/mean(win)  (cheaptrick.py:98, d4c.py:109)."""
    cnt = (2 * half + 1).astype(np.float64)[:, None]
    sw = seg * win
    ratio = (sw.sum(axis=1, keepdims=True) / cnt) / (win.sum(axis=1, keepdims=True) / cnt)
    return sw - win * ratio


def mirror_low_band(sig, fs, f0, reach):
    """Add the spectrum mirrored about f0 to the bins below f0, then make the
    second half the mirror image of the first.

    sig [F, N] (full FFT length), f0 [F].  `reach` selects the knot set:
    bins with f < f0 + fs/N for CheapTrick (cheaptrick.py:67-74) and bins with
    f < 1.2 f0 for D4C (d4c.py:213-222).  Knots sit at f0 - f_k (descending), the
    value there is sig[k]; queries are the bin frequencies below f0.
    """
    F, N = sig.shape
    df = fs / N
    fax = np.arange(N) / N * fs
    out = sig.copy()
    for i in range(F):
        lim = f0[i] + df if reach == "one_bin" else 1.2 * f0[i]
        nk = int(np.count_nonzero(fax < lim))
        nq = int(np.count_nonzero(fax < f0[i]))
        if nq == 0:
            continue
        knots_x = f0[i] - fax[:nk]
        out[i, :nq] = sig[i, :nq] + lerp_extrap(knots_x, sig[i, :nk], fax[:nq])
    out[:, N - 1:N // 2:-1] = out[:, 1:N // 2]
    return out


def box_integral(full, fs, half_width):
    """Running-integral difference used for rectangular smoothing
    (cheaptrick.py:103-131, d4c.py:179-188).  full [F, N] is a symmetric
    spectrum; returns [F, N/2+1] = I(f + h) - I(f - h) with h = half_width[F] and
    I the piecewise-linear cumulative sum over the doubled axis starting at -fs.
    The caller applies its own normalisation (1.5/f0 or 1/width)."""
    F, N = full.shape
    df = fs / N
    doubled = np.concatenate([full, full], axis=1)
    integ = np.cumsum(doubled * df, axis=1)
    x0 = (np.arange(2 * N) / N * fs - fs + df / 2)
    step = x0[1] - x0[0]
    centre = np.arange(N // 2 + 1) / N * fs
    d_integ = np.concatenate([np.diff(integ, axis=1), np.zeros((F, 1))], axis=1)
    rows = np.arange(F)[:, None]

    def sample(xq):
        xq = np.maximum(x0[0], np.minimum(x0[-1], xq))
        p = (xq - x0[0]) / step
        b = np.floor(p)
        frac = p - b
        b = b.astype(np.int64)
        return integ[rows, b] + d_integ[rows, b] * frac

    h = np.asarray(half_width, dtype=np.float64)[:, None]
    return sample(centre[None, :] + h) - sample(centre[None, :] - h)


def frame_count(n_samples, fs, period_ms):
    """int(1000*N/fs/period + 1)  (dio.py:28, harvest.py:21,46)."""
    return int(1000 * n_samples / fs / period_ms + 1)
