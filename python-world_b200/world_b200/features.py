"""Size-dependent tables for the feature heads and spectrum / time-axis edits.

Host-side companions of csrc/wb_features.h: everything here depends only on sizes and scalar
parameters (never on the signal), is evaluated once per call with the reference's own NumPy
expressions so the tables agree with it bit for bit, and is uploaded as a few hundred values.
The per-frame arithmetic runs in the CUDA kernels.
"""
import numpy as np


def hz2mel(hz):
    """main.py:258-264"""
    return 2595 * np.log10(1 + hz / 700.)


def mel2hz(mel):
    """main.py:266-272"""
    return 700 * (10 ** (mel / 2595.0) - 1)


def mel_filterbank(nfilt=20, nfft=512, samplerate=16000, lowfreq=0, highfreq=None):
    """Triangular mel filters [nfilt, nfft/2+1] (get_filterbanks, main.py:274-303)."""
    highfreq = highfreq or samplerate / 2
    assert highfreq <= samplerate / 2, "highfreq is greater than samplerate/2"
    edges = np.floor((nfft + 1) * mel2hz(np.linspace(hz2mel(lowfreq), hz2mel(highfreq), nfilt + 2)) / samplerate)
    fb = np.zeros([nfilt, nfft // 2 + 1])
    for j in range(nfilt):
        lo, mid, hi = edges[j], edges[j + 1], edges[j + 2]
        rise = np.arange(int(lo), int(mid))
        fall = np.arange(int(mid), int(hi))
        fb[j, rise] = (rise - lo) / (mid - lo)
        fb[j, fall] = (hi - fall) / (hi - mid)
    return fb


def preemphasis_abs(prefac, n_bins):
    """|H(e^{jw})| of the pre-emphasis filter [1, -prefac] at w = pi k / n_bins (freqz, main.py:313; SciPy
    evaluates an FIR response at an integer number of points through a zero-padded real FFT)."""
    return np.abs(np.fft.rfft(np.array([1.0, -prefac]), n=2 * n_bins)[:n_bins])


def mel_source_bins(n_bins, fs=16000, lowhz=0, highhz=8000):
    """Integer source bin of every point of the mel-spaced axis (main.py:331-337)."""
    pts = np.linspace(hz2mel(lowhz), hz2mel(highhz), n_bins)
    return np.floor(((n_bins - 1) * 2 + 1) * mel2hz(pts) / fs)


def mel_positions(fft_size):
    """Positions of the mel-spaced samples on the linear bin axis (decode_mcep, main.py:348-355: 8 kHz / 16 kHz fixed)."""
    pts = np.linspace(hz2mel(0), hz2mel(8000), int(fft_size // 2 + 1))
    return np.floor(fft_size * mel2hz(pts) / 16000)


def interp_brackets(xp, xq):
    """For numpy.interp(xq, xp, .): bracket j (largest index with xp[j] <= xq, clipped to [0, len-1]) and the
    query to evaluate at (queries left of xp[0] take the value at xp[0])."""
    xp = np.asarray(xp, dtype=np.float64)
    xq = np.asarray(xq, dtype=np.float64)
    j = np.searchsorted(xp, xq, side="right") - 1
    below = j < 0
    j = np.clip(j, 0, len(xp) - 1)
    q = np.where(below, xp[0], xq)
    return j.astype(np.int32), q.astype(np.float64)


# ------------------------------------------------------------------------------------------------
# Launchers.  `ops` is the object that owns the library handle and the device arrays: engine.Engine in
# the product (torch CUDA tensors), the host-emulation driver in the CPU test tier (NumPy arrays).  It
# provides L, h, f64(), i32(), empty(), ptr(), _stream(), _check().  Arrays are [..., bins] with the
# frame axes flattened into rows.
# ------------------------------------------------------------------------------------------------
def _rows(a):
    n = 1
    for s in a.shape[:-1]:
        n *= int(s)
    return n


def lfbank(ops, spec, prefac=0.97, fs=16000, nfilt=32, lowfreq=0, highfreq=None):
    d = int(spec.shape[-1])
    fb = ops.f64(mel_filterbank(nfilt, (d - 1) * 2, fs, lowfreq, highfreq))
    ha = ops.f64(preemphasis_abs(prefac, d))
    out = ops.empty(*spec.shape[:-1], nfilt)
    ops._check(ops.L.wb_lfbank(ops.h, ops._stream(), ops.ptr(spec), _rows(spec), d, ops.ptr(ha), ops.ptr(fb), int(nfilt),
                               ops.ptr(out)))
    return out


def mcep(ops, spec, n0=12, fs=16000, lowhz=0, highhz=8000):
    d = int(spec.shape[-1])
    bins = ops.i32(mel_source_bins(d, fs, lowhz, highhz).astype(np.int32))
    out = ops.empty(*spec.shape[:-1], int(n0))
    ops._check(ops.L.wb_mcep(ops.h, ops._stream(), ops.ptr(spec), _rows(spec), d, ops.ptr(bins), int(n0), ops.ptr(out)))
    return out


def mcep_decode(ops, cepstrum, fft_size):
    n0 = int(cepstrum.shape[-1])
    d = int(fft_size) // 2 + 1
    pos = mel_positions(fft_size)
    j, q = interp_brackets(pos, np.arange(d))
    out = ops.empty(*cepstrum.shape[:-1], d)
    pos_d, j_d, q_d = ops.f64(pos), ops.i32(j), ops.f64(q)  # named: the tables must outlive the enqueue
    ops._check(ops.L.wb_mcep_decode(ops.h, ops._stream(), ops.ptr(cepstrum), _rows(cepstrum), n0, int(fft_size),
                                    ops.ptr(pos_d), ops.ptr(j_d), ops.ptr(q_d), ops.ptr(out)))
    return out


def warp_rows(ops, spec, factor, out=None):
    """numpy.interp((k/D)**factor, k/D, row) for every row of spec [..., D] (warp_spectrum, main.py:189-194);
    out=spec warps in place."""
    d = int(spec.shape[-1])
    grid = np.arange(0, d) / d
    j, q = interp_brackets(grid, grid ** factor)
    if out is None:
        out = ops.empty(*spec.shape)
    grid_d, j_d, q_d = ops.f64(grid), ops.i32(j), ops.f64(q)
    ops._check(ops.L.wb_interp_rows(ops.h, ops._stream(), ops.ptr(spec), _rows(spec), d, ops.ptr(grid_d),
                                    ops.ptr(j_d), ops.ptr(q_d), d, ops.ptr(out)))
    return out


def interp_knots(ops, x, knot_x, knot_y, out=None):
    """numpy.interp(x, knot_x, knot_y) element-wise on a device array; out=x works in place."""
    kx, ky = np.asarray(knot_x, dtype=np.float64), np.asarray(knot_y, dtype=np.float64)
    assert kx.ndim == 1 and kx.shape == ky.shape and len(kx) >= 1
    if out is None:
        out = ops.empty(*x.shape)
    n = 1
    for s in x.shape:
        n *= int(s)
    kx_d, ky_d = ops.f64(kx), ops.f64(ky)
    ops._check(ops.L.wb_interp_knots(ops.h, ops._stream(), ops.ptr(x), n, ops.ptr(kx_d), ops.ptr(ky_d), len(kx),
                                     ops.ptr(out)))
    return out
