"""world/cheaptrick.py drop-in: cheaptrick(x, fs, source_object, q1, fft_size) -> dict (GPU).

Like the reference, overwrites source_object['f0'] in place (500 at unvoiced / below-limit frames,
cheaptrick.py:27,33) and consumes np.random for the eps-dither (cheaptrick.py:117)."""
import numpy as np

from . import _single as S

EPS = 2.220446049250313e-16


def cheaptrick(x, fs, source_object, q1=-0.15, fft_size=None):
    E = S.eng()
    X, ns = S.dev1(E, x)
    f0_seq = source_object['f0']
    T, F0, V = S.frames1(E, source_object['temporal_positions'], f0_seq, source_object['vuv'])
    n = E.L.wb_cheaptrick_fft_size(int(fs)) if fft_size is None else int(fft_size)
    dither = np.abs(np.random.rand(len(f0_seq), n // 2 + 1)) * EPS
    f0u, spec, ps = E.cheaptrick(X, ns, int(fs), T, F0, V, E.i32([len(f0_seq)]), q1=q1, fft_size=n,
                                 dither=E.f64(dither[None]), want_ps=True)
    f0_seq[:] = f0u[0].cpu().numpy()
    return {'temporal_positions': source_object['temporal_positions'],
            'spectrogram': S.ref_matrix(spec[0]),
            'fs': fs,
            'ps spectrogram': S.ref_matrix(ps[0])}
