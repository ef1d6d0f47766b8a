"""world/synthesis.py drop-in: synthesis(source_object, filter_object) -> ndarray (GPU).

The per-pulse Gaussian noise is drawn with np.random.randn on the host in the reference's order
(synthesis.py:93) and uploaded, so a seeded call reproduces the reference's waveform."""
import numpy as np

from . import _single as S


def synthesis(source_object, filter_object, normalize=False):
    E = S.eng()
    tp = np.asarray(source_object['temporal_positions'], dtype=np.float64)
    fs = filter_object['fs']
    T, F0, V = S.frames1(E, tp, source_object['f0'], source_object['vuv'])
    spec = S.dev_matrix(E, filter_object['spectrogram'])
    ap = S.dev_matrix(E, source_object['aperiodicity'])
    ylen = E.synthesis_length(tp[0], tp[-1], fs)
    y, out_len = E.synthesis(T, F0, V, spec, ap, E.i32([len(tp)]), int(fs), ylen, noise="legacy",
                             normalize=normalize)
    return y[0, :int(out_len[0])].cpu().numpy()
