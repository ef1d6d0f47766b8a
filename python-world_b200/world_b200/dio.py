"""world/dio.py drop-in: dio(x, fs, ...) -> dict (GPU)."""
import numpy as np

from . import _single as S


def dio(x, fs, f0_floor=71, f0_ceil=800, channels_in_octave=2, target_fs=4000, frame_period=5, allowed_range=0.1):
    E = S.eng()
    X, ns = S.dev1(E, x)
    tp, f0, vuv, nf, cand, raw = E.dio(X, ns, int(fs), float(f0_floor), float(f0_ceil), int(channels_in_octave),
                                       int(target_fs), float(frame_period), float(allowed_range), want_candidates=True)
    return {'f0': f0[0].cpu().numpy(),
            'f0_candidates': np.ascontiguousarray(cand[0].cpu().numpy().T),
            'raw_f0_candidates': raw[0].cpu().numpy(),
            'temporal_positions': tp[0].cpu().numpy(),
            'vuv': vuv[0].cpu().numpy()}
