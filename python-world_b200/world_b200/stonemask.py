"""world/stonemask.py drop-in: stonemask(x, fs, temporal_positions, f0) -> ndarray (GPU)."""
from . import _single as S


def stonemask(x, fs, temporal_positions, f0):
    E = S.eng()
    X, ns = S.dev1(E, x)
    T, F0 = S.frames1(E, temporal_positions, f0)
    return E.stonemask(X, ns, int(fs), T, F0, E.i32([len(f0)]))[0].cpu().numpy()
