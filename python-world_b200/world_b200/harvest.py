"""world/harvest.py drop-in: harvest(x, fs, f0_floor, f0_ceil, frame_period) -> dict (GPU)."""
from . import _single as S


def harvest(x, fs, f0_floor=71, f0_ceil=800, frame_period=5):
    E = S.eng()
    X, ns = S.dev1(E, x)
    tp, f0, vuv, nf = E.harvest(X, ns, int(fs), float(f0_floor), float(f0_ceil), float(frame_period))
    return {'temporal_positions': tp[0].cpu().numpy(), 'f0': f0[0].cpu().numpy(), 'vuv': vuv[0].cpu().numpy()}
