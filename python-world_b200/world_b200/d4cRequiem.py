"""world/d4cRequiem.py drop-in: d4cRequiem(x, fs, f0_object, threshold, fft_size) -> the same dict (GPU)."""
from . import _single as S


def d4cRequiem(x, fs, f0_object, threshold=0.85, fft_size=None):
    E = S.eng()
    X, ns = S.dev1(E, x)
    f0_seq = f0_object['f0']
    T, F0, V = S.frames1(E, f0_object['temporal_positions'], f0_seq, f0_object['vuv'])
    if E.L.wb_d4c_band_count(int(fs), 1) <= 0:
        raise AssertionError("number_of_aperiodicities > 0")  # d4cRequiem.py:21
    f0o, ap = E.d4c_requiem(X, ns, int(fs), T, F0, V, E.i32([len(f0_seq)]), threshold=threshold, fft_size=fft_size)
    f0_seq[:] = f0o[0].cpu().numpy()
    f0_object['aperiodicity'] = S.ref_matrix(ap[0])
    return f0_object
