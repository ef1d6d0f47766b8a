"""world/d4c.py drop-in: d4c(x, fs, f0_object, threshold, fft_size_for_spectrum) -> the same dict (GPU)."""
from . import _single as S


def d4c(x, fs, f0_object, threshold=0.85, fft_size_for_spectrum=None):
    E = S.eng()
    X, ns = S.dev1(E, x)
    f0_seq = f0_object['f0']
    T, F0, V = S.frames1(E, f0_object['temporal_positions'], f0_seq, f0_object['vuv'])
    if E.L.wb_d4c_band_count(int(fs), 0) <= 0:
        raise AssertionError("number_of_aperiodicity > 0")  # d4c.py:35
    f0o, ap, co = E.d4c(X, ns, int(fs), T, F0, V, E.i32([len(f0_seq)]), threshold=threshold,
                        fft_size_for_spectrum=fft_size_for_spectrum, want_coarse=True)
    f0_seq[:] = f0o[0].cpu().numpy()
    f0_object['aperiodicity'] = S.ref_matrix(ap[0])
    f0_object['coarse_ap'] = S.ref_matrix(co[0])
    return f0_object
