"""world/synthesisRequiem.py drop-in: synthesisRequiem(source_object, filter_object, seeds_signals) -> ndarray (GPU).

generate_noise.current_index mirrors the reference's function attribute (synthesisRequiem.py:131-141): the
cyclic read position of the noise seeds persists across calls until it is reset to None."""
import numpy as np

from . import _single as S


def generate_noise(*_args, **_kw):
    raise NotImplementedError("generate_noise runs inside the CUDA kernel; only its .current_index state lives here")


generate_noise.current_index = None


def synthesisRequiem(source_object, filter_object, seeds_signals, normalize=False):
    E = S.eng()
    tp = np.asarray(source_object['temporal_positions'], dtype=np.float64)
    fs = filter_object['fs']
    T, F0, V = S.frames1(E, tp, source_object['f0'], source_object['vuv'])
    spec = S.dev_matrix(E, filter_object['spectrogram'])
    ap = S.dev_matrix(E, source_object['aperiodicity'])
    pulse = E.f64(np.ascontiguousarray(seeds_signals['pulse']))
    noise = E.f64(np.ascontiguousarray(seeds_signals['noise']))
    rows = pulse.shape[1]
    cur = generate_noise.current_index
    cursor = np.zeros(rows) if cur is None else np.asarray(cur, dtype=np.float64)
    ylen = E.synthesis_length(tp[0], tp[-1], fs)
    y, out_len, cur_out = E.synthesis_requiem(T, F0, V, spec, ap, E.i32([len(tp)]), int(fs), ylen, pulse, noise,
                                              cursor=cursor, normalize=normalize)
    generate_noise.current_index = cur_out[0].cpu().numpy()
    return y[0, :int(out_len[0])].cpu().numpy()
