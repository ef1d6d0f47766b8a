"""ctypes declarations of the C-ABI in include/world_b200.h (one table, used by the
product loader world_b200._lib and by the test-only host-emulation loader)."""
import ctypes as C

P = C.c_void_p
I = C.c_int
D = C.c_double
U64 = C.c_uint64

SIGNATURES = {
    "wb_is_cuda_build": (I, []),
    "wb_version": (C.c_char_p, []),
    "wb_create": (I, [C.POINTER(P), I]),
    "wb_destroy": (I, [P]),
    "wb_last_error": (C.c_char_p, [P]),
    "wb_frame_count": (I, [I, I, D]),
    "wb_cheaptrick_fft_size": (I, [I]),
    "wb_d4c_band_count": (I, [I, I]),
    "wb_synthesis_length": (I, [D, D, I]),
    # h, stream, x, x_stride, n_samples, batch, fs, tpos, f0, vuv, n_frames, f_stride, ...
    "wb_cheaptrick": (I, [P, P, P, I, P, I, I, P, P, P, P, I, D, I, P, U64, P, P, P]),
    "wb_d4c": (I, [P, P, P, I, P, I, I, P, P, P, P, I, D, I, P, P, P]),
    "wb_d4c_requiem": (I, [P, P, P, I, P, I, I, P, P, P, P, I, D, I, P, P]),
    # h, batch, max_samples, fs, f0_floor, f0_ceil, *bytes
    "wb_harvest_workspace_bytes": (I, [P, I, I, I, D, D, C.POINTER(C.c_size_t)]),
    "wb_harvest_workspace_layout": (I, [P, I, I, I, D, D, C.POINTER(C.c_size_t), C.POINTER(I)]),
    # h, stream, x, x_stride, n_samples, batch, max_samples, fs, floor, ceil, period, ws, ws_bytes,
    # f_stride, tpos, f0, vuv, n_frames
    "wb_harvest": (I, [P, P, P, I, P, I, I, I, D, D, D, P, C.c_size_t, I, P, P, P, P]),
    "wb_harvest_stages": (I, [P, P, P, I, P, I, I, I, D, D, D, P, C.c_size_t, I, P, P, P, P, I, I]),
    # h, batch, max_samples, fs, floor, ceil, channels_in_octave, target_fs, period, *bytes
    "wb_dio_workspace_bytes": (I, [P, I, I, I, D, D, I, I, D, C.POINTER(C.c_size_t)]),
    # h, stream, x, x_stride, n_samples, batch, max_samples, fs, floor, ceil, cio, target_fs, period, allowed,
    # ws, ws_bytes, f_stride, tpos, f0, vuv, n_frames, f0_candidates|NULL, raw_f0_candidates|NULL
    "wb_dio": (I, [P, P, P, I, P, I, I, I, D, D, I, I, D, D, P, C.c_size_t, I, P, P, P, P, P, P]),
    "wb_dio_band_count": (I, [D, D, I]),
    # h, stream, x, x_stride, n_samples, batch, fs, tpos, f0, n_frames, f_stride, out
    "wb_stonemask": (I, [P, P, P, I, P, I, I, P, P, P, I, P]),
    "wb_debug_nuttall": (I, [I, P]),
    # h, stream, spec, rows, n_bins, preemph_abs, filterbank, n_filt, out
    "wb_lfbank": (I, [P, P, P, I, I, P, P, I, P]),
    # h, stream, spec, rows, n_bins, mel_bin, n0, out
    "wb_mcep": (I, [P, P, P, I, I, P, I, P]),
    # h, stream, cepstrum, rows, n0, fft_size, mel_pos, bracket, query, out
    "wb_mcep_decode": (I, [P, P, P, I, I, I, P, P, P, P]),
    # h, stream, in, rows, n_in, knots, bracket, query, n_out, out
    "wb_interp_rows": (I, [P, P, P, I, I, P, P, P, I, P]),
    # h, stream, x, n, knot_x, knot_y, n_knots, out
    "wb_interp_knots": (I, [P, P, P, C.c_longlong, P, P, I, P]),
    # h, stream, pcm, pcm_stride, n_samples, batch, divisor, x, x_stride
    "wb_pcm16_to_f64": (I, [P, P, P, I, P, I, D, P, I]),
    # h, stream, y, y_stride, n_samples, batch, gain, pcm, pcm_stride
    "wb_f64_to_pcm16": (I, [P, P, P, I, P, I, D, P, I]),
    # h, batch, y_stride, requiem_rows, *bytes
    "wb_synthesis_workspace_bytes": (I, [P, I, I, I, C.POINTER(C.c_size_t)]),
    # h, stream, tpos, f0, vuv, n_frames, batch, f_stride, fs, y_stride, ws, ws_bytes, requiem_rows,
    # out_len, n_pulses, noise_total
    "wb_synthesis_timebase": (I, [P, P, P, P, P, P, I, I, I, I, P, C.c_size_t, I, P, P, P]),
    # h, stream, tpos, f0, vuv, spec, ap, n_frames, batch, f_stride, fs, fft, ws, ws_bytes, noise, noise_stride,
    # seed, y, y_stride, normalize
    "wb_synthesis": (I, [P, P, P, P, P, P, P, P, I, I, I, I, P, C.c_size_t, P, I, U64, P, I, I]),
    # h, stream, tpos, f0, vuv, spec, band_ap, n_frames, batch, f_stride, fs, fft, rows, pulse_seed, seed_fft,
    # noise_seed, noise_len, cursor_in, cursor_out, ws, ws_bytes, y, y_stride, normalize
    "wb_synthesis_requiem": (I, [P, P, P, P, P, P, P, P, I, I, I, I, I, P, I, P, I, P, P, P, C.c_size_t, P, I, I]),
}


class EncodeParams(C.Structure):
    """wb_encode_params (include/world_b200.h)."""
    _fields_ = [("fs", I), ("f0_method", I), ("f0_floor", D), ("f0_ceil", D), ("channels_in_octave", I),
                ("target_fs", I), ("frame_period_ms", D), ("allowed_range", D), ("fft_size", I), ("requiem", I),
                ("q1", D), ("threshold", D), ("seed", U64)]


F0_METHODS = {"harvest": 0, "dio": 1}

SIGNATURES.update({
    # h, params*, batch, max_samples, *bytes
    "wb_encode_workspace_bytes": (I, [P, C.POINTER(EncodeParams), I, I, C.POINTER(C.c_size_t)]),
    # h, stream, params*, x, x_stride, n_samples, batch, max_samples, ws, ws_bytes, f_stride, dither, tpos, f0, vuv,
    # n_frames, spectrogram, aperiodicity|NULL, coarse_ap|NULL, ps|NULL
    "wb_encode": (I, [P, P, C.POINTER(EncodeParams), P, I, P, I, I, P, C.c_size_t, I, P, P, P, P, P, P, P, P, P]),
    # h, stream, coarse_ap, rows, fs, fft_size_for_spectrum, aperiodicity
    "wb_d4c_expand": (I, [P, P, P, C.c_longlong, I, I, P]),
    # h, batch, y_stride, requiem_rows, *bytes
    # h, stream, in, n, out
    "wb_f64_to_f32": (I, [P, P, P, C.c_longlong, P]),
    "wb_f32_to_f64": (I, [P, P, P, C.c_longlong, P]),
    # h, stream, threads, iters, out, *flops
    "wb_probe_dfma": (I, [P, P, C.c_longlong, I, P, C.POINTER(D)]),
    "wb_decode_workspace_bytes": (I, [P, I, I, I, C.POINTER(C.c_size_t)]),
    # h, stream, fs, fft, tpos, f0, vuv, spec, ap, n_frames, batch, f_stride, requiem_rows, pulse_seed, seed_fft,
    # noise_seed, noise_len, cursor_in, cursor_out, noise, noise_stride, seed, ws, ws_bytes, y, y_stride, normalize, out_len
    "wb_decode": (I, [P, P, I, I, P, P, P, P, P, P, I, I, I, P, I, P, I, P, P, P, I, U64, P, C.c_size_t, P, I, I, P]),
})


def declare(lib):
    """Attach restype/argtypes; raises AttributeError if a symbol is missing."""
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib
