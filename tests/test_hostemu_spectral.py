"""CPU tier: the CUDA kernel bodies, compiled for the host with one emulated thread
per block (tests/hostemu), against the reference goldens.  Checks kernel LOGIC where
no GPU exists; the GPU tier repeats the same comparisons on the real kernels."""
import numpy as np

from conftest import legacy_dither, spec_close


def test_cheaptrick_emu(emu, syn16k):
    g = syn16k
    tp, f0, vuv = g["harvest_d4c_temporal_positions"], g["harvest_d4c_f0_tracker"], g["harvest_d4c_vuv"]
    dz = legacy_dither(len(f0), 513)
    f0u, spec, ps = emu.cheaptrick(g["x"], int(g["fs"]), tp, f0, vuv, dither=dz[None])
    assert np.array_equal(f0u[0], g["harvest_d4c_f0_after_cheaptrick"])
    p99, mx = spec_close(spec[0].T, g["harvest_d4c_spectrogram"])
    assert p99 < 1e-6 and mx < 1e-5
    assert np.max(np.abs(ps[0].T[:, ::4] - g["harvest_d4c_ps_spectrogram"])) < 1e-12


def test_cheaptrick_emu_48k(emu, syn48k):
    g = syn48k
    f0, vuv = g["f0_tracker"], g["vuv"]
    tp = np.arange(len(f0)) * 0.005
    f0u, spec, _ = emu.cheaptrick(g["x"], int(g["fs"]), tp, f0, vuv, dither=legacy_dither(len(f0), 1025)[None],
                                  want_ps=False)
    assert np.array_equal(f0u[0], g["f0_after_cheaptrick"])
    p99, mx = spec_close(spec[0].T[:, ::4], g["spectrogram"])
    assert p99 < 1e-6 and mx < 1e-4


def test_d4c_emu(emu, syn16k):
    g = syn16k
    tp, vuv = g["harvest_d4c_temporal_positions"], g["harvest_d4c_vuv"]
    f0 = g["harvest_d4c_f0_after_cheaptrick"]
    f0o, ap, co = emu.d4c(g["x"], int(g["fs"]), tp, f0, vuv)
    assert np.array_equal(f0o[0], g["harvest_d4c_f0"])
    assert np.max(np.abs(ap[0].T - g["harvest_d4c_aperiodicity"])) < 1e-8
    assert np.max(np.abs(co[0].T - g["harvest_d4c_coarse_ap"])) < 1e-6
    f0o, apr = emu.d4c_requiem(g["x"], int(g["fs"]), tp, f0, vuv)
    assert np.max(np.abs(apr[0].T[:, ::4] - g["harvest_req_aperiodicity"])) < 1e-6


def test_d4c_emu_mwm_subset(emu, mwm):
    """22 050 Hz fixture (two aperiodicity bands), every 8th frame."""
    g = mwm
    st = int(g["dio_d4c_frame_stride"])
    tp = g["dio_d4c_temporal_positions"][::st]
    f0 = g["dio_d4c_f0_after_cheaptrick"][::st]
    vuv = g["dio_d4c_vuv"][::st]
    f0o, ap, co = emu.d4c(g["x"], int(g["fs"]), tp, f0, vuv)
    assert np.max(np.abs(ap[0].T - g["dio_d4c_aperiodicity"])) < 1e-8
    assert np.max(np.abs(co[0].T - g["dio_d4c_coarse_ap"][:, ::st])) < 1e-6
