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


def test_fused_encode_and_coarse_transport(emu, syn16k):
    """wb_encode (csrc/wb_pipeline.cu) sequences the stage entry points like main.py:106-152: its outputs equal the
    stage calls', the band values rebuild the aperiodicity bit for bit (wb_d4c_expand), the host expansion of
    world_b200.main follows the reference's expressions (d4c.py:56-59), and WB_AP_NONE is World.get_spectrum."""
    import os
    import sys
    from world_b200 import main as wmain
    g = syn16k
    x = g["x"][:6000]
    F = emu.L.wb_frame_count(len(x), 16000, 5.0)
    dz = np.abs(np.random.RandomState(0).rand(1, F, 513)) * 2.220446049250313e-16
    d = emu.encode(x, 16000, "harvest", dither=dz)
    hv = emu.harvest(x, 16000)
    f0u, spec, _ = emu.cheaptrick(x, 16000, hv["temporal_positions"], hv["f0"], hv["vuv"], dither=dz, want_ps=False)
    f0o, ap, co = emu.d4c(x, 16000, hv["temporal_positions"], f0u, hv["vuv"])
    assert np.array_equal(d["f0"], f0o) and np.array_equal(d["vuv"], hv["vuv"]) and np.array_equal(d["spectrogram"], spec)
    assert np.array_equal(d["aperiodicity"], ap) and np.array_equal(d["coarse_ap"], co)
    assert np.array_equal(emu.expand_aperiodicity(co, 16000), ap)
    host = wmain.expand_coarse_ap(co, 16000)
    assert np.max(np.abs(host - ap)) < 1e-14
    # coarse-only output skips the matrix; requiem and spectrum-only modes
    c = emu.encode(x, 16000, "harvest", dither=dz, aperiodicity="coarse")
    assert c["aperiodicity"] is None and np.array_equal(c["coarse_ap"], co)
    r = emu.encode(x, 16000, "dio", is_requiem=True, dither=dz)
    f0r, apr = emu.d4c_requiem(x, 16000, r["temporal_positions"], emu.cheaptrick(
        x, 16000, r["temporal_positions"], emu.stonemask(x, 16000, r["temporal_positions"], emu.dio(x, 16000)["f0"]),
        r["vuv"], dither=dz, want_ps=False)[0], r["vuv"])
    assert np.array_equal(r["aperiodicity"], apr) and np.array_equal(r["f0"], f0r)
    s = emu.encode(x, 16000, "harvest", dither=dz, aperiodicity="none")
    assert np.array_equal(s["f0"], f0u) and np.array_equal(s["spectrogram"], spec)
    # errors: unknown tracker (main.py:136-137), wrong frame stride
    import ctypes as C
    from world_b200 import _abi
    q = _abi.EncodeParams(16000, 7, 71.0, 800.0, 2, 4000, 5.0, 0.1, 0, 0, -0.15, 0.85, 0)
    nbytes = C.c_size_t()
    assert emu.L.wb_encode_workspace_bytes(emu.h, C.byref(q), 1, 6000, C.byref(nbytes)) == -1
