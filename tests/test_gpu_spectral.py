"""GPU tier: CheapTrick / D4C / D4C-Requiem CUDA kernels through the C-ABI against the
reference goldens and the oracle."""
import numpy as np
import pytest

from conftest import legacy_dither, spec_close

pytestmark = pytest.mark.gpu


def _dev(engine, g, x, tp, f0, vuv):
    import torch
    X = engine.f64(np.atleast_2d(x))
    ns = engine.i32([X.shape[1]])
    T = engine.f64(np.atleast_2d(tp))
    F0 = engine.f64(np.atleast_2d(f0))
    V = engine.f64(np.atleast_2d(vuv))
    nf = engine.i32([T.shape[1]])
    return X, ns, T, F0, V, nf


def test_cheaptrick_gpu_syn16k(engine, syn16k):
    g = syn16k
    tp, f0, vuv = g["harvest_d4c_temporal_positions"], g["harvest_d4c_f0_tracker"], g["harvest_d4c_vuv"]
    X, ns, T, F0, V, nf = _dev(engine, g, g["x"], tp, f0, vuv)
    dz = engine.f64(legacy_dither(len(f0), 513)[None])
    f0u, spec, ps = engine.cheaptrick(X, ns, int(g["fs"]), T, F0, V, nf, dither=dz, want_ps=True)
    assert np.array_equal(f0u.cpu().numpy()[0], g["harvest_d4c_f0_after_cheaptrick"])
    p99, mx = spec_close(spec.cpu().numpy()[0].T, g["harvest_d4c_spectrogram"])
    assert p99 < 1e-6 and mx < 1e-5
    assert np.max(np.abs(ps.cpu().numpy()[0].T[:, ::4] - g["harvest_d4c_ps_spectrogram"])) < 1e-12


def test_cheaptrick_gpu_mwm_and_48k(engine, mwm, syn48k):
    for g, tag in ((mwm, "dio_d4c_"), (syn48k, "")):
        f0, vuv = g[tag + "f0_tracker"], g[tag + "vuv"]
        tp = np.arange(len(f0)) * 0.005
        fs = int(g["fs"])
        n = 1024 if fs < 40000 else 2048
        X, ns, T, F0, V, nf = _dev(engine, g, g["x"], tp, f0, vuv)
        dz = engine.f64(legacy_dither(len(f0), n // 2 + 1)[None])
        f0u, spec, _ = engine.cheaptrick(X, ns, fs, T, F0, V, nf, dither=dz)
        assert np.array_equal(f0u.cpu().numpy()[0], g[tag + "f0_after_cheaptrick"])
        st = int(g[tag + "frame_stride"])
        p99, mx = spec_close(spec.cpu().numpy()[0].T[:, ::st], g[tag + "spectrogram"])
        assert p99 < 1e-5 and mx < 1e-3, (tag, p99, mx)


def test_d4c_gpu(engine, syn16k, mwm):
    g = syn16k
    tp, vuv = g["harvest_d4c_temporal_positions"], g["harvest_d4c_vuv"]
    f0 = g["harvest_d4c_f0_after_cheaptrick"]
    X, ns, T, F0, V, nf = _dev(engine, g, g["x"], tp, f0, vuv)
    f0o, ap, co = engine.d4c(X, ns, int(g["fs"]), T, F0, V, nf, want_coarse=True)
    assert np.array_equal(f0o.cpu().numpy()[0], g["harvest_d4c_f0"])
    assert np.max(np.abs(ap.cpu().numpy()[0].T - g["harvest_d4c_aperiodicity"])) < 1e-8
    assert np.max(np.abs(co.cpu().numpy()[0].T - g["harvest_d4c_coarse_ap"])) < 1e-6
    f0o, apr = engine.d4c_requiem(X, ns, int(g["fs"]), T, F0, V, nf)
    assert np.max(np.abs(apr.cpu().numpy()[0].T[:, ::4] - g["harvest_req_aperiodicity"])) < 1e-6
    # 22 050 Hz, full file, two bands
    g = mwm
    st = int(g["dio_d4c_frame_stride"])
    X, ns, T, F0, V, nf = _dev(engine, g, g["x"], g["dio_d4c_temporal_positions"],
                               g["dio_d4c_f0_after_cheaptrick"], g["dio_d4c_vuv"])
    f0o, ap, co = engine.d4c(X, ns, int(g["fs"]), T, F0, V, nf, want_coarse=True)
    assert np.array_equal(f0o.cpu().numpy()[0], g["dio_d4c_f0"])
    assert np.max(np.abs(ap.cpu().numpy()[0].T[:, ::st] - g["dio_d4c_aperiodicity"])) < 1e-8
    assert np.max(np.abs(co.cpu().numpy()[0].T - g["dio_d4c_coarse_ap"])) < 1e-6


def test_batch_ragged_gpu(engine, syn16k):
    """Two utterances of different length in one batch: frames beyond n_frames stay untouched,
    and each utterance equals its single-utterance result."""
    import torch
    g = syn16k
    fs = int(g["fs"])
    tp, f0, vuv = g["harvest_d4c_temporal_positions"], g["harvest_d4c_f0_tracker"], g["harvest_d4c_vuv"]
    x = g["x"]
    S2, F2 = 9000, 100
    X = engine.f64(np.stack([x, np.r_[x[:S2], np.zeros(len(x) - S2)]]))
    ns = engine.i32([len(x), S2])
    T = engine.f64(np.stack([tp, tp]))
    F0 = engine.f64(np.stack([f0, f0]))
    V = engine.f64(np.stack([vuv, vuv]))
    nf = engine.i32([len(f0), F2])
    f0u, spec, _ = engine.cheaptrick(X, ns, fs, T, F0, V, nf, seed=1)
    X1 = engine.f64(x[None, :S2])
    f0u1, spec1, _ = engine.cheaptrick(X1, engine.i32([S2]), fs, T[:1, :F2].contiguous(), F0[:1, :F2].contiguous(),
                                       V[:1, :F2].contiguous(), engine.i32([F2]), seed=1)
    a, b = spec[1, :F2].cpu().numpy(), spec1[0].cpu().numpy()
    m = b > 1e-10
    assert np.max(np.abs(np.log10(a[m]) - np.log10(b[m]))) < 1e-4
