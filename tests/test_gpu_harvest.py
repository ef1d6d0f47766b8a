"""GPU tier: Harvest CUDA pipeline through the C-ABI against reference goldens."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

F0_RTOL = 1e-6  # identical vuv on >= 99.5 % of frames, voiced rel. error <= 1e-6


def _run(engine, x, fs, n_samples=None):
    X = engine.f64(np.atleast_2d(x))
    ns = engine.i32([X.shape[1]] * X.shape[0] if n_samples is None else n_samples)
    tp, f0, vuv, nf = engine.harvest(X, ns, fs, max_samples=int(max(ns.cpu().numpy())))
    return tp.cpu().numpy(), f0.cpu().numpy(), vuv.cpu().numpy(), nf.cpu().numpy()


def _check(f0, vuv, f0g, vg):
    assert np.array_equal(vuv, vg)  # SURVEY 8d allows 0.5 % flips; none are measured, so none are accepted
    both = (vuv > 0) & (vg > 0)
    assert np.max(np.abs(f0[both] - f0g[both]) / f0g[both]) <= F0_RTOL


def test_harvest_gpu_syn16k(engine, syn16k):
    g = syn16k
    tp, f0, vuv, nf = _run(engine, g["x"], 16000)
    assert nf[0] == len(g["harvest_d4c_f0_tracker"])
    assert np.array_equal(tp[0], g["harvest_d4c_temporal_positions"])
    _check(f0[0], vuv[0], g["harvest_d4c_f0_tracker"], g["harvest_d4c_vuv"])
    assert np.array_equal(vuv[0], g["harvest_d4c_vuv"])


def test_harvest_gpu_mwm_and_48k(engine, mwm, syn48k):
    tp, f0, vuv, nf = _run(engine, mwm["x"], int(mwm["fs"]))
    _check(f0[0], vuv[0], mwm["harvest_req_f0_tracker"], mwm["harvest_req_vuv"])
    tp, f0, vuv, nf = _run(engine, syn48k["x"], 48000)
    _check(f0[0], vuv[0], syn48k["f0_tracker"], syn48k["vuv"])


def test_harvest_gpu_ragged_and_silence(engine, syn16k):
    x = syn16k["x"]
    X = np.stack([x, np.r_[x[:9000], np.zeros(7000)], np.zeros(16000)])
    tp, f0, vuv, nf = _run(engine, X, 16000, n_samples=[16000, 9000, 16000])
    assert list(nf) == [201, 113, 201]
    _check(f0[0], vuv[0], syn16k["harvest_d4c_f0_tracker"], syn16k["harvest_d4c_vuv"])
    tp1, f01, vuv1, nf1 = _run(engine, x[:9000], 16000)
    assert np.allclose(f0[1, :113], f01[0], rtol=1e-9, atol=0)
    assert np.all(f0[2] == 0) and np.all(vuv[2] == 0)  # the reference raises IndexError here


def test_dio_stonemask_gpu(engine, mwm, syn16k):
    for g in (mwm, syn16k):
        x, fs = g["x"], int(g["fs"])
        X = engine.f64(x[None])
        ns = engine.i32([len(x)])
        tp, f0, vuv, nf, cand, raw = engine.dio(X, ns, fs, want_candidates=True)
        assert np.array_equal(vuv.cpu().numpy()[0], g["dio_d4c_vuv"])
        assert np.max(np.abs(f0.cpu().numpy()[0] - g["dio_d4c_dio_f0"])) < 1e-8
        assert np.max(np.abs(raw.cpu().numpy()[0] - g["dio_d4c_dio_raw_f0_candidates"])) < 1e-6
        f = engine.stonemask(X, ns, fs, tp, engine.f64(g["dio_d4c_dio_f0"][None]), nf)
        fr = f.cpu().numpy()[0]
        gt = g["dio_d4c_f0_tracker"]
        m = gt > 0
        assert np.max(np.abs(fr[m] - gt[m]) / gt[m]) < 1e-9
        assert np.all(fr[~m] == 0)
