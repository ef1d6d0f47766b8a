"""GPU tier: sampling rates around the decimation-ratio edges against the reference goldens (tests/golden/rates.npz):
8 000 Hz (Harvest pass-through), 11 025 Hz (ratio 1, still filtered), 44 100 Hz."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, legacy_dither

pytestmark = pytest.mark.gpu


def test_rates_gpu(engine):
    g = dict(np.load(os.path.join(GOLDEN, "rates.npz")))
    for fs in (8000, 11025, 44100):
        t = "r%d_" % fs
        x = g[t + "x"]
        X, ns = engine.f64(x[None]), engine.i32([len(x)])
        tp, f0, vuv, nf = engine.harvest(X, ns, fs)
        assert np.array_equal(vuv.cpu().numpy()[0], g[t + "harvest_vuv"]), fs
        v = g[t + "harvest_vuv"] > 0
        assert np.max(np.abs(f0.cpu().numpy()[0] - g[t + "harvest_f0"])[v] / g[t + "harvest_f0"][v]) < 1e-6, fs
        F0, V = engine.f64(g[t + "harvest_f0"][None]), engine.f64(g[t + "harvest_vuv"][None])
        nb = g[t + "spectrogram"].shape[0]
        f0u, spec, _ = engine.cheaptrick(X, ns, fs, tp, F0, V, nf, dither=engine.f64(legacy_dither(tp.shape[1], nb)[None]))
        assert np.array_equal(f0u.cpu().numpy()[0], g[t + "f0_after_cheaptrick"])
        S, W = spec.cpu().numpy()[0].T[:, ::4], g[t + "spectrogram"]
        m = W > 1e-10
        assert np.max(np.abs(np.log10(S[m]) - np.log10(W[m]))) < 1e-4
        f0o, ap, _ = engine.d4c(X, ns, fs, tp, f0u, V, nf)
        assert np.max(np.abs(ap.cpu().numpy()[0].T[:, ::4] - g[t + "aperiodicity"])) < 1e-5
        tq, fq, vq, nq = engine.dio(X, ns, fs)
        assert np.array_equal(vq.cpu().numpy()[0], g[t + "dio_vuv"])
        assert np.max(np.abs(fq.cpu().numpy()[0] - g[t + "dio_f0"])) < 1e-8
        s = engine.stonemask(X, ns, fs, tq, engine.f64(g[t + "dio_f0"][None]), nq).cpu().numpy()[0]
        sm = g[t + "stonemask_f0"]
        k = sm > 0
        assert np.max(np.abs(s[k] - sm[k]) / sm[k]) < 1e-9 and np.all(s[~k] == 0)
