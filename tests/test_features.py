"""CPU tier: the feature-head / edit / PCM oracle (oracle/features.py) against the reference goldens
(tests/golden/features.npz, make_golden_features.py), the product's table builders against the oracle's, and the
kernel bodies under host emulation against both."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import features as o_ft

TOL = 1e-10  # log-domain / cepstral values of O(1..20); the heads are sums of <= 2049 float64 terms


@pytest.fixture(scope="module")
def feat():
    return dict(np.load(os.path.join(GOLDEN, "features.npz")))


CASES = (("k16_", 16000), ("k48_", 48000))


def test_oracle_feature_heads(feat):
    for tag, fs in CASES:
        mag = feat[tag + "spec"]
        assert np.max(np.abs(o_ft.encode_lfbank(mag, fs=fs) - feat[tag + "lfbank"])) < 1e-12
        got = o_ft.encode_lfbank(mag, prefac=0.9, fs=fs, nfilt=20, lowfreq=100, highfreq=fs / 2 - 500)
        assert np.max(np.abs(got - feat[tag + "lfbank_20_hi"])) < 1e-12
        assert np.max(np.abs(o_ft.encode_mcep(mag, n0=40, fs=fs, highhz=min(8000, fs // 2)) - feat[tag + "mcep40"])) < 1e-12
        assert np.max(np.abs(o_ft.encode_mcep(mag, fs=fs) - feat[tag + "mcep12"])) < 1e-12
        dec = o_ft.decode_mcep(feat[tag + "mcep40"], (mag.shape[1] - 1) * 2)
        assert np.max(np.abs(np.log(dec) - np.log(feat[tag + "decoded"]))) < 1e-12


def test_oracle_edits(feat):
    for tag, _ in CASES:
        spec = np.ascontiguousarray(feat[tag + "spec"].T)
        for f in (0.8, 1.25):
            assert np.array_equal(o_ft.warp_spectrum(spec, f), feat[tag + "warp_%g" % f])
    assert np.array_equal(o_ft.modify_duration(feat["dur_tp"], [1, 1.5], [0, 1, 3, -1]), feat["dur_out"])


def test_tables_match_oracle():
    from world_b200 import features as F
    for nf, nfft, fs, lo, hi in ((32, 1024, 16000, 0, None), (20, 2048, 48000, 100, 23500.0), (40, 512, 8000, 0, None)):
        assert np.array_equal(F.mel_filterbank(nf, nfft, fs, lo, hi), o_ft.filterbank(nf, nfft, fs, lo, hi))
    from scipy.signal import freqz
    for d in (513, 1025):
        assert np.max(np.abs(F.preemphasis_abs(0.97, d) - np.abs(freqz([1, -0.97], [1], d)[1]))) < 1e-15
    xp = np.array([0., 0., 1., 3., 3., 3., 7.])
    xq = np.array([-1., 0., 0.5, 1., 2., 3., 5., 7., 9.])
    j, q = F.interp_brackets(xp, xq)
    fp = np.arange(7.) ** 2
    want = np.interp(xq, xp, fp)
    got = np.array([fp[-1] if jj >= 6 else (fp[jj] if xp[jj] == qq else (fp[jj + 1] - fp[jj]) / (xp[jj + 1] - xp[jj]) * (qq - xp[jj]) + fp[jj])
                    for jj, qq in zip(j, q)])
    assert np.array_equal(got, want)


def test_emu_feature_heads(emu, feat):
    from world_b200 import features as F
    for tag, fs in CASES:
        mag = np.ascontiguousarray(feat[tag + "spec"])
        assert np.max(np.abs(F.lfbank(emu, mag, fs=fs) - feat[tag + "lfbank"])) < TOL
        got = F.lfbank(emu, mag, prefac=0.9, fs=fs, nfilt=20, lowfreq=100, highfreq=fs / 2 - 500)
        assert np.max(np.abs(got - feat[tag + "lfbank_20_hi"])) < TOL
        assert np.max(np.abs(F.mcep(emu, mag, n0=40, fs=fs, highhz=min(8000, fs // 2)) - feat[tag + "mcep40"])) < TOL
        assert np.max(np.abs(F.mcep(emu, mag, fs=fs) - feat[tag + "mcep12"])) < TOL
        dec = F.mcep_decode(emu, np.ascontiguousarray(feat[tag + "mcep40"]), (mag.shape[1] - 1) * 2)
        assert np.max(np.abs(np.log(dec) - np.log(feat[tag + "decoded"]))) < TOL


def test_emu_edits_and_pcm(emu, feat, mwm):
    from world_b200 import features as F
    for tag, _ in CASES:
        mag = np.ascontiguousarray(feat[tag + "spec"])  # rows = frames
        for f in (0.8, 1.25):
            assert np.array_equal(F.warp_rows(emu, mag, f).T, feat[tag + "warp_%g" % f])
        inplace = mag.copy()
        F.warp_rows(emu, inplace, 1.25, out=inplace)
        assert np.array_equal(inplace.T, feat[tag + "warp_1.25"])
    tp = feat["dur_tp"]
    got = F.interp_knots(emu, tp.copy(), np.r_[0, [1, 1.5], tp[-1]], [0, 1, 3, tp[-1]])
    assert np.array_equal(got, feat["dur_out"])
    pcm = mwm["x_int16"][None, :5000]
    x = emu.pcm16_to_f64(pcm, [5000])
    assert np.array_equal(x[0], o_ft.pcm16_to_float(pcm[0]))
    y = np.r_[x[0, :4000] * 1.3, [1.0, -1.0, 0.999999, -0.5, 0.0, 0.99996, -0.99999]][None]
    n = y.shape[1]
    assert np.array_equal(emu.f64_to_pcm16(y, [n])[0], o_ft.float_to_pcm16(y[0]))
    assert np.all(emu.f64_to_pcm16(y, [10])[0, 10:] == 0)


def test_interp_brackets_random():
    """The bracket tables the interpolation kernels consume reproduce numpy.interp for random non-decreasing knots
    with repeats, queries on / between / outside the knots (the kernel formula is restated here in NumPy)."""
    from world_b200 import features as F
    rng = np.random.default_rng(7)
    for _ in range(200):
        n = int(rng.integers(2, 40))
        xp = np.sort(rng.integers(0, 25, size=n)).astype(np.float64) * (1.0 if rng.random() < 0.5 else 0.37)
        fp = rng.standard_normal(n)
        xq = np.r_[rng.uniform(xp[0] - 2, xp[-1] + 2, size=30), xp[rng.integers(0, n, size=10)]]
        j, q = F.interp_brackets(xp, xq)
        got = np.empty(len(xq))
        for i, (jj, qq) in enumerate(zip(j, q)):
            if jj >= n - 1:
                got[i] = fp[n - 1]
            elif xp[jj] == qq:
                got[i] = fp[jj]
            else:
                got[i] = (fp[jj + 1] - fp[jj]) / (xp[jj + 1] - xp[jj]) * (qq - xp[jj]) + fp[jj]
        assert np.array_equal(got, np.interp(xq, xp, fp))
