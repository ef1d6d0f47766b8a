"""CPU tier: sampling rates around the decimation-ratio edges (tests/golden/rates.npz, make_golden_rates.py):
8 000 Hz (Harvest pass-through), 11 025 Hz (ratio 1 but still filtered, harvest.py:61-69), 44 100 Hz.  The oracle
against the reference goldens, and the kernel bodies (host emulation) against the same goldens."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, legacy_dither
from oracle import cheaptrick as o_ct
from oracle import d4c as o_d4c
from oracle import dio as o_dio
from oracle import harvest as o_hv

RATES = (8000, 11025, 44100)


@pytest.fixture(scope="module")
def rates():
    return dict(np.load(os.path.join(GOLDEN, "rates.npz")))


def _check_spec(got, want):
    m = want > 1e-12
    return float(np.max(np.abs(np.log10(got[m]) - np.log10(want[m]))))


def test_oracle_rates(rates):
    g = rates
    for fs in RATES:
        t = "r%d_" % fs
        x = g[t + "x"]
        r = o_hv.harvest(x, fs)
        assert np.array_equal(r["vuv"], g[t + "harvest_vuv"]), fs
        assert np.max(np.abs(r["f0"] - g[t + "harvest_f0"])) < 1e-9, fs
        np.random.seed(0)
        c = o_ct.cheaptrick(x, fs, r["temporal_positions"], g[t + "harvest_f0"], g[t + "harvest_vuv"])
        assert np.array_equal(c["f0"], g[t + "f0_after_cheaptrick"])
        assert _check_spec(c["spectrogram"][:, ::4], g[t + "spectrogram"]) < 1e-6
        d = o_d4c.d4c(x, fs, r["temporal_positions"], g[t + "f0_after_cheaptrick"], g[t + "harvest_vuv"])
        assert np.max(np.abs(d["aperiodicity"][:, ::4] - g[t + "aperiodicity"])) < 1e-8
        q = o_dio.dio(x, fs)
        assert np.array_equal(q["vuv"], g[t + "dio_vuv"]) and np.max(np.abs(q["f0"] - g[t + "dio_f0"])) < 1e-9
        s = o_dio.stonemask(x, fs, q["temporal_positions"], g[t + "dio_f0"])
        assert np.max(np.abs(s - g[t + "stonemask_f0"])) < 1e-8


def test_emu_rates(emu, rates):
    g = rates
    for fs in RATES:
        t = "r%d_" % fs
        x = g[t + "x"]
        r = emu.harvest(x, fs)
        assert np.array_equal(r["vuv"][0], g[t + "harvest_vuv"]), fs
        v = g[t + "harvest_vuv"] > 0
        assert np.max(np.abs(r["f0"][0] - g[t + "harvest_f0"])[v] / g[t + "harvest_f0"][v]) < 1e-9, fs
        tp = r["temporal_positions"][0]
        nb = g[t + "spectrogram"].shape[0]
        f0u, spec, _ = emu.cheaptrick(x, fs, tp, g[t + "harvest_f0"], g[t + "harvest_vuv"],
                                      dither=legacy_dither(len(tp), nb)[None], want_ps=False)
        assert np.array_equal(f0u[0], g[t + "f0_after_cheaptrick"])
        assert _check_spec(spec[0].T[:, ::4], g[t + "spectrogram"]) < 1e-6
        f0o, ap, _ = emu.d4c(x, fs, tp, g[t + "f0_after_cheaptrick"], g[t + "harvest_vuv"])
        assert np.max(np.abs(ap[0].T[:, ::4] - g[t + "aperiodicity"])) < 1e-8
        q = emu.dio(x, fs)
        assert np.array_equal(q["vuv"][0], g[t + "dio_vuv"]) and np.max(np.abs(q["f0"][0] - g[t + "dio_f0"])) < 1e-8
        s = emu.stonemask(x, fs, tp, g[t + "dio_f0"])
        sm = g[t + "stonemask_f0"]
        m = sm > 0
        assert np.max(np.abs(s[0][m] - sm[m]) / sm[m]) < 1e-9 and np.all(s[0][~m] == 0)
