"""CPU tier: DIO + StoneMask -- oracle pinned to the reference goldens, kernel bodies (host emulation)
against the same goldens, and the bit-exact Nuttall table the DIO index bias depends on."""
import numpy as np

from oracle import common as OC
from oracle import dio as o_dio


def _x(g):
    return g["x"], int(g["fs"])


def test_oracle_dio_stonemask(mwm, syn16k):
    for g in (mwm, syn16k):
        x, fs = _x(g)
        r = o_dio.dio(x, fs)
        assert np.array_equal(r["vuv"], g["dio_d4c_vuv"])
        assert np.max(np.abs(r["f0"] - g["dio_d4c_dio_f0"])) < 1e-9
        assert np.max(np.abs(r["raw_f0_candidates"] - g["dio_d4c_dio_raw_f0_candidates"])) < 1e-7
        assert np.max(np.abs(r["f0_candidates"] - g["dio_d4c_dio_f0_candidates"])) < 1e-7
        f = o_dio.stonemask(x, fs, r["temporal_positions"], g["dio_d4c_dio_f0"])
        assert np.max(np.abs(f - g["dio_d4c_f0_tracker"])) < 1e-10


def test_nuttall_table_bit_exact(emu):
    """The library's Nuttall table must equal the matrix-product form of the reference bit for bit
    (argmax of an even-length window decides DIO's index bias, dio.py:131)."""
    for n in list(range(4, 200, 4)) + [37, 385, 493, 513, 557, 769]:
        assert np.array_equal(emu.nuttall(n), OC.nuttall(n)), n


def test_emu_dio_stonemask(emu, mwm, syn16k):
    for g in (mwm, syn16k):
        x, fs = _x(g)
        r = emu.dio(x, fs)
        assert np.array_equal(r["vuv"][0], g["dio_d4c_vuv"])
        assert np.max(np.abs(r["f0"][0] - g["dio_d4c_dio_f0"])) < 1e-9
        raw, G = r["raw_f0_candidates"][0], g["dio_d4c_dio_raw_f0_candidates"]
        assert np.array_equal(raw != 0, G != 0)
        assert np.max(np.abs(raw - G)) < 1e-7
        assert np.max(np.abs(r["f0_candidates"][0].T - g["dio_d4c_dio_f0_candidates"])) < 1e-7
        f = emu.stonemask(x, fs, r["temporal_positions"], g["dio_d4c_dio_f0"])
        assert np.max(np.abs(f[0] - g["dio_d4c_f0_tracker"])) < 1e-9


def test_emu_dio_ragged(emu, syn16k):
    x = syn16k["x"]
    X = np.stack([x, np.r_[x[:9001], np.zeros(6999)]])
    r = emu.dio(X, 16000, n_samples=[16000, 9001])
    r1 = emu.dio(x[:9001], 16000)
    n1 = r1["f0"].shape[1]
    assert list(r["n_frames"]) == [201, n1]
    assert np.array_equal(r["f0"][1, :n1], r1["f0"][0])
    ro = o_dio.dio(x[:9001], 16000)
    assert np.array_equal(r1["vuv"][0], ro["vuv"])
    assert np.max(np.abs(r1["f0"][0] - ro["f0"])) < 1e-9
