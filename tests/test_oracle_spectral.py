"""CPU tier: the oracle (NumPy restatement) pinned against outputs of the unmodified
reference (tests/golden/*.npz) for CheapTrick, D4C and D4C-Requiem."""
import numpy as np

from oracle import cheaptrick as o_ct
from oracle import d4c as o_d4c


def _src(g, tag):
    return g[tag + "temporal_positions"], g[tag + "f0_tracker"], g[tag + "vuv"]


def test_cheaptrick_syn16k(syn16k):
    g = syn16k
    tp, f0, vuv = _src(g, "harvest_d4c_")
    np.random.seed(0)
    r = o_ct.cheaptrick(g["x"], int(g["fs"]), tp, f0, vuv)
    S = g["harvest_d4c_spectrogram"]
    assert np.max(np.abs(r["spectrogram"] - S) / S) < 1e-9
    assert np.array_equal(r["f0"], g["harvest_d4c_f0_after_cheaptrick"])
    assert np.max(np.abs(r["ps_spectrogram"][:, ::4] - g["harvest_d4c_ps_spectrogram"])) < 1e-13


def test_cheaptrick_mwm_and_48k(mwm, syn48k):
    for g, tag, x in ((mwm, "dio_d4c_", mwm["x"]), (syn48k, "", syn48k["x"])):
        tp = g.get(tag + "temporal_positions")
        f0, vuv = g[tag + "f0_tracker"], g[tag + "vuv"]
        if tp is None:
            tp = np.arange(len(f0)) * 0.005
        np.random.seed(0)
        r = o_ct.cheaptrick(x, int(g["fs"]), tp, f0, vuv)
        st = int(g[tag + "frame_stride"])
        S = g[tag + "spectrogram"]
        m = S > 1e-12
        assert np.max(np.abs(np.log10(r["spectrogram"][:, ::st][m]) - np.log10(S[m]))) < 1e-6
        assert np.array_equal(r["f0"], g[tag + "f0_after_cheaptrick"])


def test_d4c_syn16k(syn16k):
    g = syn16k
    tp, _, vuv = _src(g, "harvest_d4c_")
    f0 = g["harvest_d4c_f0_after_cheaptrick"]
    d = o_d4c.d4c(g["x"], int(g["fs"]), tp, f0, vuv)
    assert np.max(np.abs(d["aperiodicity"] - g["harvest_d4c_aperiodicity"])) < 1e-10
    assert np.max(np.abs(d["coarse_ap"] - g["harvest_d4c_coarse_ap"])) < 1e-8
    assert np.array_equal(d["f0"], g["harvest_d4c_f0"])
    r = o_d4c.d4c_requiem(g["x"], int(g["fs"]), tp, f0, vuv)
    assert np.max(np.abs(r["aperiodicity"][:, ::4] - g["harvest_req_aperiodicity"])) < 1e-8


def test_d4c_mwm(mwm):
    g = mwm
    st = int(g["dio_d4c_frame_stride"])
    d = o_d4c.d4c(g["x"], int(g["fs"]), g["dio_d4c_temporal_positions"], g["dio_d4c_f0_after_cheaptrick"],
                  g["dio_d4c_vuv"])
    assert np.max(np.abs(d["aperiodicity"][:, ::st] - g["dio_d4c_aperiodicity"])) < 1e-9
    assert np.max(np.abs(d["coarse_ap"] - g["dio_d4c_coarse_ap"])) < 1e-7
    r = o_d4c.d4c_requiem(g["x"], int(g["fs"]), g["harvest_req_temporal_positions"],
                          g["harvest_req_f0_after_cheaptrick"], g["harvest_req_vuv"])
    assert np.max(np.abs(r["aperiodicity"][:, ::st] - g["harvest_req_aperiodicity"])) < 1e-7
