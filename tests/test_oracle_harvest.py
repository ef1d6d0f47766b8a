"""CPU tier: Harvest oracle pinned stage by stage against the reference goldens, and the kernel
bodies (host emulation) against the same goldens."""
import numpy as np

from oracle import harvest as o_hv


def _dense(r, nc, shape, keep_only=False):
    f = np.zeros(shape)
    s = np.zeros(shape)
    for j in range(shape[1]):
        for q in range(r["l_n"][0, j]):
            if keep_only and not r["l_keep"][0, j, q]:
                continue
            sh, k = divmod(int(r["l_slot"][0, j, q]), 15)
            f[sh * nc + k, j] = r["l_f0"][0, j, q]
            s[sh * nc + k, j] = r["l_sc"][0, j, q]
    return f, s


def test_oracle_harvest_stages(syn16k):
    g = syn16k
    st = {}
    r = o_hv.harvest(g["x"], int(g["fs"]), stages=st)
    assert np.max(np.abs(st["y"] - g["hv_y"])) < 1e-14
    for name, key, tol in (("raw", "hv_raw", 1e-8), ("detect", "hv_detect", 1e-8),
                           ("refined_f0", "hv_refined_f0", 1e-9), ("reliable_f0", "hv_reliable_f0", 1e-9),
                           ("connected", "hv_connected", 1e-9), ("smoothed", "hv_smoothed", 1e-9)):
        a, b = st[name], g[key]
        assert a.shape == b.shape, name
        assert np.array_equal(a != 0, b != 0), name
        assert np.max(np.abs(a - b)) < tol, name
    assert st["ncand"] == int(g["hv_ncand"])
    assert np.max(np.abs(st["refined_score"] - g["hv_refined_score"]) / np.maximum(g["hv_refined_score"], 1)) < 1e-9
    assert np.array_equal(r["vuv"], g["harvest_d4c_vuv"])
    assert np.max(np.abs(r["f0"] - g["harvest_d4c_f0_tracker"])) < 1e-9


def test_oracle_harvest_other_rates(mwm, syn48k):
    r = o_hv.harvest(mwm["x"], int(mwm["fs"]))
    assert np.array_equal(r["vuv"], mwm["harvest_req_vuv"])
    assert np.max(np.abs(r["f0"] - mwm["harvest_req_f0_tracker"])) < 1e-9
    r = o_hv.harvest(syn48k["x"], 48000)
    assert np.array_equal(r["vuv"], syn48k["vuv"])
    assert np.max(np.abs(r["f0"] - syn48k["f0_tracker"])) < 1e-9


def test_emu_harvest_stages(emu, syn16k):
    g = syn16k
    r = emu.harvest(g["x"], 16000, debug=True)
    assert r["status"][0] == 0
    F1 = g["hv_raw"].shape[1]
    assert np.max(np.abs(r["y"][0, :r["y_len"][0]] - g["hv_y"])) < 1e-14
    raw = r["raw"][0][:, :F1]
    assert np.array_equal(raw != 0, g["hv_raw"] != 0)
    assert np.max(np.abs(raw - g["hv_raw"])) < 1e-8
    nc = int(g["hv_ncand"])
    assert r["base_n"][0, :F1].max() == nc
    f, s = _dense(r, nc, g["hv_refined_f0"].shape)
    assert np.array_equal(f != 0, g["hv_refined_f0"] != 0)
    assert np.max(np.abs(f - g["hv_refined_f0"])) < 1e-9
    assert np.max(np.abs(s - g["hv_refined_score"]) / np.maximum(g["hv_refined_score"], 1)) < 1e-9
    f, s = _dense(r, nc, g["hv_reliable_f0"].shape, keep_only=True)
    assert np.array_equal(f != 0, g["hv_reliable_f0"] != 0)
    assert np.array_equal(r["vuv"][0], g["harvest_d4c_vuv"])
    assert np.max(np.abs(r["f0"][0] - g["harvest_d4c_f0_tracker"])) < 1e-9
    assert np.array_equal(r["temporal_positions"][0], g["harvest_d4c_temporal_positions"])


def test_emu_harvest_rates_ragged_silence(emu, mwm, syn48k, syn16k):
    r = emu.harvest(syn48k["x"], 48000)
    assert np.array_equal(r["vuv"][0], syn48k["vuv"])
    assert np.max(np.abs(r["f0"][0] - syn48k["f0_tracker"])) < 1e-9
    x = mwm["x"][:33075]  # 1.5 s of the 22 050 Hz fixture (decimation ratio 3, 7350 Hz)
    ro = o_hv.harvest(x, 22050)
    r = emu.harvest(x, 22050)
    assert np.array_equal(r["vuv"][0], ro["vuv"])
    assert np.max(np.abs(r["f0"][0] - ro["f0"])) < 1e-9
    x = syn16k["x"]
    X = np.stack([x, np.r_[x[:9000], np.zeros(7000)], np.zeros(16000)])
    r = emu.harvest(X, 16000, n_samples=[16000, 9000, 16000])
    assert list(r["n_frames"]) == [201, 113, 201]
    r1 = emu.harvest(x[:9000], 16000)
    assert np.array_equal(r["f0"][1, :113], r1["f0"][0])
    assert np.all(r["f0"][2] == 0) and np.all(r["vuv"][2] == 0)
