"""Golden vectors for the partial pipelines (World.get_f0 / get_spectrum / encode_w_gvn_f0,
main.py:27-104) and for decode after a NON-uniform World.modify_duration (synthesis.py:50-52,121
and synthesisRequiem.py:78 with its truncated hop), by RUNNING the unmodified reference.
Build container only:  python tests/golden/make_golden_partial.py"""
import copy
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refload  # noqa: E402


def main():
    ref = refload.load()
    W = importlib.import_module("refworld.main").World()
    g = dict(np.load(os.path.join(HERE, "syn16k_1s.npz")))
    fs, x = int(g["fs"]), g["x"]
    out = {"fs": np.int64(fs)}
    # ---- partial pipelines
    for m in ("harvest", "dio"):
        refload.reseed(0)
        tp, f0, vuv = W.get_f0(fs, np.array(x), f0_method=m)
        out["get_f0_%s_tp" % m], out["get_f0_%s_f0" % m], out["get_f0_%s_vuv" % m] = tp, f0, vuv
    refload.reseed(0)
    sp = W.get_spectrum(fs, np.array(x), f0_method="dio")
    out["get_spectrum_f0"] = np.array(sp["f0"])  # as CheapTrick leaves it (cheaptrick.py:27,33)
    out["get_spectrum_spectrogram"] = np.array(sp["spectrogram"][:, ::4])
    # an external, fully voiced contour (the reference asserts f0 >= 3 fs / fft_size everywhere, main.py:86)
    F = len(g["harvest_d4c_temporal_positions"])
    f0_ext = 120.0 + 40.0 * np.sin(2 * np.pi * np.arange(F) / 70.0)
    # (is_requiem=True raises KeyError('coarse_ap') in the reference, main.py:102: d4cRequiem sets no such key)
    src = {"temporal_positions": np.array(g["harvest_d4c_temporal_positions"]), "f0": f0_ext.copy(), "vuv": np.ones(F)}
    refload.reseed(0)
    d = W.encode_w_gvn_f0(fs, np.array(x), src, fft_size=1024, is_requiem=False)
    out["gvn_d4c_f0_in"] = f0_ext
    out["gvn_d4c_f0"] = np.array(d["f0"])
    out["gvn_d4c_spectrogram"] = np.array(d["spectrogram"][:, ::4])
    out["gvn_d4c_aperiodicity"] = np.array(d["aperiodicity"][:, ::4])
    out["gvn_d4c_coarse_ap"] = np.array(d["coarse_ap"])
    # ---- decode after a non-uniform duration edit, both synthesisers
    for tag, req in (("harvest_d4c_", False), ("harvest_req_", True)):
        refload.reseed(0)
        dat = W.encode(fs, np.array(x), f0_method="harvest", is_requiem=req)
        out["dur_" + tag + "spectrogram"] = np.array(dat["spectrogram"])
        out["dur_" + tag + "aperiodicity"] = np.array(dat["aperiodicity"])
        out["dur_" + tag + "f0"] = np.array(dat["f0"])
        out["dur_" + tag + "vuv"] = np.array(dat["vuv"])
        d2 = copy.deepcopy(dat)
        W.modify_duration(d2, [0.3, 0.6], [0.0, 0.2, 0.8, -1])
        out["dur_" + tag + "tp"] = np.array(d2["temporal_positions"])
        refload.reseed(0)
        W.decode(d2)
        out["dur_" + tag + "out"] = np.array(d2["out"])
    np.savez_compressed(os.path.join(HERE, "partial.npz"), **out)
    print({k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
