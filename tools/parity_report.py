"""Parity numbers of the CUDA path against the committed reference goldens (run on the GPU box)."""
import os, random, sys
import numpy as np
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "python-world_b200"))
from world_b200 import main, synthesisRequiem
from world_b200.synthesis import synthesis


def reseed():
    np.random.seed(0); random.seed(0); synthesisRequiem.generate_noise.current_index = None


def rms(a, b):
    return float(np.sqrt(np.mean((a - b) ** 2)))


W = main.World()
g = dict(np.load(os.path.join(ROOT, "tests/golden/syn16k_1s.npz")))
dat = {"f0": g["harvest_d4c_f0"].copy(), "vuv": g["harvest_d4c_vuv"], "fs": 16000,
       "temporal_positions": g["harvest_d4c_temporal_positions"], "spectrogram": g["harvest_d4c_spectrogram"],
       "aperiodicity": g["harvest_d4c_aperiodicity"], "is_requiem": False}
reseed()
y = synthesis(dat, dat)
print("decode-only (golden features, syn16k d4c): waveform RMS diff %.3e" % rms(y, g["harvest_d4c_out"]))
for name in ("syn16k_1s", "mwm_full"):
    g = dict(np.load(os.path.join(ROOT, "tests/golden/%s.npz" % name)))
    x = g["x_int16"] / 32767.0 if "x_int16" in g else g["x"]
    fs = int(g["fs"])
    for tag, method, req in (("dio_d4c_", "dio", False), ("harvest_req_", "harvest", True)):
        reseed()
        d = W.encode(fs, x, f0_method=method, is_requiem=req)
        st = int(g[tag + "frame_stride"])
        v = g[tag + "vuv"] > 0
        S, Sg = d["spectrogram"][:, ::st], g[tag + "spectrogram"]
        m = Sg > 1e-10
        dl = np.abs(np.log10(S[m]) - np.log10(Sg[m]))
        apd = np.max(np.abs(d["aperiodicity"][:, ::st] - g[tag + "aperiodicity"]))
        f0e = np.max(np.abs(d["f0"][v] - g[tag + "f0"][v]) / np.maximum(g[tag + "f0"][v], 1e-9))
        reseed()
        W.decode(d)
        print("%-10s %-13s vuv mismatches %d | voiced f0 max rel %.2e | spectrogram |dlog10| p99 %.2e max %.2e | "
              "aperiodicity max abs %.2e | waveform RMS diff %.3e (signal RMS %.3e)"
              % (name, tag, int(np.sum(d["vuv"] != g[tag + "vuv"])), f0e, np.percentile(dl, 99), dl.max(), apd,
                 rms(d["out"], g[tag + "out"]), float(np.sqrt(np.mean(g[tag + "out"] ** 2)))))
