"""Golden vectors for the feature heads / edits by RUNNING the unmodified reference's
World.encode_lfbank / encode_mcep / decode_mcep / warp_spectrum / modify_duration on spectrograms taken from
the committed fixtures.  Build container only:  python tests/golden/make_golden_features.py"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refload  # noqa: E402


def main():
    refload.load()
    W = importlib.import_module("refworld.main").World()
    g16 = dict(np.load(os.path.join(HERE, "syn16k_1s.npz")))
    g48 = dict(np.load(os.path.join(HERE, "syn48k_05s.npz")))
    out = {}
    for tag, spec, fs in (("k16_", g16["harvest_d4c_spectrogram"][:, ::5], 16000), ("k48_", g48["spectrogram"][:, ::2], 48000)):
        mag = np.ascontiguousarray(spec.T)  # the reference's callers pass data['spectrogram'].T (test/spectralFeatures.py:27)
        out[tag + "spec"] = mag
        out[tag + "lfbank"] = W.encode_lfbank(mag.copy(), fs=fs)
        out[tag + "lfbank_20_hi"] = W.encode_lfbank(mag.copy(), prefac=0.9, fs=fs, nfilt=20, lowfreq=100, highfreq=fs / 2 - 500)
        out[tag + "mcep40"] = W.encode_mcep(mag.copy(), n0=40, fs=fs, highhz=min(8000, fs // 2))
        out[tag + "mcep12"] = W.encode_mcep(mag.copy(), fs=fs)
        nfft = (mag.shape[1] - 1) * 2
        out[tag + "decoded"] = W.decode_mcep(out[tag + "mcep40"].copy(), fft_size=nfft)
        for f in (0.8, 1.25):
            d = {"spectrogram": spec.copy()}
            W.warp_spectrum(d, f)
            out[tag + "warp_%g" % f] = d["spectrogram"]
    tp = np.arange(0, 401) * 0.005
    d = {"temporal_positions": tp.copy()}
    W.modify_duration(d, [1, 1.5], [0, 1, 3, -1])
    out["dur_tp"] = tp
    out["dur_out"] = d["temporal_positions"]
    np.savez_compressed(os.path.join(HERE, "features.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
