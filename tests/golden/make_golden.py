"""Generate the committed golden fixtures by RUNNING the unmodified reference.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py

The reference has no golden vectors of its own, so parity
is pinned on outputs of the reference executed here.  Every call is preceded by
refload.reseed(0) (np.random, random, generate_noise cursor) because encode
consumes np.random through CheapTrick's eps-dither (cheaptrick.py:117) and decode
through randn / the velvet-noise seeds (synthesis.py:93, get_seeds_signals.py:48-72).

Fixtures (np.savez_compressed, float64 unless noted):
  mwm_full.npz     test/test-mwm.wav (22 050 Hz, int16 samples stored as int16),
                   BASELINE config 1 (dio+stonemask, d4c, synthesis) and the
                   example/prosody.py path (harvest, requiem); per-frame matrices
                   are kept at every 8th frame to stay small, F0/vuv/out in full.
  syn16k_1s.npz    synthetic 16 kHz 1 s (config-2 shape): stage-level Harvest
                   intermediates, CheapTrick, D4C, D4C-Requiem, both decoders.
  syn48k_05s.npz   synthetic 48 kHz 0.5 s (config-5 shape): Harvest + CheapTrick.
"""
import copy
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", "..", "python-world_b200"))

import refload  # noqa: E402
from world_b200 import synth_input  # noqa: E402

STRIDE = 8


def _copy_source(src):
    return {k: (np.array(v) if isinstance(v, np.ndarray) else v) for k, v in src.items()}


def encode_stages(ref, fs, x, f0_method, is_requiem, out, tag):
    """main.py:106-152 unrolled so every stage boundary is captured."""
    dio, sm, hv = ref.dio, ref.stonemask, ref.harvest
    refload.reseed(0)
    if f0_method == "dio":
        src = dio.dio(x, fs)
        out[tag + "dio_f0"] = np.array(src["f0"])
        out[tag + "dio_raw_f0_candidates"] = np.array(src["raw_f0_candidates"])
        out[tag + "dio_f0_candidates"] = np.array(src["f0_candidates"])
        src["f0"] = sm.stonemask(x, fs, src["temporal_positions"], src["f0"])
    else:
        src = hv.harvest(x, fs)
    out[tag + "f0_tracker"] = np.array(src["f0"])
    out[tag + "vuv"] = np.array(src["vuv"])
    out[tag + "temporal_positions"] = np.array(src["temporal_positions"])
    flt = ref.cheaptrick.cheaptrick(x, fs, src)
    out[tag + "f0_after_cheaptrick"] = np.array(src["f0"])
    if is_requiem:
        src = ref.d4cRequiem.d4cRequiem(x, fs, src)
    else:
        src = ref.d4c.d4c(x, fs, src)
        out[tag + "coarse_ap"] = np.array(src["coarse_ap"])
    dat = {"temporal_positions": src["temporal_positions"], "vuv": src["vuv"], "fs": fs,
           "f0": src["f0"], "aperiodicity": src["aperiodicity"],
           "ps spectrogram": flt["ps spectrogram"], "spectrogram": flt["spectrogram"],
           "is_requiem": is_requiem}
    out[tag + "f0"] = np.array(dat["f0"])
    return dat


def decode(ref, dat):
    """main.py:198-214."""
    refload.reseed(0)
    d = copy.deepcopy(dat)
    if d["is_requiem"]:
        seeds = ref.get_seeds_signals.get_seeds_signals(d["fs"])
        y = ref.synthesisRequiem.synthesisRequiem(d, d, seeds)
    else:
        seeds = None
        y = ref.synthesis.synthesis(d, d)
    m = np.max(np.abs(y))
    raw = np.array(y)
    if m > 1.0:
        y = y / m
    return y, raw, seeds


def harvest_stages(ref, fs, x, out, tag):
    """harvest.py:17-54 unrolled."""
    hv = ref.harvest
    f0_floor, f0_ceil = 71, 800
    n1 = int(1000 * len(x) / fs / 1 + 1)
    tp1 = np.arange(0, n1) * 1 / 1000
    fl, ce = f0_floor * 0.9, f0_ceil * 1.1
    bl = np.arange(np.ceil(np.log2(ce / fl) * 40)) + 1
    bl = fl * 2.0 ** (bl / 40)
    y, afs = hv.CalculateDownsampledSignal(np.array(x), fs, 8000)
    fft_size = int(2 ** np.ceil(np.log2(len(y) + int(fs / fl * 4 + 0.5) + 1)))
    ysp = np.fft.fft(y, fft_size)
    raw = hv.CalculateCandidates(len(tp1), bl, len(y), tp1, afs, ysp, f0_floor, f0_ceil)
    cand, ncand = hv.DetectCandidates(raw)
    ov = hv.OverlapF0Candidates(cand, ncand)
    rf, rs = hv.RefineCandidates(y, afs, tp1, ov, f0_floor, f0_ceil)
    uf, us = hv.RemoveUnreliableCandidates(rf, rs)
    conn, vuv = hv.FixF0Contour(uf, us)
    sm = hv.SmoothF0(conn)
    out[tag + "hv_y"] = y
    out[tag + "hv_actual_fs"] = np.float64(afs)
    out[tag + "hv_raw"] = raw
    out[tag + "hv_detect"] = cand
    out[tag + "hv_ncand"] = np.int64(ncand)
    out[tag + "hv_refined_f0"] = rf
    out[tag + "hv_refined_score"] = rs
    out[tag + "hv_reliable_f0"] = uf
    out[tag + "hv_reliable_score"] = us
    out[tag + "hv_connected"] = conn
    out[tag + "hv_smoothed"] = sm


def put_matrices(out, tag, dat, stride=STRIDE):
    out[tag + "frame_stride"] = np.int64(stride)
    out[tag + "spectrogram"] = np.array(dat["spectrogram"][:, ::stride])
    out[tag + "aperiodicity"] = np.array(dat["aperiodicity"][:, ::stride])
    out[tag + "ps_spectrogram"] = np.array(dat["ps spectrogram"][:, ::stride * 4])


def main():
    ref = refload.load()
    from scipy.io import wavfile

    # ---- test-mwm.wav, full file --------------------------------------------
    fs, xi = wavfile.read(os.path.join(refload.REF_ROOT, "test", "test-mwm.wav"))
    x = xi / 32767.0  # test/speed.py:14
    out = {"fs": np.int64(fs), "x_int16": xi.astype(np.int16)}
    for tag, method, req in (("dio_d4c_", "dio", False), ("harvest_req_", "harvest", True)):
        dat = encode_stages(ref, fs, np.array(x), method, req, out, tag)
        put_matrices(out, tag, dat)
        y, raw, seeds = decode(ref, dat)
        out[tag + "out"] = y
        out[tag + "out_peak_before_rescale"] = np.float64(np.max(np.abs(raw)))
    np.savez_compressed(os.path.join(HERE, "mwm_full.npz"), **out)
    print("mwm_full", {k: getattr(v, "shape", None) for k, v in out.items()})

    # ---- synthetic 16 kHz 1 s -------------------------------------------------
    fs = 16000
    x = synth_input.utterance(fs, 1.0, 2, 0)
    out = {"fs": np.int64(fs), "x": x}
    harvest_stages(ref, fs, x, out, "")
    for tag, method, req in (("harvest_d4c_", "harvest", False), ("harvest_req_", "harvest", True),
                             ("dio_d4c_", "dio", False)):
        dat = encode_stages(ref, fs, np.array(x), method, req, out, tag)
        put_matrices(out, tag, dat, stride=1 if tag == "harvest_d4c_" else 4)
        y, raw, seeds = decode(ref, dat)
        out[tag + "out"] = y
        if seeds is not None:
            out[tag + "seed_pulse"] = seeds["pulse"]
            out[tag + "seed_noise"] = seeds["noise"]
    np.savez_compressed(os.path.join(HERE, "syn16k_1s.npz"), **out)
    print("syn16k_1s", {k: getattr(v, "shape", None) for k, v in out.items()})

    # ---- synthetic 48 kHz 0.5 s ------------------------------------------------
    fs = 48000
    x = synth_input.utterance(fs, 0.5, 5, 0)
    out = {"fs": np.int64(fs), "x": x}
    refload.reseed(0)
    src = ref.harvest.harvest(np.array(x), fs)
    out["f0_tracker"] = np.array(src["f0"])
    out["vuv"] = np.array(src["vuv"])
    flt = ref.cheaptrick.cheaptrick(x, fs, src)
    out["f0_after_cheaptrick"] = np.array(src["f0"])
    out["spectrogram"] = np.array(flt["spectrogram"][:, ::4])
    out["frame_stride"] = np.int64(4)
    np.savez_compressed(os.path.join(HERE, "syn48k_05s.npz"), **out)
    print("syn48k_05s", {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
