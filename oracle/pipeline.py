"""encode() of the reference restated on the oracle stages (main.py:106-152).

TEST INFRASTRUCTURE / CPU baseline only (see oracle/__init__.py)."""
import numpy as np

from . import cheaptrick as o_ct
from . import d4c as o_d4c
from . import harvest as o_hv


def encode(fs, x, f0_method="harvest", f0_floor=71, f0_ceil=800, frame_period=5, fft_size=None, is_requiem=False,
           dither="legacy"):
    if fft_size is not None:
        f0_floor = 3.0 * fs / fft_size
    if f0_method == "harvest":
        src = o_hv.harvest(x, fs, f0_floor, f0_ceil, frame_period)
    elif f0_method == "dio":
        from . import dio as o_dio
        src = o_dio.dio(x, fs, f0_floor, f0_ceil, frame_period=frame_period)
        src["f0"] = o_dio.stonemask(x, fs, src["temporal_positions"], src["f0"])
    else:
        raise Exception
    ct = o_ct.cheaptrick(x, fs, src["temporal_positions"], src["f0"], src["vuv"], fft_size=fft_size, dither=dither)
    if is_requiem:
        ap = o_d4c.d4c_requiem(x, fs, src["temporal_positions"], ct["f0"], src["vuv"], fft_size=fft_size)
    else:
        ap = o_d4c.d4c(x, fs, src["temporal_positions"], ct["f0"], src["vuv"], fft_size_for_spectrum=fft_size)
    return {"temporal_positions": src["temporal_positions"], "vuv": src["vuv"], "fs": fs, "f0": ap["f0"],
            "aperiodicity": ap["aperiodicity"], "ps spectrogram": ct["ps_spectrogram"],
            "spectrogram": ct["spectrogram"], "is_requiem": is_requiem}


def encode_quiet(fs, x):
    """Worker for the multi-process CPU baseline: full Harvest+CheapTrick+D4C, result discarded."""
    d = encode(fs, x, "harvest")
    return len(d["f0"])
