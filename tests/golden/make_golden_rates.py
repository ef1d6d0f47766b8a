"""Golden vectors at sampling rates around the decimation-ratio edges, by RUNNING the unmodified reference:
8 000 Hz (Harvest passes the signal through, harvest.py:61-63), 11 025 Hz (ratio rounds to 1 but the reference still
runs the zero-phase Chebyshev filter, harvest.py:64-69; DIO ratio 2), 44 100 Hz (ratio 6 / 11).  Harvest + CheapTrick +
D4C and DIO + StoneMask on 0.3 s of the synthetic generator.  Build container only:
    python tests/golden/make_golden_rates.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", "..", "python-world_b200"))
import refload  # noqa: E402
from world_b200 import synth_input  # noqa: E402


def main():
    ref = refload.load()
    out = {}
    for fs in (8000, 11025, 44100):
        x = synth_input.utterance(fs, 0.3, 2, 1)
        tag = "r%d_" % fs
        out[tag + "x"] = x
        refload.reseed(0)
        src = ref.harvest.harvest(np.array(x), fs)
        out[tag + "harvest_f0"] = np.array(src["f0"])
        out[tag + "harvest_vuv"] = np.array(src["vuv"])
        flt = ref.cheaptrick.cheaptrick(np.array(x), fs, src)
        out[tag + "f0_after_cheaptrick"] = np.array(src["f0"])
        out[tag + "spectrogram"] = np.array(flt["spectrogram"][:, ::4])
        src = ref.d4c.d4c(np.array(x), fs, src)
        out[tag + "aperiodicity"] = np.array(src["aperiodicity"][:, ::4])
        d = ref.dio.dio(np.array(x), fs)
        out[tag + "dio_f0"] = np.array(d["f0"])
        out[tag + "dio_vuv"] = np.array(d["vuv"])
        out[tag + "stonemask_f0"] = ref.stonemask.stonemask(np.array(x), fs, d["temporal_positions"], d["f0"])
    np.savez_compressed(os.path.join(HERE, "rates.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
