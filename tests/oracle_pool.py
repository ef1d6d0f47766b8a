"""Test helper: run oracle/pipeline.encode for several utterances in worker processes (spawned, so that they never
inherit a CUDA context).  Test infrastructure only."""
import multiprocessing as mp
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _work(job):
    for p in (ROOT, os.path.join(ROOT, "python-world_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import numpy as np
    from oracle import pipeline
    from world_b200 import synth_input
    fs, seconds, config, index, is_requiem = job
    x = synth_input.utterance(fs, seconds, config, index)
    np.random.seed(0)  # CheapTrick's eps-dither consumes np.random (cheaptrick.py:117)
    d = pipeline.encode(fs, x, "harvest", is_requiem=is_requiem)
    return {k: d[k] for k in ("temporal_positions", "vuv", "f0", "spectrogram", "aperiodicity")}


def encode_many(jobs, workers=None):
    """jobs: (fs, seconds, config, utterance index, is_requiem) tuples -> list of oracle encode dicts."""
    workers = workers or max(1, min(len(jobs), (os.cpu_count() or 2) // 2))
    with mp.get_context("spawn").Pool(workers) as pool:
        return pool.map(_work, jobs)
