"""Decode throughput (BASELINE config 4 shape: synthesis from precomputed features), per flavour."""
import argparse, os, sys, time
import numpy as np
import torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "python-world_b200"))
from world_b200 import engine as eng, synth_input, get_seeds_signals

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--iters", type=int, default=3)
a = ap.parse_args()
E = eng.default_engine(0)
fs = 16000
uniq = 16
xs = synth_input.batch(fs, 4.0, 4, uniq)
X = E.f64(xs); ns = E.i32([xs.shape[1]] * uniq)
for req in (False, True):
    d = E.encode(X, ns, fs, is_requiem=req)
    rep = a.batch // uniq
    tile = lambda t: t.repeat((rep,) + (1,) * (t.dim() - 1)).contiguous()
    tp, f0, vuv, spec, apx, nf = (tile(d[k]) for k in ("temporal_positions", "f0", "vuv", "spectrogram", "aperiodicity", "n_frames"))
    B, F = tp.shape
    ylen = E.synthesis_length(0.0, float(tp[0, -1]), fs)
    if req:
        sd = get_seeds_signals.get_seeds_signals(fs)
        P, N = E.f64(sd["pulse"]), E.f64(sd["noise"])
        fn = lambda: E.synthesis_requiem(tp, f0, vuv, spec, apx, nf, fs, ylen, P, N)
    else:
        fn = lambda: E.synthesis(tp, f0, vuv, spec, apx, nf, fs, ylen, noise="device", seed=1)
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(a.iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    print("decode %-9s batch %d: %.2f ms  %.0f frames/s  %.1f x realtime-seconds/s" %
          ("requiem" if req else "synthesis", B, ms, B * F / ms * 1e3, B * 4.0 / ms * 1e3))
