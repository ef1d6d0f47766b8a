"""Throughput of the other BASELINE configs on one GPU (their per-GPU shard), CUDA events, inputs resident:
  config 3  256 x 16 kHz x 4 s, full encode with is_requiem=True (Harvest + CheapTrick + D4C-Requiem)
  config 4  decode only, batch 4096 (both flavours)
  config 5  128 x 48 kHz x 4 s, Harvest + CheapTrick (FFT 2048)
(config 2 is bench.py; config 1 is the reference's own CPU case, see profiles/parity_*.txt)."""
import os, sys
import numpy as np
import torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "python-world_b200"))
from world_b200 import engine as eng, synth_input, get_seeds_signals

E = eng.default_engine(0)


def timed(fn, iters=3):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


def tile(t, rep):
    return t.repeat((rep,) + (1,) * (t.dim() - 1)).contiguous()


# config 3 shard
fs, B = 16000, 256
xs = synth_input.batch(fs, 4.0, 3, B)
X, ns = E.f64(xs), E.i32([xs.shape[1]] * B)
ms = timed(lambda: E.encode(X, ns, fs, f0_method="harvest", is_requiem=True, streams=2))
print("config 3 shard (256 x 16 kHz x 4 s, requiem encode): %.1f ms  %.2f M frames/s" % (ms, B * 801 / ms / 1e3))
del X
# config 5 shard
fs, B = 48000, 128
xs = synth_input.batch(fs, 4.0, 5, B)
X, ns = E.f64(xs), E.i32([xs.shape[1]] * B)


def c5():
    tp, f0, vuv, nf = E.harvest(X, ns, fs)
    E.cheaptrick(X, ns, fs, tp, f0, vuv, nf)


ms = timed(c5)
print("config 5 shard (128 x 48 kHz x 4 s, Harvest + CheapTrick): %.1f ms  %.2f M frames/s" % (ms, B * 801 / ms / 1e3))
del X
# config 4
fs, uniq, B = 16000, 32, 4096
xs = synth_input.batch(fs, 4.0, 4, uniq)
X, ns = E.f64(xs), E.i32([xs.shape[1]] * uniq)
for req in (False, True):
    d = E.encode(X, ns, fs, is_requiem=req)
    tp, f0, vuv, spec, apx, nf = (tile(d[k], B // uniq) for k in ("temporal_positions", "f0", "vuv", "spectrogram", "aperiodicity", "n_frames"))
    ylen = E.synthesis_length(0.0, float(tp[0, -1]), fs)
    if req:
        sd = get_seeds_signals.get_seeds_signals(fs)
        P, N = E.f64(sd["pulse"]), E.f64(sd["noise"])
        fn = lambda: E.synthesis_requiem(tp, f0, vuv, spec, apx, nf, fs, ylen, P, N)
    else:
        fn = lambda: E.synthesis(tp, f0, vuv, spec, apx, nf, fs, ylen, noise="device", seed=1)
    ms = timed(fn, iters=2)
    print("config 4 (decode, batch 4096, %s): %.1f ms  %.2f M frames/s  %.0f x real time" %
          ("requiem" if req else "synthesis", ms, B * 801 / ms / 1e3, B * 4.0 / ms * 1e3))
    del tp, f0, vuv, spec, apx
    torch.cuda.empty_cache()
