"""Per-kernel timing on the GPU (CUDA events), config-2 shape by default."""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "python-world_b200"))
from world_b200 import engine as eng, synth_input  # noqa: E402


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    return sorted(a.elapsed_time(b) for a, b in ev)[len(ev) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--fs", type=int, default=16000)
    ap.add_argument("--seconds", type=float, default=4.0)
    a = ap.parse_args()
    E = eng.default_engine(0)
    fs = a.fs
    uniq = min(a.batch, 8)
    xs = synth_input.batch(fs, a.seconds, 2, uniq)
    x = np.concatenate([xs] * (a.batch // uniq), axis=0)
    B, S = x.shape
    F = int(1000 * S / fs / 5 + 1)
    t = np.arange(F) * 0.005
    rng = np.random.default_rng(0)
    f0 = 140 * 2 ** (0.3 * np.sin(2 * np.pi * 0.7 * t[None] + rng.random((B, 1)) * 6))
    vuv = (np.sin(2 * np.pi * 1.1 * t[None] + rng.random((B, 1)) * 6) > -0.5).astype(np.float64)
    X, ns, T = E.f64(x), E.i32([S] * B), E.f64(np.tile(t, (B, 1)))
    F0, V, nf = E.f64(f0), E.f64(vuv), E.i32([F] * B)
    frames = B * F
    print("batch %d, frames %d, voiced %.2f" % (B, frames, vuv.mean()))
    ms = timeit(lambda: E.cheaptrick(X, ns, fs, T, F0, V, nf))
    print("cheaptrick        %8.3f ms  %10.0f frames/s" % (ms, frames / ms * 1e3))
    ms = timeit(lambda: E.cheaptrick(X, ns, fs, T, F0, V, nf, want_ps=True))
    print("cheaptrick+ps     %8.3f ms  %10.0f frames/s" % (ms, frames / ms * 1e3))
    f0u, _, _ = E.cheaptrick(X, ns, fs, T, F0, V, nf)
    ms = timeit(lambda: E.d4c(X, ns, fs, T, f0u, V, nf))
    print("d4c               %8.3f ms  %10.0f frames/s" % (ms, frames / ms * 1e3))
    ms = timeit(lambda: E.d4c_requiem(X, ns, fs, T, f0u, V, nf))
    print("d4c_requiem       %8.3f ms  %10.0f frames/s" % (ms, frames / ms * 1e3))


if __name__ == "__main__":
    main()
