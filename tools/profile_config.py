"""One warm-up step and one step of a BASELINE config for ncu (kernels launched on one stream, in order).
usage: profile_config.py <config 2|3|4|4r|5> <batch>
Launches per step: config 2 / 3: 16, config 5: 15, config 4: 4 (synthesis) / 6 (requiem, '4r') -- so the second
step is selected with  ncu --profile-from-start off -s <launches> -c <launches>."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "python-world_b200"))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from world_b200 import engine as eng  # noqa: E402

tag = sys.argv[1]
cfg = bench.CONFIGS[int(tag[0])]
B = int(sys.argv[2]) if len(sys.argv) > 2 else cfg["batch"]
E = eng.default_engine(0)
fs = cfg["fs"]
if cfg["kind"] == "encode":
    xs = bench.make_inputs(cfg, 0, B)
    X, ns = E.f64(xs), E.i32([xs.shape[1]] * B)
    step = lambda: E.encode(X, ns, fs, f0_method="harvest", is_requiem=cfg["requiem"], aperiodicity=cfg["aperiodicity"])
else:
    req = tag.endswith("r")
    src = bench.make_inputs(cfg, 0, 32)
    d0 = E.encode(E.f64(src), E.i32([src.shape[1]] * 32), fs, f0_method="harvest", is_requiem=req)
    reps = (B + 31) // 32
    f = {k: d0[k].repeat((reps,) + (1,) * (d0[k].dim() - 1))[:B].contiguous()
         for k in ("temporal_positions", "f0", "vuv", "spectrogram", "aperiodicity", "n_frames")}
    F = bench.frames_of(cfg)
    ylen = E.synthesis_length(0.0, (F - 1) * 5.0 / 1000.0, fs)
    seeds = None
    if req:
        from world_b200.get_seeds_signals import get_seeds_signals
        np.random.seed(0)
        sd = get_seeds_signals(fs)
        seeds = (E.f64(sd["pulse"]), E.f64(sd["noise"]))
    step = lambda: E.decode(f["temporal_positions"], f["f0"], f["vuv"], f["spectrogram"], f["aperiodicity"], f["n_frames"],
                            fs, ylen, is_requiem=req, seeds=seeds, seed=1)
torch.cuda.synchronize()
torch.cuda.profiler.start()  # ncu --profile-from-start off: launches are counted from here
for _ in range(2):
    step()
torch.cuda.synchronize()
