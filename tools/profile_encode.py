"""Small encode() run for ncu (one launch of every kernel after a warm-up)."""
import os, sys
import numpy as np
import torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "python-world_b200"))
from world_b200 import engine as eng, synth_input
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
req = len(sys.argv) > 2 and sys.argv[2] == "req"
E = eng.default_engine(0)
xs = synth_input.batch(16000, 4.0, 2, B)
X = E.f64(xs); ns = E.i32([xs.shape[1]] * B)
for _ in range(2):
    E.encode(X, ns, 16000, is_requiem=req)
torch.cuda.synchronize()
