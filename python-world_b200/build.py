"""Build the in-tree CUDA library for sm_100a:  python python-world_b200/build.py

nvcc cross-compiles without a GPU.  The .so lands next to the Python package
(git-ignored, but it travels to the GPU box with the gpurun snapshot)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "world_b200", "libworld_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-shared"]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def up_to_date():
    if not os.path.exists(OUT):
        return False
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "world_b200.h")]
    return all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps)


def build(force=False, verbose=False, out=None, defines=()):
    """out / defines: a variant build (tests of rarely taken paths, tuning sweeps) next to the product library."""
    if out is None and not force and up_to_date():
        return OUT
    cmd = [NVCC] + FLAGS + list(defines) + (["-Xptxas", "-v"] if verbose else []) + ["-o", out or OUT] + sources()
    print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return out or OUT


if __name__ == "__main__":
    out = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else None
    build(force="--force" in sys.argv, verbose="-v" in sys.argv, out=out,
          defines=[a for a in sys.argv[1:] if a.startswith("-D")])
