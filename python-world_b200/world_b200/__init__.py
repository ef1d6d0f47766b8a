"""world_b200 -- B200-native WORLD vocoder analysis/synthesis engine.

Drop-in for the encode()/decode() hot path of tuanad121/Python-WORLD
(`from world_b200 import main; main.World().encode(fs, x, ...)`), implemented as
hand-written sm_100a CUDA kernels behind a C-ABI (include/world_b200.h) that this
package calls through ctypes.  torch tensors are used only as HBM containers.
There is no CPU fallback: importing the engine without the compiled CUDA library
or without a GPU raises.
"""
__all__ = ["main", "engine"]
