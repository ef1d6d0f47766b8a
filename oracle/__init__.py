"""CPU oracle for the WORLD encode()/decode() hot path -- TEST INFRASTRUCTURE.

A NumPy restatement of the reference algorithm (tuanad121/Python-WORLD,
world/*.py), pinned against outputs of the unmodified reference run in the
build container (tests/golden/*.npz, see tests/golden/make_golden.py).

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline and
--impl reference) may import this package, and only as the checker / CPU
baseline.  The product path (python-world_b200/) never does.
"""
