#!/bin/bash
# One GPU-box iteration: parity tests, bench line, and (optionally) an ncu source-level capture of the spectral kernels.
#   tools/gpu_iter.sh TAG [ncu]
T=${1:-iter}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $O/${T}_pytest.txt
python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>$O/${T}_bench.err | tail -1 > $O/${T}_bench.json
python - <<PY
import json
d = json.load(open("$O/${T}_bench.json"))
print("ms_per_step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"])
print(d["roofline"]["stage_ms"])
PY
if [ "$2" = "ncu" ]; then
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
      --kernel-name regex:'d4c|cheaptrick|channels_fft|refine_items' --profile-from-start off -s 6 -c 6 -f -o $O/${T}_src \
      python tools/profile_config.py 2 64 > $O/${T}_ncu.log 2>&1
  tail -2 $O/${T}_ncu.log
fi
