#!/bin/bash
# per-frame spectral kernels against their block size (WB_D4C_THREADS / WB_CT_THREADS tuning knobs)
run() { python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['roofline']['stage_ms']; print(round(d['ms_per_step'],2), 'cheaptrick', round(s['cheaptrick'],2), 'd4c', round(s['d4c'],2))"; }
echo default; run
for t in 128 512; do echo "WB_D4C_THREADS=$t"; WB_D4C_THREADS=$t run; done
for t in 128 512; do echo "WB_CT_THREADS=$t"; WB_CT_THREADS=$t run; done
