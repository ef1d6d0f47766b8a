#!/bin/bash
# channels kernel time against the direct-FIR / overlap-save threshold (WB_HV_FFT_MIN_TAPS)
for t in ${@:-100000 200 128 96 64 32}; do
  echo "min_taps=$t"
  WB_HV_FFT_MIN_TAPS=$t python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['stage_ms'])"
done
