#!/bin/bash
# bench line of every library variant under gpurun_variants/ (built with build.py --out or by hand): stage times side by side
#   tools/sweep_libs.sh [config]
C=${1:-2}
for so in gpurun_variants/lib_*.so; do
  WORLD_B200_LIB=$PWD/$so python bench.py --config $C --steps 5 --warmup 3 --no-cpu-baseline --no-e2e-variants 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['roofline']['stage_ms']
print('$so', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']/1e6,3), ' '.join('%s %.2f'%(k,v) for k,v in s.items()))"
done
