#!/bin/bash
# One `ncu --set full` source-level capture of the kernels matching REGEX on the config-2 workload (batch 64):
#   tools/gpu_ncu_one.sh TAG REGEX [launch-skip] [launch-count]
T=${1:-one}
O=gpurun_out
mkdir -p $O
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    --kernel-name regex:"$2" -s ${3:-1} -c ${4:-1} -f -o $O/${T}_src \
    python tools/profile_config.py 2 64 > $O/${T}_ncu.log 2>&1
tail -2 $O/${T}_ncu.log
