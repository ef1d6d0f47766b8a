#!/bin/bash
# One GPU-box pass that regenerates the evidence under profiles/ (run through gpurun; outputs land in gpurun_out/).
#   tools/round_evidence.sh r01
R=${1:-r01}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -3 > $O/${R}_pytest_gpu.txt
python bench.py --steps 5 --warmup 3 2>/dev/null | tail -1 > $O/${R}_bench_n1.json
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > $O/${R}_bench_reference_n1.json
python tools/parity_report.py > $O/parity_${R}.txt 2>&1
python tools/bench_decode.py > $O/${R}_decode_batch256.txt 2>&1
# launch list of the bench command (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 240 --csv --log-file $O/${R}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --streams 1 --no-cpu-baseline > /dev/null 2>&1
# one full-set capture of every kernel of encode() at the bench batch
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -s 16 -c 16 -f -o $O/${R}_full_batch256 \
    python tools/profile_encode.py 256 > $O/ncu_full.log 2>&1
tail -2 $O/ncu_full.log
cat $O/${R}_pytest_gpu.txt
cat $O/${R}_bench_n1.json
# memory / race checks of the small parity cases (every kernel of encode, decode and the feature heads)
SAN="tests/test_gpu_features.py tests/test_gpu_harvest.py::test_harvest_gpu_syn16k tests/test_gpu_harvest.py::test_dio_stonemask_gpu tests/test_gpu_spectral.py tests/test_gpu_decode.py::test_batch_decode_device_noise"
timeout 600 compute-sanitizer --tool memcheck python -m pytest $SAN -q -x > $O/${R}_sanitizer_memcheck.log 2>&1
tail -2 $O/${R}_sanitizer_memcheck.log
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_harvest.py::test_harvest_gpu_syn16k tests/test_gpu_features.py tests/test_gpu_spectral.py -q -x > $O/${R}_sanitizer_racecheck.log 2>&1
tail -2 $O/${R}_sanitizer_racecheck.log
