#!/bin/bash
# One GPU-box pass that regenerates the measured evidence of a round (run through gpurun; outputs land in gpurun_out/):
# bench lines of configs 2-5 (+ reference arm), the ncu launch list of the bench command, and one `ncu --set full`
# capture per config from which tools/ncu_kernels_json.py builds the flop / traffic model bench.py reads.
#   tools/gpu_evidence.sh r02 [quick]
R=${1:-r02}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -3 > $O/${R}_pytest_gpu.txt
cat $O/${R}_pytest_gpu.txt
python tools/parity_report.py > $O/parity_${R}.txt 2>&1
if [ "$2" != "quick" ]; then
# full-set captures, one step of every config at its bench batch; the reports stay on the box (gpurun_out/ is capped
# at 64 MiB), only the per-stage JSON (flop / traffic model of bench.py) and the text summaries come back
T=/tmp/wb_ncu
mkdir -p $T
# start from the committed model so that configs not re-captured in this pass (CAPS) keep their entries
if [ -f profiles/${R}_kernels.json ]; then cp profiles/${R}_kernels.json $O/${R}_kernels.json; else echo '{}' > $O/${R}_kernels.json; fi
CAPS=${CAPS:-"2 3 5 4 4r"}
cap() {  # tag launches batch frames
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled --profile-from-start off \
      -s $2 -c $2 -f -o $T/full_config$1 python tools/profile_config.py $1 $3 > $O/${R}_ncu_$1.log 2>&1
  tail -1 $O/${R}_ncu_$1.log
  python tools/ncu_kernels_json.py $T/full_config$1.ncu-rep config$1 $4 \
      "ncu --set full --clock-control none, second step of tools/profile_config.py $1 $3 (one launch of every kernel)" \
      $O/${R}_kernels.json > $O/${R}_kernels.json.new && mv $O/${R}_kernels.json.new $O/${R}_kernels.json
  python tools/ncu_summary.py $T/full_config$1.ncu-rep > $O/${R}_ncu_full_config$1_summary.txt 2>&1
}
for c in $CAPS; do
  case $c in
    2) cap 2 16 256 205056 ;;
    3) cap 3 16 256 205056 ;;
    5) cap 5 15 128 102528 ;;
    4) cap 4 4 512 410112 ;;
    4r) cap 4r 6 512 410112 ;;
  esac
done
python tools/ncu_lines.py $T/full_config2.ncu-rep d4c 40 > $O/${R}_source_hotspots.txt 2>&1
python tools/ncu_lines.py $T/full_config2.ncu-rep channels_fft 40 >> $O/${R}_source_hotspots.txt 2>&1
ls -la $O | head -40
# the bench lines below take their flop / traffic model from the captures just made
cp $O/${R}_kernels.json profiles/${R}_kernels.json
fi
python bench.py --config 2 --steps 20 --warmup 3 2>$O/${R}_bench_c2.err | tail -1 > $O/${R}_bench_config2.json
for c in 3 5; do
  python bench.py --config $c --steps 5 --warmup 3 2>$O/${R}_bench_c$c.err | tail -1 > $O/${R}_bench_config$c.json
done
python bench.py --config 4 --steps 3 --warmup 3 2>$O/${R}_bench_c4.err | tail -1 > $O/${R}_bench_config4_synthesis.json
python bench.py --config 4 --flavour requiem --steps 3 --warmup 3 2>$O/${R}_bench_c4r.err | tail -1 > $O/${R}_bench_config4_requiem.json
python bench.py --impl reference --steps 3 --warmup 3 2>/dev/null | tail -1 > $O/${R}_bench_reference_config2.json
for f in $O/${R}_bench_config*.json; do python - "$f" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
r = d["roofline"]
print(sys.argv[1].split("/")[-1], "value %.4g  e2e %.4g  ms/step %.2f  top %s %.2f ms  fp64_frac %s" % (
    d["value"], d["e2e"]["value"], d["ms_per_step"], r["kernel"], r["kernel_ms"], r.get("fp64_frac")))
PY
done
[ "$2" = "quick" ] && exit 0
# launch list of the bench command (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 200 --csv --log-file $O/${R}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --streams 1 --no-cpu-baseline --no-e2e-variants > /dev/null 2>&1
python tools/launch_shares.py $O/${R}_launches_bench.csv > $O/${R}_launches_bench_summary.txt
gzip -f $O/${R}_launches_bench.csv
# memory / race checks of the small parity cases (every kernel of encode, decode and the feature heads)
SAN="tests/test_gpu_features.py tests/test_gpu_harvest.py::test_harvest_gpu_syn16k tests/test_gpu_harvest.py::test_dio_stonemask_gpu tests/test_gpu_spectral.py tests/test_gpu_decode.py::test_batch_decode_device_noise tests/test_gpu_pipeline.py::test_fused_encode_equals_stages tests/test_gpu_pipeline.py::test_coarse_transport"
timeout 900 compute-sanitizer --tool memcheck python -m pytest $SAN -q -x 2>&1 | tail -15 > $O/${R}_sanitizer_memcheck.log
tail -2 $O/${R}_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_harvest.py::test_harvest_gpu_syn16k tests/test_gpu_features.py tests/test_gpu_spectral.py -q -x 2>&1 | tail -15 > $O/${R}_sanitizer_racecheck.log
tail -2 $O/${R}_sanitizer_racecheck.log
# one full-length (4 s) utterance of config 2: the overlap-save kernel's interpolation tables then run on into the
# idle spectrum buffer and the longest streams go through more than one table window
timeout 900 compute-sanitizer --tool racecheck python tools/profile_config.py 2 1 2>&1 | tail -6 > $O/${R}_sanitizer_racecheck_4s.log
tail -2 $O/${R}_sanitizer_racecheck_4s.log
timeout 900 compute-sanitizer --tool memcheck python tools/profile_config.py 2 1 2>&1 | tail -6 > $O/${R}_sanitizer_memcheck_4s.log
tail -2 $O/${R}_sanitizer_memcheck_4s.log
