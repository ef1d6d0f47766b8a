"""Turn what tools/round_evidence.sh left in gpurun_out/ into the tracked files under profiles/ (run here, after the
GPU-box pass):  python tools/collect_profiles.py r01"""
import collections, csv, gzip, json, os, shutil, subprocess, sys
R = sys.argv[1] if len(sys.argv) > 1 else "r01"
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
rep = os.path.join(G, "%s_full_batch256.ncu-rep" % R)
py = sys.executable
open(os.path.join(P, "%s_ncu_full_batch256_summary.txt" % R), "w").write(
    subprocess.run([py, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout)
how = ("ncu --set full --clock-control none --import-source on -s 16 -c 16, python tools/profile_encode.py 256 (second "
       "encode() of batch 256 x 16 kHz x 4 s; one launch of every kernel); tools/round_evidence.sh")
open(os.path.join(P, "%s_kernels.json" % R), "w").write(
    subprocess.run([py, os.path.join(ROOT, "tools", "ncu_kernels_json.py"), rep, how], capture_output=True, text=True).stdout)
# launch list of the bench command
src = os.path.join(G, "%s_launches_bench.csv" % R)
rows = list(csv.reader(l for l in open(src) if not l.startswith("==")))
h = rows[0]
ix = {k: i for i, k in enumerate(h)}
agg = collections.defaultdict(list)
for r in rows[1:]:
    if len(r) != len(h) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[ix["Metric Value"]].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ix["Metric Unit"]], 1e-6)
    agg[r[ix["Kernel Name"]]].append(v)
tot = sum(sum(v) for v in agg.values())
n = sum(len(v) for v in agg.values())
with open(os.path.join(P, "%s_launches_bench_summary.txt" % R), "w") as f:
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 240  python bench.py --steps 2 --warmup 3 --streams 1 --no-cpu-baseline\n")
    f.write("(cold-cache, serialised launches: compare SHARES)  total %.1f ms over %d launches\n" % (tot, n))
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        f.write("%6.2f%% %10.3f ms %5d launches  avg %8.3f ms  %s\n" % (100 * sum(v) / tot, sum(v), len(v), sum(v) / len(v), k[:70]))
with open(src, "rb") as a, gzip.open(os.path.join(P, "%s_launches_bench.csv.gz" % R), "wb") as b:
    shutil.copyfileobj(a, b)
for name in ("%s_bench_n1.json", "%s_bench_reference_n1.json", "parity_%s.txt", "%s_decode_batch256.txt", "%s_pytest_gpu.txt",
             "%s_sanitizer_memcheck.log", "%s_sanitizer_racecheck.log", "%s_bench_n2.json"):
    p = os.path.join(G, name % R)
    if os.path.exists(p):
        shutil.copy(p, os.path.join(P, name % R))
k = json.load(open(os.path.join(P, "%s_kernels.json" % R)))
for name, v in k.items():
    if name != "_source":
        print("%-20s %7.2f ms  %5.2f GB dram  fp64 %4.1f%%  issue %4.1f%%" % (name, v["time_ms"], v["dram_bytes"] / 1e9, v["fp64_pipe_pct"], v["issue_pct"]))
