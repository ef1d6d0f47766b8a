"""Share of every kernel in an `ncu --metrics gpu__time_duration.sum --csv` launch list (cold-cache, serialised
launches: compare SHARES with bench.py's stage_ms, not absolute times).  usage: launch_shares.py launches.csv"""
import collections
import csv
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if not l.startswith("==")))
h = rows[0]
ix = {k: i for i, k in enumerate(h)}
agg = collections.defaultdict(list)
for r in rows[1:]:
    if len(r) != len(h) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ix["Metric Unit"]], 1e-6)
    agg[r[ix["Kernel Name"]]].append(float(r[ix["Metric Value"]].replace(",", "")) * scale)
tot = sum(sum(v) for v in agg.values())
print("total %.1f ms over %d launches" % (tot, sum(len(v) for v in agg.values())))
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print("%6.2f%% %10.3f ms %5d launches  avg %8.3f ms  %s" % (100 * sum(v) / tot, sum(v), len(v), sum(v) / len(v), k[:90]))
