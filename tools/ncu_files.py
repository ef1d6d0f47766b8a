"""Samples / instructions of one kernel aggregated per source file and per 10-line bucket.
usage: ncu_files.py report.ncu-rep <kernel-id>"""
import csv, subprocess, sys, collections
rep, kid = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-id", ":::" + kid],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; hdr = None; agg = collections.Counter(); aggi = collections.Counter(); b = collections.Counter(); bi = collections.Counter()
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; S = hdr.index("# Samples"); I = hdr.index("Instructions Executed"); continue
    if hdr and r[0] != "" and len(r) == len(hdr):
        try: ln = int(r[0])
        except ValueError: continue
        agg[cur] += int(r[S]); aggi[cur] += int(r[I])
        b[(cur, ln // 10 * 10)] += int(r[S]); bi[(cur, ln // 10 * 10)] += int(r[I])
ts = sum(agg.values()); ti = sum(aggi.values())
for f, s in agg.most_common(): print("%5.1f%% smp %5.1f%% ins  %s" % (100.0 * s / ts, 100.0 * aggi[f] / ti, f))
print()
for (f, l), s in b.most_common(40): print("%5.1f%% smp %5.1f%% ins  %s:%d-%d" % (100.0 * s / ts, 100.0 * bi[(f, l)] / ti, f, l, l + 9))
