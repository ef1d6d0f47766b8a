"""Per-source-line hot spots of one kernel from an ncu report (needs -lineinfo + --import-source on).
usage: ncu_lines.py report.ncu-rep <kernel name fragment> [top-n]
Prints the stall-reason mix of the launch and the lines with the most warp-state samples / executed instructions."""
import csv
import subprocess
import sys

rep, frag = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[0]
names = [r[h.index("Kernel Name")] for r in rows[2:]]
idx = next((i for i, k in enumerate(names) if frag in k), None)
if idx is None:
    sys.exit("no kernel matching %r in %s" % (frag, names))
r = rows[2 + idx]
stalls = [(float(r[i]), h[i]) for i in range(len(h)) if "smsp__average_warps_issue_stalled" in h[i]
          and h[i].endswith("_per_issue_active.ratio") and r[i] not in ("", "n/a")]
tot = sum(v for v, _ in stalls) or 1.0
print(names[idx])
print("stall reasons: " + " ".join("%s %.0f%%" % (k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), 100 * v / tot)
                                    for v, k in sorted(stalls, reverse=True)[:8]))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", str(idx),
                      "--launch-count", "1"], capture_output=True, text=True).stdout
cur_file = None
hdr = None
lines = {}
for r in csv.reader(out.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        hdr = None
        continue
    if r[0] == "Line No":
        hdr = r
        S, I = hdr.index("# Samples"), hdr.index("Instructions Executed")
        continue
    if hdr and r[0] != "" and len(r) == len(hdr):
        try:
            key = (cur_file, int(r[0]), r[1].strip())
            e = lines.setdefault(key, [0, 0])
            e[0] += int(r[S])
            e[1] += int(r[I])
        except ValueError:
            pass
tot_s = sum(v[0] for v in lines.values()) or 1
tot_i = sum(v[1] for v in lines.values()) or 1
files = {}
for (f, ln, src), (s_, i_) in lines.items():
    e = files.setdefault(f, [0, 0])
    e[0] += s_
    e[1] += i_
print("samples %d instructions %d" % (tot_s, tot_i))
for f, (s_, i_) in sorted(files.items(), key=lambda kv: -kv[1][1]):
    print("%5.1f%% smp %5.1f%% ins  %s" % (100.0 * s_ / tot_s, 100.0 * i_ / tot_i, f))
for (f, ln, src), (s_, i_) in sorted(lines.items(), key=lambda kv: -kv[1][1])[:n]:
    print("%5.1f%% smp %5.1f%% ins  %s:%d  %s" % (100.0 * s_ / tot_s, 100.0 * i_ / tot_i, f, ln, src[:100]))
