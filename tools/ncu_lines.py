"""Per-source-line hot spots of one kernel from an ncu report (needs -lineinfo + --import-source on).
usage: ncu_lines.py report.ncu-rep <kernel-id> [top-n]"""
import csv, subprocess, sys
rep, kid = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-id", ":::" + kid],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = None; hdr = None; lines = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": fn = r[1]; continue
    if r[0] == "Line No": hdr = r; S = hdr.index("# Samples"); I = hdr.index("Instructions Executed"); continue
    if hdr and r[0] != "" and len(r) == len(hdr):
        lines.append((int(r[S]), int(r[I]), cur_file, r[0], r[1].strip()))
tot_s = sum(l[0] for l in lines); tot_i = sum(l[1] for l in lines)
print(fn, "samples", tot_s, "instructions", tot_i)
for s_, i_, f, ln, src in sorted(lines, reverse=True)[:n]:
    print("%5.1f%% smp %5.1f%% ins  %s:%s  %s" % (100.0 * s_ / max(tot_s, 1), 100.0 * i_ / max(tot_i, 1), f, ln, src[:100]))
