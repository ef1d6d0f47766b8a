"""Summarise an ncu source page (SASS) for one kernel: top instructions by stall samples and totals per opcode."""
import csv, subprocess, sys, collections
rep, kid = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", ":::" + kid], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
print(rows[0][:2])
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
S = idx["# Samples"]; I = idx["Instructions Executed"]
tot_s = sum(int(r[S]) for r in data); tot_i = sum(int(r[I]) for r in data)
print("instructions", tot_i, "samples", tot_s, "sass lines", len(data))
ops = collections.Counter(); opi = collections.Counter()
for r in data:
    op = r[idx["Source"]].split()[0] if r[idx["Source"]].split()[0][0] != '@' else r[idx["Source"]].split()[1]
    op = op.split('.')[0]
    ops[op] += int(r[S]); opi[op] += int(r[I])
print("by opcode (samples%, instr%):")
for op, s in ops.most_common(18):
    print("  %-10s %5.1f%% %5.1f%%" % (op, 100.0 * s / tot_s, 100.0 * opi[op] / tot_i))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {h: sum(int(r[idx[h]]) for r in data) for h in stalls}
print("stall reasons:", ", ".join("%s %.1f%%" % (h[6:], 100.0 * v / tot_s) for h, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]))
n = int(sys.argv[3]) if len(sys.argv) > 3 else 0
if n:
    print("top instructions by samples:")
    order = sorted(range(len(data)), key=lambda i: -int(data[i][S]))[:n]
    for i in order:
        r = data[i]
        print("  #%5d %6s smp %9s inst  %s" % (i, r[S], r[I], r[idx["Source"]].strip()[:90]))
