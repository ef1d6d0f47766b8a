"""Per-stage numbers of an `ncu --set full` report as JSON (profiles/rNN_kernels.json): launch duration, DRAM bytes
read + written, FP64 operations (dadd + dmul + 2 dfma thread instructions), FP64 pipe and issue-slot utilisation,
shared-memory bank conflicts.  bench.py reads `fp64_flop_per_frame` (the flop model of its FP64 roofline) and
`dram_bytes` (roofline.traffic) of the slowest stage from it.

usage: ncu_kernels_json.py report.ncu-rep <config key> <frames of the launch> "how it was captured" [existing.json]
(merges into existing.json when given; prints the merged JSON)"""
import csv
import json
import subprocess
import sys

rep, cfg_key, frames, how = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
merged = json.load(open(sys.argv[5])) if len(sys.argv) > 5 else {}
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units = rows[0], rows[1]
ix = {k: i for i, k in enumerate(h)}
# kernel-name fragment -> stage key of Engine.profile_stages / profile_decode (longest fragment wins)
STAGE = {"wb_hv_dec_": "hv_decimate", "wb_hv_fft_fwd": "hv_channels", "wb_hv_channels": "hv_channels",
         "wb_hv_detect": "hv_detect", "wb_hv_refine": "hv_refine", "wb_hv_prune": "hv_prune",
         "wb_hv_contour": "hv_contour", "wb_cheaptrick_body": "cheaptrick", "wb_d4c_body": "d4c",
         "wb_sy_timebase": "sy_timebase", "wb_sy_prefix": "sy_timebase", "wb_sy_pulses": "sy_synthesis",
         "wb_sy_normalise": "sy_synthesis", "wb_rq_": "rq_synthesis"}


def num(r, key, scale_units=True):
    if key not in ix:
        return None
    try:
        v = float(r[ix[key]])
    except ValueError:
        return None
    u = units[ix[key]]
    if scale_units:
        v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3}.get(u, 1.0)
    return v


res = {}
requiem = any("wb_rq_" in r[ix["Kernel Name"]] for r in rows[2:]) or cfg_key.endswith("requiem") or cfg_key == "config3"
for r in rows[2:]:
    name = r[ix["Kernel Name"]]
    key = None
    for pat in sorted(STAGE, key=len, reverse=True):
        if pat in name:
            key = STAGE[pat]
            break
    if key is None:
        continue
    if key == "d4c" and requiem:
        key = "d4c_requiem"
    if key == "sy_synthesis" and requiem:
        key = "rq_synthesis"
    cyc = num(r, "sm__cycles_elapsed.max", False) or 0.0
    per = lambda op: (num(r, "smsp__sass_thread_inst_executed_op_%s_pred_on.sum.per_cycle_elapsed" % op, False) or 0.0)
    flops = (per("dadd") + per("dmul") + 2.0 * per("dfma")) * cyc
    rd, wr = num(r, "dram__bytes_read.sum"), num(r, "dram__bytes_write.sum")
    e = res.setdefault(key, {"kernels": [], "time_ms": 0.0, "dram_bytes": 0.0, "fp64_flop": 0.0})
    e["kernels"].append(name)
    t = num(r, "gpu__time_duration.sum") or 0.0
    e["time_ms"] += t
    e["dram_bytes"] += (rd or 0) + (wr or 0)
    e["fp64_flop"] += flops
    if t >= max(e.get("_top_ms", 0.0), 1e-12):  # utilisation figures of the stage's longest kernel
        e["_top_ms"] = t
        e["fp64_pipe_pct"] = num(r, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", False)
        e["issue_pct"] = num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active", False)
        e["registers"] = num(r, "launch__registers_per_thread", False)
        e["warps_active_pct"] = num(r, "sm__warps_active.avg.pct_of_peak_sustained_active", False)
        conf = num(r, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", False)
        wav = num(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", False)
        e["smem_conflict_frac"] = (conf / wav) if conf is not None and wav else None
for e in res.values():
    e.pop("_top_ms", None)
    e["fp64_flop_per_frame"] = e["fp64_flop"] / frames
    e["kernels"] = sorted(set(e["kernels"]))
res["_source"] = how
res["_frames"] = frames
merged[cfg_key] = res
print(json.dumps(merged, indent=1))
