"""Per-kernel numbers of an `ncu --set full` report as JSON (profiles/rNN_kernels.json): launch duration, DRAM bytes
read + written, FP64 pipe and issue-slot utilisation.  bench.py reads `dram_bytes` (roofline.traffic) and
`fp64_pipe_pct` of the dominant kernel from it.  usage: ncu_kernels_json.py report.ncu-rep "how it was captured" """
import csv, json, re, subprocess, sys
rep, how = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units = rows[0], rows[1]
ix = {k: i for i, k in enumerate(h)}
SHORT = {"wb_hv_channels_fft": "hv_channels_fft", "wb_hv_channels": "hv_channels_direct", "wb_hv_fft_fwd": "hv_fft_fwd",
         "wb_hv_refine_items, 128": "hv_refine", "wb_hv_contour": "hv_contour", "wb_hv_prune": "hv_prune",
         "wb_hv_detect": "hv_detect", "wb_cheaptrick_body": "cheaptrick", "wb_d4c_body": "d4c"}


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


res = {"_source": how}
for r in rows[2:]:
    name = r[ix["Kernel Name"]]
    key = None
    for pat in sorted(SHORT, key=len, reverse=True):
        if pat in name:
            key = SHORT[pat]
            break
    if key is None or key in res:
        continue
    rd, wr = num(r, "dram__bytes_read.sum"), num(r, "dram__bytes_write.sum")
    res[key] = {"kernel": name, "time_ms": num(r, "gpu__time_duration.sum"),
                "dram_bytes": (rd or 0) + (wr or 0),
                "fp64_pipe_pct": num(r, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", False),
                "issue_pct": num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active", False),
                "registers": num(r, "launch__registers_per_thread", False),
                "warps_active_pct": num(r, "sm__warps_active.avg.pct_of_peak_sustained_active", False)}
print(json.dumps(res, indent=1))
