"""Text summary of an ncu report: one line per profiled kernel launch (for profiles/)."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units = rows[0], rows[1]
ix = {k: i for i, k in enumerate(h)}
cols = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs"), ("launch__occupancy_limit_shared_mem", "occ_smem"),
        ("launch__occupancy_limit_registers", "occ_regs"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_pipe%"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"), ("smsp__inst_executed.sum", "inst")]
for r in rows[2:]:
    name = r[ix["Kernel Name"]]
    parts = []
    for c, lab in cols:
        if c in ix:
            v = r[ix[c]]
            try:
                v = "%.4g" % float(v)
            except ValueError:
                pass
            parts.append("%s=%s%s" % (lab, v, (" " + units[ix[c]]) if units[ix[c]] not in ("", "%") and lab in ("time", "dram_rd", "dram_wr") else ""))
    print(name)
    print("    " + "  ".join(parts))
