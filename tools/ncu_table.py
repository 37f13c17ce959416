#!/usr/bin/env python
"""One line per kernel launch of an ncu report with the metrics the profiles/ tables quote, then the warp stall ratios.
Usage: python tools/ncu_table.py report.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]
cols = [("ms", "gpu__time_duration.sum"), ("grid", "launch__grid_size"), ("regs", "launch__registers_per_thread"), ("warp_inst", "smsp__inst_executed.sum"),
        ("issue%", "smsp__issue_active.avg.pct_of_peak_sustained_active"), ("alu%", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
        ("fmaheavy%", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"), ("warps/smsp", "smsp__warps_active.avg.per_cycle_active"),
        ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), ("l1_data_pipe%", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
        ("dram_rd_GB", "dram__bytes_read.sum"), ("dram_wr_GB", "dram__bytes_write.sum")]
kn = h.index("Kernel Name")
print("kernel".ljust(22), " ".join(c[0].rjust(12) for c in cols))
for r in rows[2:]:
    name = r[kn].split("(")[0].split("::")[-1][:22]
    vals = []
    for _, m in cols:
        v = r[h.index(m)] if m in h else ""
        try: v = f"{float(v):.4g}"
        except ValueError: pass
        vals.append(v.rjust(12))
    print(name.ljust(22), " ".join(vals))
st = [k for k in h if "issue_stalled" in k and "per_issue_active" in k and "not_issued" not in k]
print("\nwarp stall reasons per issued instruction (> 0.05)")
for r in rows[2:]:
    name = r[kn].split("(")[0].split("::")[-1][:22]
    d = {k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): float(r[h.index(k)] or 0) for k in st}
    print(name.ljust(22), " ".join(f"{k}={v:.2f}" for k, v in sorted(d.items(), key=lambda kv: -kv[1]) if v > 0.05))
