#!/usr/bin/env python
"""Summarises an ncu report's SASS source page: instruction mix, top stall sites, shared-memory conflict sites.
Usage: python tools/ncu_hot.py report.ncu-rep [top_n] [kernel_index]   (kernel_index: section number, or a kernel-name substring)"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
starts = [i for i, l in enumerate(lines) if l.startswith('"Address"')]
karg = sys.argv[3] if len(sys.argv) > 3 else "0"
if karg.lstrip("-").isdigit():
    kidx = int(karg)
else:  # kernel-name substring: first section whose header names it
    kidx = next(i for i, st in enumerate(starts) if karg in lines[st - 1])
start = starts[kidx]
end = starts[kidx + 1] - 1 if kidx + 1 < len(starts) else len(lines)
print(lines[start - 1][:160])
rows = list(csv.DictReader(lines[start:end]))
tot_inst = sum(int(r["Instructions Executed"]) for r in rows)
tot_samp = sum(int(r["# Samples"]) for r in rows)
mix = collections.Counter()
for r in rows:
    op = r["Source"].split()[0] if not r["Source"].strip().startswith("@") else r["Source"].split()[1]
    mix[op.split(".")[0]] += int(r["Instructions Executed"])
print(f"instructions executed (warp level): {tot_inst}, samples: {tot_samp}")
print("mix:", ", ".join(f"{k} {v / tot_inst:.1%}" for k, v in mix.most_common(22)))
stall_cols = [c for c in rows[0] if c.startswith("stall_") and "Not Issued" not in c]
agg = collections.Counter()
for r in rows:
    for c in stall_cols:
        agg[c] += int(r[c] or 0)
print("stalls:", ", ".join(f"{k[6:]} {v / max(tot_samp, 1):.1%}" for k, v in agg.most_common(10)))
print("-- top sample sites")
for i, r in sorted(enumerate(rows), key=lambda t: -int(t[1]["# Samples"]))[:top]:
    st = max(stall_cols, key=lambda c: int(r[c] or 0))
    print(f"{i:5d} {int(r['# Samples']) / max(tot_samp, 1):6.2%} {st[6:]:12s} {r['Source'].strip()[:90]}")
print("-- shared memory excessive wavefronts")
for i, r in sorted(enumerate(rows), key=lambda t: -int(t[1]["L1 Wavefronts Shared Excessive"] or 0))[:12]:
    if int(r["L1 Wavefronts Shared Excessive"] or 0) == 0:
        break
    print(f"{i:5d} excess {r['L1 Wavefronts Shared Excessive']:>10s} of {r['L1 Wavefronts Shared']:>10s} ideal {r['L1 Wavefronts Shared Ideal']:>10s} {r['Source'].strip()[:80]}")
