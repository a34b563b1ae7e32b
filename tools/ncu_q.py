#!/usr/bin/env python
"""Query an ncu report: python tools/ncu_q.py file.ncu-rep KERNEL_SUBSTR regex [regex ...] (metric names matched with re.search)"""
import csv, io, re, subprocess, sys
rep, kern, pats = sys.argv[1], sys.argv[2], sys.argv[3:]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
for r in data:
    name = r[hdr.index("Kernel Name")]
    if kern not in name:
        continue
    print("###", name[:60])
    for i, h in enumerate(hdr):
        if any(re.search(p, h) for p in pats):
            print(f"  {h} = {r[i]} {units[i]}")
