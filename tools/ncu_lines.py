#!/usr/bin/env python
"""Per-source-line hot spots from an ncu report:
   ncu -i rep.ncu-rep --page source --print-source cuda,sass --csv > src.csv ; python tools/ncu_lines.py src.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
fname, hdr, data = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
        hdr["SASS"] = 3
    elif hdr and len(r) > 8 and r[2] == "-":
        data.append((fname, r))
ie, sm, lst = hdr["Instructions Executed"], hdr["# Samples"], hdr["L2 Theoretical Sectors Local"]
f = lambda x: float(x) if x.replace(".", "").isdigit() else 0.0
tot_i = sum(f(r[ie]) for _, r in data)
tot_s = sum(f(r[sm]) for _, r in data)
print("total warp instructions %.3g, samples %d" % (tot_i, tot_s))
print("%7s %7s  %-22s %s" % ("inst%", "smpl%", "file:line", "source"))
for fn, r in sorted(data, key=lambda x: -f(x[1][ie]))[:top_n]:
    print("%6.2f%% %6.2f%%  %-22s %s" % (100 * f(r[ie]) / tot_i, 100 * f(r[sm]) / max(tot_s, 1), "%s:%s" % (fn[:14], r[0]), r[1].strip()[:110]))
