#!/usr/bin/env python
"""Top SASS instructions of one kernel in an .ncu-rep by warp-instructions executed."""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass",
                      "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
recs = []
for row in rows:
    if row and row[0] == "Address":
        hdr = row
        continue
    if hdr and len(row) >= len(hdr):
        d = dict(zip(hdr, row))
        def num(k):
            try:
                return float(d.get(k, "0").replace(",", "") or 0)
            except ValueError:
                return 0.0
        recs.append((num("Instructions Executed"), num("Thread Instructions Executed"), num("# Samples"), d["Address"], d["Source"]))
tot = sum(r[0] for r in recs) or 1
ts = sum(r[2] for r in recs) or 1
print(f"{len(recs)} SASS instructions, total executed {tot/1e9:.3f} G")
for i, r in enumerate(sorted(recs, key=lambda r: -r[0])[:top]):
    print(f"{100*r[0]/tot:5.1f}%  thr {r[1]/max(r[0],1):5.1f}  smp {100*r[2]/ts:5.1f}%  {r[3][-5:]}  {r[4][:100]}")
