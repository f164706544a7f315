#!/usr/bin/env python
"""Top source lines of one kernel in an .ncu-rep (needs -lineinfo + --import-source on):
per line, warp instructions executed, average active threads, stall samples."""
import csv, io, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
files = {}
cur = None
agg = collections.defaultdict(lambda: [0, 0, 0])   # (file,line) -> inst, thread_inst, samples
src = {}
hdr = None
fname = None
for row in csv.reader(io.StringIO(out)):
    if not row:
        continue
    if row[0] == "File Name":
        fname = row[1].split("/")[-1]
        continue
    if row[0] == "Line No":
        hdr = row
        continue
    if hdr is None or len(row) < len(hdr):
        continue
    d = dict(zip(hdr, row))
    try:
        ln = int(row[0])
    except ValueError:
        continue
    src[(fname, ln)] = row[1]
    def num(k):
        try:
            return float(d.get(k, "0").replace(",", "") or 0)
        except ValueError:
            return 0.0
    a = agg[(fname, ln)]
    a[0] += num("Instructions Executed")
    a[1] += num("Thread Instructions Executed")
    a[2] += num("# Samples")
tot_i = sum(a[0] for a in agg.values()) or 1
tot_s = sum(a[2] for a in agg.values()) or 1
print(f"total warp inst {tot_i/1e9:.3f} G, samples {tot_s:.0f}")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    if a[0] == 0:
        continue
    print(f"{f}:{ln:<5d} inst {100*a[0]/tot_i:5.1f}%  thr/inst {a[1]/max(a[0],1):5.1f}  samples {100*a[2]/tot_s:5.1f}%  | {src[(f, ln)].strip()[:90]}")
