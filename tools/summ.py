#!/usr/bin/env python
"""Summarise gpurun_out/<tag>: bench JSON lines and ncu launch lists."""
import csv, sys, json, collections, glob, os
d = sys.argv[1]
for f in sorted(glob.glob(os.path.join(d, "bench_*.log"))):
    for line in open(f):
        if line.startswith("{"):
            j = json.loads(line)
            r = j.get("roofline") or {}
            ra = j.get("random_access") or {}
            e = j.get("e2e") or {}
            print(os.path.basename(f), "value %.3f G/s" % (j["value"] / 1e9), "ms/step %.2f" % j["ms_per_step"],
                  "kernel_ms %.2f" % r.get("kernel_ms", 0), "frac %.4f" % r.get("frac", 0),
                  "ra_frac %.2f" % ra.get("frac_of_random_access_ceiling", 0),
                  "e2e %.3f" % ((e.get("value") or 0) / 1e9), "launches", j.get("gpu_launches"),
                  "fp", j.get("graph", {}).get("fingerprint"))
        elif "rror" in line or "Traceback" in line:
            print(os.path.basename(f), line.strip()[:200])
for f in sorted(glob.glob(os.path.join(d, "launches*.csv"))):
    rows = [r for r in csv.reader(open(f)) if len(r) > 10]
    if not rows:
        continue
    hdr = rows[0]
    agg = collections.defaultdict(lambda: collections.defaultdict(list))
    for r in rows[1:]:
        x = dict(zip(hdr, r))
        agg[(x["Kernel Name"].split("(")[0][-48:], x["Grid Size"])][x["Metric Name"]].append(float(x["Metric Value"].replace(",", "")))
    print("==", os.path.basename(f))
    for k, v in agg.items():
        def m(name, div=1.0):
            x = v.get(name)
            return sum(x) / len(x) / div if x else float("nan")
        print("  %-50s %-12s n=%d  %.3f ms  rd %.2f GB wr %.2f GB  inst %.3f G  thr/inst %.1f  ipc %.2f" % (
            k[0], k[1], len(v["gpu__time_duration.sum"]), m("gpu__time_duration.sum", 1e6), m("dram__bytes_read.sum", 1e9),
            m("dram__bytes_write.sum", 1e9), m("smsp__inst_executed.sum", 1e9),
            m("smsp__thread_inst_executed_per_inst_executed.ratio"), m("sm__inst_executed.avg.per_cycle_elapsed")))
