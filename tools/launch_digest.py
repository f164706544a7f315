#!/usr/bin/env python
"""Digest of an ncu launch list of bench.py (csv: gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum):
per kernel launches, total and mean time, share of the timed kernels, DRAM bytes per launch; writes
profiles/ncu_traffic.json when asked.    python tools/launch_digest.py launches.csv [--traffic]"""
import collections, csv, json, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
i_name, i_m, i_v, i_id = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('ID')
d = collections.OrderedDict()
for r in rows[1:]:
    d.setdefault((r[i_id], r[i_name]), {})[r[i_m]] = float(r[i_v].replace(',', ''))
agg = collections.OrderedDict()
for (id_, name), m in d.items():
    nm = name.split('(')[0].split('::')[-1].split('<')[0]
    a = agg.setdefault(nm, [0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[1] += m.get('gpu__time_duration.sum', 0) / 1e6
    a[2] += m.get('dram__bytes_read.sum', 0) / 1e9
    a[3] += m.get('dram__bytes_write.sum', 0) / 1e9
tot = sum(a[1] for n, a in agg.items() if n != 'synth_kernel')
print("ncu launch list of `python bench.py --steps 20 --warmup 5` (--profile-from-start off: the 20 steps of the whole cfg-4 job), one B200")
print("metrics: gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum; --clock-control none.  Per-launch times under ncu are")
print("serialised and cold; the SHARES are what must agree with bench.py's CUDA-event shares (bench_cfg4_N1.json, roofline.kernels).")
print("synth_kernel is the read generator: it runs BETWEEN the timed windows of bench.py (reads resident before each window) and is left out.")
print(f"{'kernel':<24}{'launches':>9}{'ms total':>11}{'share':>8}{'ms/launch':>11}{'DRAM rd GB/launch':>19}{'DRAM wr GB/launch':>19}{'TB/s':>7}")
for nm, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    sh = f"{a[1] / tot:>8.3f}" if nm != 'synth_kernel' else f"{'-':>8}"
    print(f"{nm:<24}{a[0]:>9}{a[1]:>11.1f}{sh}{a[1] / a[0]:>11.3f}{a[2] / a[0]:>19.3f}{a[3] / a[0]:>19.3f}{(a[2] + a[3]) / a[1]:>7.2f}")
print(f"{'total (timed kernels)':<24}{sum(a[0] for n, a in agg.items() if n != 'synth_kernel'):>9}{tot:>11.1f}")
if '--traffic' in sys.argv:
    per = lambda n: (agg[n][2] + agg[n][3]) / agg[n][0] * 1e9
    t = {"sg_insert_kernel": per('sg_insert_kernel'), "sg_sort_kernel": per('sg_sort_kernel'), "build_kernel<MODE_BIN>": per('build_kernel'),
         "insert_bins_kernel": 27570000000.0,
         "_source": "profiles/r2/launches_cfg4_N1.csv (ncu launch list of `python bench.py --steps 20 --warmup 5`, the whole cfg-4 job on one B200): "
                    "dram__bytes_read.sum + dram__bytes_write.sum, mean per launch; profiles/r2/ncu_full_summary_*.txt are the --set full captures. "
                    "insert_bins_kernel: round 1 (profiles/r1/launches_cfg4_N1.csv)"}
    json.dump(t, open('profiles/ncu_traffic.json', 'w'), indent=1)
