#!/usr/bin/env python
"""Print an ncu launch list (--csv --metrics ... log file): one line per launch."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
i_name, i_m, i_v, i_id = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('ID')
d = collections.OrderedDict()
for r in rows[1:]:
    d.setdefault((r[i_id], r[i_name]), {})[r[i_m]] = float(r[i_v].replace(',', ''))
short = {'dram__bytes_read.sum': 'rd_GB', 'dram__bytes_write.sum': 'wr_GB', 'gpu__time_duration.sum': 'ms',
         'smsp__inst_executed.sum': 'Minst', 'sm__inst_executed.avg.per_cycle_elapsed': 'ipc',
         'smsp__thread_inst_executed_per_inst_executed.ratio': 'thr/inst'}
scale = {'rd_GB': 1e-9, 'wr_GB': 1e-9, 'ms': 1e-6, 'Minst': 1e-6}
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
for (id_, name), m in list(d.items())[:n]:
    nm = name.split('(')[0].split('::')[-1][:36]
    print(f"{id_:>4} {nm:<36}", ' '.join(f"{short.get(k, k)}={v * scale.get(short.get(k, k), 1):.4g}" for k, v in m.items()))
