"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: python profiles/launch_summary.py <csv>"""
import csv, sys
from collections import defaultdict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    v = float(r[vi].replace(',', '')) * {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0}.get(r[ui], 1e-6)
    agg[r[ki][:90]][0] += 1
    agg[r[ki][:90]][1] += v
tot = sum(v[1] for v in agg.values())
print(f'{len(rows) - 1} launches, {tot:.3f} ms of kernel time (cold-cache, serialised: shares matter, not absolutes)')
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{v[1]:10.3f} ms {v[0]:5d}x {100 * v[1] / tot:5.1f}%  {k}')
