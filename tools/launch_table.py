"""Summarises an ncu gpu__time_duration launch list (csv) per kernel. Usage: launch_table.py CSV"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
h = rows[hdr]; ki = h.index('Kernel Name'); vi = h.index('Metric Value')
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) > vi:
        agg.setdefault(r[ki].split('(')[0], []).append(float(r[vi].replace(',', '')))
tot = sum(sum(v) for v in agg.values())
for k, v in agg.items():
    print(f"{k[:70]:70s} launches={len(v):4d} mean_us={sum(v)/len(v)/1e3:9.2f} min_us={min(v)/1e3:9.2f} share={100*sum(v)/tot:5.1f}%")
