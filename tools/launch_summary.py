"""Per-kernel summary of an ncu launch list (ncu --metrics gpu__time_duration.sum --csv).  usage: launch_summary.py <csv>"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
H = rows[hdr]
ki, vi = H.index('Kernel Name'), H.index('Metric Value')
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) > vi:
        agg.setdefault(r[ki][:70], []).append(float(r[vi].replace(',', '')))
tot = 0.0
for k, v in agg.items():
    print(f"{k:70s} n={len(v):3d} last={v[-1] / 1e3:9.1f} us  min={min(v) / 1e3:9.1f}")
    if 'generate' not in k:
        tot += v[-1] / 1e3
print(f"sum of the last launch of every kernel (without the generator): {tot:.1f} us")
