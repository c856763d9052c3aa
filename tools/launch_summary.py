"""Aggregate an ncu launch list (--csv, gpu__time_duration.sum) per kernel: python tools/launch_summary.py list.csv"""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = None
agg = collections.OrderedDict()
for r in rows:
    if 'Kernel Name' in r:
        hdr = {h: i for i, h in enumerate(r)}
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    name = r[hdr['Kernel Name']]
    name = re.sub(r'dai::|\(anonymous namespace\)::|<unnamed>::|unnamed>::', '', name)
    name = name.split('(')[0]
    try:
        v = float(r[hdr['Metric Value']].replace(',', ''))
    except Exception:
        continue
    unit = r[hdr['Metric Unit']]
    us = v / 1000.0 if unit in ('ns', 'nsecond') else (v if unit in ('us', 'usecond') else v * 1000.0)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += us
tot = sum(a[1] for a in agg.values())
print("kernel n total_us avg_us share")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-70s %5d %10.1f %8.1f %5.1f%%" % (k[:70], n, t, t / n, 100 * t / tot))
print("total us %.1f" % tot)
