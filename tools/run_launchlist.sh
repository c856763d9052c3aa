DAI_GRAPHS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 700 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --quick --no-extras --no-cpu-baseline > gpurun_out/r02_ncu_list.log 2>&1
python - <<'PY'
import csv, collections, re
rows = list(csv.reader(open('gpurun_out/r02_launches.csv')))
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
PY
