"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per (kernel, grid)."""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
agg = collections.OrderedDict(); tot = 0
for row in csv.DictReader(lines):
    try: t = float(row['Metric Value'].replace(',', ''))
    except Exception: continue
    u = row['Metric Unit']
    t *= {'ns': 1, 'us': 1e3, 'ms': 1e6, 'nsecond': 1, 'usecond': 1e3, 'msecond': 1e6}.get(u, 1)
    key = (re.sub(r'\(.*', '', row['Kernel Name'])[:58], row.get('Grid Size', ''))
    a = agg.setdefault(key, [0, 0.0]); a[0] += 1; a[1] += t; tot += t
print(f"total {tot/1e6:.3f} ms over {sum(v[0] for v in agg.values())} launches")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[0]:58s} {k[1]:14s} n={v[0]:4d} total={v[1]/1e6:8.3f} ms avg={v[1]/v[0]/1e3:8.1f} us {100*v[1]/tot:5.1f}%")
