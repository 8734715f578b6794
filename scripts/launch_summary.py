"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.  usage: launch_summary.py file.csv [skip_rows]"""
import collections, csv, sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5 and r[0].isdigit()]
rows = rows[int(sys.argv[2]) if len(sys.argv) > 2 else 0:]
agg = collections.OrderedDict()
for r in rows:
    name, v, u = r[4][:110], float(r[-1].replace(",", "")), r[-2]
    us = v / 1000 if u.startswith("n") else (v * 1000 if u.startswith("m") else v)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(v[1] for v in agg.values())
print(f"{len(rows)} launches, {tot:.1f} us")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n:5d} {us / n:9.1f} us {100 * us / tot:5.1f}%  {k}")
