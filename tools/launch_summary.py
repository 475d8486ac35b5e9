#!/usr/bin/env python
"""Per-launch table of an ncu `--metrics gpu__time_duration.sum --csv` log: python tools/launch_summary.py file.csv [skip_first_n]"""
import csv, io, sys
txt = open(sys.argv[1]).read()
rows = list(csv.DictReader(io.StringIO(txt[txt.index('"ID"'):])))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else len(rows) // 2
tot = 0.0
agg = {}
for r in rows[skip:]:
    t = float(r['Metric Value'].replace(',', ''))
    t = t / 1000 if r['Metric Unit'] == 'ns' else t * 1000 if r['Metric Unit'] == 'ms' else t
    tot += t
    name = r['Kernel Name'].split('(')[0][-48:]
    print(f"  {name:48s} grid {r['Grid Size']:>14s} blk {r['Block Size']:>12s} {t:9.1f} us")
    agg[name] = agg.get(name, 0) + t
print("  total us", round(tot, 1))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
    print(f"    {k:48s} {v:9.1f} us  {100 * v / tot:5.1f} %")
