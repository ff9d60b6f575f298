"""ncu launch-list CSV -> per-kernel totals (markdown).  usage: launch_table.py <csv> [skip_first_n_launches]"""
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if r and r[0].isdigit()]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = rows[skip:]
agg = {}
for r in rows:
  name, val, unit = r[4], float(r[-1]), r[-2]
  us = val / 1000.0 if unit in ("ns", "nsecond") else val
  a = agg.setdefault(name.split("(")[0].replace("void ", "")[:100], [0, 0.0])
  a[0] += 1
  a[1] += us
total = sum(v[1] for v in agg.values())
print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
  print("| `%s` | %d | %.1f | %.1f%% |" % (k, n, us, 100 * us / total))
print("\ntotal %.1f us over %d launches" % (total, len(rows)))
