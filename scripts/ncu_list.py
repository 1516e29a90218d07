"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel (name + grid) count, mean, total us."""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
d = collections.OrderedDict()
for r in rows[skip:]:
    d.setdefault((r[4][:70], r[8]), []).append(float(r[-1]) / 1000)
tot = sum(sum(v) for v in d.values())
for (k, g), v in d.items():
    print(f"{len(v):4d} x {sum(v)/len(v):9.1f} us = {sum(v):9.1f} us ({100*sum(v)/tot:4.1f} %)  grid {g:16s} {k}")
print(f"total {tot:.1f} us")
