"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (per learner step)."""
import collections
import csv
import sys

path = sys.argv[1]
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] in ("ns", "nsecond") else v * 1e3 if r[ui] in ("ms", "msecond") else v
    name = r[ki].split('(')[0].replace('void ', '')
    agg[name][0] += 1
    agg[name][1] += v
nsteps = max(agg["adam_kernel"][0], agg["riqn::adam_kernel"][0], 1)
tot = sum(v[1] for v in agg.values())
print(f"# {len(rows)} launches, {nsteps} learner steps (adam_kernel count); cold-cache serialised times: compare SHARES")
print(f"# {'kernel':30s} {'launches/step':>13s} {'us/step':>10s} {'share':>7s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:32s} {v[0] / nsteps:13.1f} {v[1] / nsteps:10.1f} {v[1] / tot * 100:6.1f}%")
print(f"# total {tot / nsteps:.1f} us/step")
