"""Per-kernel totals of one step from an ncu launch list (gpu__time_duration.sum CSV).
usage: python tools/launch_summary.py launches.csv [launches_per_step]
The list holds several identical steps (warm-up + timed + e2e); the LAST complete device step is summarised."""
import csv
import re
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
names = [re.sub(r"\(.*", "", r[4]).replace("void ", "") for r in rows]
ns = [float(r[-1]) for r in rows]
# a step starts at the domain's fill_i32_kernel with the grid-sized launch
starts = [i for i, r in enumerate(rows) if "fill_i32_kernel" in r[4] and r[8] == rows[0][8]]
print(f"{len(rows)} launches, {len(starts)} steps in the list")
per_step = [(starts[i], starts[i + 1] if i + 1 < len(starts) else len(rows)) for i in range(len(starts))]
for k, (a, b) in enumerate(per_step):
    print(f"  step {k}: launches {b - a:4d}  sum of kernel durations {sum(ns[a:b]) / 1e6:8.3f} ms")
# default: the step after the three warm-up steps = the timed device-resident step of `bench.py --steps 1 --warmup 3`
# (the steps after it are the host-API / e2e steps, which run the transport-encoding variant of K3)
sel = int(sys.argv[2]) if len(sys.argv) > 2 else min(3, len(per_step) - 1)
a, b = per_step[sel]
tot = OrderedDict()
for n, t in zip(names[a:b], ns[a:b]):
    c = tot.setdefault(n, [0, 0.0])
    c[0] += 1
    c[1] += t
total = sum(v[1] for v in tot.values())
print(f"step {sel}: {b - a} launches, {total / 1e6:.3f} ms of kernel time (cold-cache, serialised under ncu)")
for n, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"  {t / 1e6:9.3f} ms  {100 * t / total:5.1f}%  x{c:<3d} {n[:110]}")
