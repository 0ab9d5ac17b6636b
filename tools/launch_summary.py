"""Summarises an ncu launch list (`--metrics gpu__time_duration.sum --csv`): per-kernel totals and
shares.  usage: python tools/launch_summary.py launches.csv [n_steps_in_capture]"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
h = rows[0]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
tot = collections.OrderedDict()
for r in rows[1:]:
    v, u = float(r[vi].replace(",", "")), r[ui]
    ms = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v if u in ("ms", "msecond") else v * 1e3
    n = r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "")[:60]
    tot.setdefault(n, [0, 0.0])
    tot[n][0] += 1
    tot[n][1] += ms
total = sum(v[1] for v in tot.values())
print(f"{len(rows) - 1} launches, {total:.3f} ms of kernel time")
print(f"{'ms':>10} {'share':>7} {'n':>5}  kernel")
for n, (c, ms) in sorted(tot.items(), key=lambda x: -x[1][1]):
    print(f"{ms:10.3f} {100 * ms / total:6.1f}% {c:5d}  {n}")
