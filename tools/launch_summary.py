"""Digest of an `ncu --metrics gpu__time_duration.sum --clock-control none --csv` launch list: per-kernel totals and shares.
Usage: python tools/launch_summary.py launches.csv "<command line that was profiled>" """
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
for r in rows[1:]:
    try:
        v = float(r[ix["Metric Value"]].replace(",", ""))
    except ValueError:
        continue
    unit = r[ix["Metric Unit"]]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    agg.setdefault(r[ix["Kernel Name"]], []).append(us)
total = sum(sum(v) for v in agg.values())
ours = {k: v for k, v in agg.items() if "ursa::" in k}
print(sys.argv[2] if len(sys.argv) > 2 else "")
print("(cold-cache serialised launch times: compare SHARES)")
print("all launches: %d, total %.1f us; kernels of libursa_b200.so: %d launches, %.1f us (%.1f %%)"
      % (sum(len(v) for v in agg.values()), total, sum(len(v) for v in ours.values()), sum(sum(v) for v in ours.values()),
         100 * sum(sum(v) for v in ours.values()) / total))
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print("%12.1f us %6.2f%% n=%5d avg %9.1f us  %s" % (sum(v), 100 * sum(v) / total, len(v), sum(v) / len(v), k[:140]))
