#!/usr/bin/env python
"""tools/launch_summary.py LAUNCHES.csv [title] -> markdown table: launches, total / share / average duration per kernel
(input: ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file LAUNCHES.csv ...)."""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else path
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("==")) if r]
    hdr = rows[0]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= mv:
            continue
        name = re.sub(r"\(.*$", "", r[kn]).replace("void <unnamed>::", "").replace("<unnamed>::", "")
        t = float(r[mv].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}.get(r[mu], 1e-6)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
    tot = sum(a[1] for a in agg.values()) or 1.0
    print(f"# {title}\n")
    print(f"source: `{path}` (`ncu --metrics gpu__time_duration.sum --clock-control none --csv`; serialised, cold cache: shares matter, not absolutes)\n")
    print("| kernel | launches | total ms | share | avg ms |\n|---|---|---|---|---|")
    for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name}` | {c} | {t:.2f} | {t / tot * 100:.1f}% | {t / c:.3f} |")


if __name__ == "__main__":
    main()
