#!/usr/bin/env python
"""tools/ncu_lines.py REPORT.ncu-rep KERNEL_INDEX (ordinal among the distinct kernel functions of the report) [min_pct] -> executed instructions and stall samples per CUDA source line
(ncu --page source --print-source cuda,sass of a report captured with --import-source on; kernels built with -lineinfo)."""
import collections
import csv
import subprocess
import sys


def main():
    rep, k = sys.argv[1], sys.argv[2]
    minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = None
    fnames, active = [], False
    files, cur_file, cur = {}, None, None
    agg = collections.OrderedDict()
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1]
            continue
        if r[0] == "Function Name":
            if r[1] not in fnames:
                fnames.append(r[1])
            active = fnames.index(r[1]) == int(k)
            continue
        if not active:
            continue
        if r[0] == "Line No":
            hdr = r
            ie, ns = hdr.index("Instructions Executed"), hdr.index("# Samples")
            continue
        if hdr is None:
            continue
        if r[0].strip().isdigit():
            cur = (cur_file, int(r[0]))
            continue
        if len(r) == len(hdr) and r[2].startswith("0x"):
            try:
                a = agg.setdefault(cur, [0, 0])
                a[0] += int(r[ie])
                a[1] += int(r[ns])
            except ValueError:
                pass
    print("#", fnames[int(k)][:150] if int(k) < len(fnames) else "no such kernel", "of", len(fnames))
    tot = sum(a[0] for a in agg.values()) or 1
    ts = sum(a[1] for a in agg.values()) or 1
    print(f"# total warp instructions {tot}, stall samples {ts}")
    for (f, ln), (i, s) in sorted(agg.items(), key=lambda kv: (kv[0][0] or "", kv[0][1])):
        if i >= tot * minpct / 100 or s >= ts * minpct / 100:
            if f not in files:
                try:
                    files[f] = open(f).read().splitlines()
                except OSError:
                    files[f] = []
            src = files[f][ln - 1].strip()[:100] if 0 < ln <= len(files[f]) else ""
            print(f"{(f or '?').split('/')[-1]}:{ln:<4d} {i / tot * 100:5.1f}% inst {s / ts * 100:5.1f}% samples  {src}")


if __name__ == "__main__":
    main()
