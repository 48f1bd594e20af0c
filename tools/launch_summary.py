#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: launches, total and share.
Usage: python tools/launch_summary.py gpurun_out/launches.csv [out.csv]"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1.0)
        a = agg.setdefault(r[ki].split("(")[0], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    lines = ["kernel,launches,total_us,share_pct"]
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        lines.append("%s,%d,%.1f,%.1f" % (k.replace(",", ";"), a[0], a[1], 100 * a[1] / tot))
    lines.append("TOTAL,%d,%.1f,100.0" % (sum(a[0] for a in agg.values()), tot))
    text = "\n".join(lines) + "\n"
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text)
    sys.stdout.write(text)


if __name__ == "__main__":
    main()
