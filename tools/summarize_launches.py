#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python tools/summarize_launches.py gpurun_out/launches.csv > profiles/xxx_launches.txt"""
import csv
import re
import sys
from collections import OrderedDict


def main(path):
    rows = []
    with open(path, newline="") as fh:
        lines = [l for l in fh if not l.startswith("==")]
    rd = csv.reader(lines)
    hdr = None
    for r in rd:
        if hdr is None:
            if "Kernel Name" in r:
                hdr = r
            continue
        rows.append(r)
    ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    total = 0.0
    for r in rows:
        if r[im] != "gpu__time_duration.sum":
            continue
        v = float(r[iv].replace(",", ""))
        u = r[iu]
        us = v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3)
        name = re.sub(r"^void ", "", r[ik])
        name = re.sub(r"vus::rt::k_(elem|coop)<vus::", r"k_\1<", name)
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        a = agg.setdefault(name, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += us
        a[2] = max(a[2], us)
        total += us
    print(f"# {path}: {sum(a[0] for a in agg.values())} launches, {total / 1e3:.3f} ms summed device time (cold-cache, serialised)")
    print(f"{'share%':>7} {'total_us':>12} {'launches':>9} {'avg_us':>10} {'max_us':>10}  kernel")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{100 * a[1] / total:7.2f} {a[1]:12.1f} {a[0]:9d} {a[1] / a[0]:10.2f} {a[2]:10.1f}  {name[:110]}")


if __name__ == "__main__":
    main(sys.argv[1])
