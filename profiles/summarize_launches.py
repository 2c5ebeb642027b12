#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total
device time and share.  usage: summarize_launches.py launches.csv [first_id last_id]"""
import collections
import csv
import sys


def main():
    fn = sys.argv[1]
    lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    hi = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 60
    rows = list(csv.reader(l for l in open(fn) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, idi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("ID")
    agg = collections.defaultdict(lambda: [0, 0.0])
    order = []
    for r in rows[1:]:
        try:
            v, i = float(r[vi].replace(",", "")), int(r[idi])
        except ValueError:
            continue
        if not (lo <= i < hi):
            continue
        k = r[ki][:110]
        agg[k][0] += 1
        agg[k][1] += v
        order.append((i, k, v))
    tot = sum(v[1] for v in agg.values())
    print(f"# {fn}: ids [{lo},{hi}) {sum(v[0] for v in agg.values())} launches, {tot / 1e3:.1f} us total")
    print(f"# {'us':>9} {'n':>5} {'share':>6}  kernel")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1] / 1e3:10.1f} {v[0]:5d} {v[1] / tot * 100:5.1f}%  {k}")
    return order


if __name__ == "__main__":
    main()
