#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.

    python tools/launch_summary.py profiles/x_launches.csv "<command line that produced it>" > profiles/x_launches_summary.txt
"""
import collections
import csv
import sys


def main(path: str, header: str) -> None:
    rows = list(csv.reader(open(path)))
    hdr, data = None, []
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            data.append(r)
    ci = {h: i for i, h in enumerate(hdr)}
    agg: dict[str, list] = collections.OrderedDict()
    for r in data:
        if r[ci["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = r[ci["Kernel Name"]].split("(")[0][:88]
        v = float(r[ci["Metric Value"]].replace(",", ""))
        unit = r[ci["Metric Unit"]]
        v = v / 1000 if unit in ("ns", "nsecond") else v * 1000 if unit in ("ms", "msecond") else v
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    n = sum(a[0] for a in agg.values())
    print(f"# command: {header}")
    print("# ncu launch list summary (gpu__time_duration.sum, us; cold-cache serialised launches; "
          f"compare SHARES): total {tot:.1f} us over {n} launches")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name:90s} n={a[0]:4d} total={a[1]:9.1f} us avg={a[1] / a[0]:8.1f} us share={100 * a[1] / tot:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
