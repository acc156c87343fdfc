#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv --log-file X.csv` launch list into per-kernel totals / shares (the file
committed under profiles/).  usage: python tools/ncu_launch_summary.py X.csv [--last N] [--note "..."] > profiles/rNN_launches_*.csv"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    last, note = None, ""
    a = sys.argv[2:]
    while a:
        if a[0] == "--last":
            last, a = int(a[1]), a[2:]
        elif a[0] == "--note":
            note, a = a[1], a[2:]
        else:
            a = a[1:]
    lines = [l for l in open(path, errors="replace") if not l.startswith("==")]
    rows = list(csv.reader(lines))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hi]
    ki, vi, ui, mi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit"), h.index("Metric Name")
    launches = []
    for r in rows[hi + 1:]:
        if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", ""))
        u = r[ui]
        ms = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v * 1e3 if u in ("s", "second") else v
        name = re.sub(r"\(.*$", "", r[ki])
        name = re.sub(r"^void ", "", name).replace("ea::", "")
        launches.append((name, ms))
    if last:
        launches = launches[-last:]
    agg = collections.OrderedDict()
    for n, ms in launches:
        e = agg.setdefault(n, [0, 0.0])
        e[0] += 1
        e[1] += ms
    tot = sum(v[1] for v in agg.values())
    if note:
        print("# " + note)
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (per-launch times are cold-cache and serialised: compare SHARES)")
    print(f"# total {tot:.1f} ms over {len(launches)} launches")
    print("ms,share_pct,launches,avg_ms,kernel")
    for n, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{ms:.3f},{100 * ms / tot:.2f},{c},{ms / c:.4f},{n}")


if __name__ == "__main__":
    main()
