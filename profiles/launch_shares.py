#!/usr/bin/env python
"""Kernel shares of a step from an ncu launch list (--metrics gpu__time_duration.sum --csv):
    python profiles/launch_shares.py profiles/r1_launches_spline.csv [...]"""
import collections
import csv
import sys


def shares(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = collections.OrderedDict()
    for r in rows:
        if r is hdr or r[ik] == "Kernel Name":
            continue
        try:
            v = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}.get(r[iu], 1e-6)
        n, t = tot.get(r[ik], (0, 0.0))
        tot[r[ik]] = (n + 1, t + v)
    total = sum(t for _, t in tot.values())
    print(f"== {path}: {sum(n for n, _ in tot.values())} launches (cold-cache serialised times: compare SHARES)")
    for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:70]:70s} n={n:3d} total={t:9.3f} ms share={t / total:.3f}")


if __name__ == "__main__":
    for p in sys.argv[1:]:
        shares(p)
