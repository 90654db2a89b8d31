"""Shares of the device-resident region in an ncu launch list of bench.py
(`ncu --metrics gpu__time_duration.sum --csv`): the launches with grids of 888 and 544 problems
(4096 = 4 x 888 + 544 per step; the 64-problem launches are the e2e region).

    python tools/launch_shares.py profiles/launches_r01.csv
"""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[start]
    ki, vi, gi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size")
    tot, full = collections.Counter(), collections.defaultdict(list)
    for r in rows[start + 1:]:
        if len(r) <= vi or not (r[gi].startswith("(888") or r[gi].startswith("(544")):
            continue
        for k in ("fwd", "energy", "bwd", "finalize"):
            if k in r[ki]:
                ms = float(r[vi].replace(",", "")) / 1e6
                tot[k] += ms
                if r[gi].startswith("(888"):
                    full[k].append(ms)
    s = sum(tot.values())
    for k in ("fwd", "energy", "bwd", "finalize"):
        print(f"{k:9s} share {100 * tot[k] / s:5.1f} %   888-problem launch {sum(full[k]) / max(len(full[k]), 1):7.3f} ms "
              f"({len(full[k])} launches)")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "profiles/launches_r01.csv")
