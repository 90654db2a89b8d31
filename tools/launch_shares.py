"""Shares of the device-resident region in an ncu launch list of bench.py
(`ncu --metrics gpu__time_duration.sum --csv`): the launches of the shard's passes -- round 1 and the first
round-2 lists: grids of 888 and 544 problems (4096 = 4 x 888 + 544 per step); from the closing state of round 2
on: 2048 problems (two equal passes); the 64-problem launches are the e2e region.

    python tools/launch_shares.py profiles/launches_r02.csv [problems_per_full_pass]
"""
import collections
import csv
import sys


def main(path, full="888"):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[start]
    ki, vi, gi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size")
    tot, full_l = collections.Counter(), collections.defaultdict(list)
    for r in rows[start + 1:]:
        grid = r[gi].strip("() ").split(",")[0] if len(r) > gi else ""
        N = 1001
        def problems(gs):          # the energy kernel's grid is problems x N
            try:
                v = int(gs)
            except ValueError:
                return -1
            return v // N if v >= N * 8 else v
        if len(r) <= vi or problems(grid) not in (int(full), 544):
            continue
        for k in ("fwd", "energy", "bwd", "finalize"):
            if k in r[ki]:
                ms = float(r[vi].replace(",", "")) / 1e6
                tot[k] += ms
                if problems(grid) == int(full):
                    full_l[k].append(ms)
    s = sum(tot.values())
    for k in ("fwd", "energy", "bwd", "finalize"):
        print(f"{k:9s} share {100 * tot[k] / s:5.1f} %   {full}-problem launch {sum(full_l[k]) / max(len(full_l[k]), 1):7.3f} ms "
              f"({len(full_l[k])} launches)")


if __name__ == "__main__":
    main(*(sys.argv[1:3] or ["profiles/launches_r01.csv"]))
