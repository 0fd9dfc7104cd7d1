"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel.
usage: python tools/launch_summary.py launches.csv [first_id last_id]   (ids select one step)"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[start]
    ki, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("ID")
    gi, bi = hdr.index("Grid Size"), hdr.index("Block Size")
    lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (0, 1 << 60)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[start + 1:]:
        if len(r) <= vi or not r[ii].isdigit() or not (lo <= int(r[ii]) <= hi):
            continue
        name = r[ki]
        for junk in ("void ", "mft::", "<unnamed>::", "(anonymous namespace)::"):
            name = name.replace(junk, "")
        name = name.split("(")[0][:70]
        agg[name][0] += 1
        agg[name][1] += float(r[vi].replace(",", ""))
    total = sum(v[1] for v in agg.values())
    print(f"{'us':>10} {'share':>6} {'n':>5}  kernel")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t / 1e3:10.1f} {100 * t / total:5.1f}% {c:5d}  {n}")
    print(f"{total / 1e3:10.1f} total, {sum(v[0] for v in agg.values())} launches")


if __name__ == "__main__":
    main()
