"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel (name, grid, block) launches, total ms, share.
usage: python tools/ncu_launches.py gpurun_out/launches.csv [--by-shape]"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    by_shape = "--by-shape" in sys.argv
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    agg = OrderedDict()
    tot = 0.0
    for r in rows[1:]:
        name = re.sub(r"^void\s+", "", r[ki])
        name = re.sub(r"\(.*", "", name).replace("unnamed>::", "").replace("ttn::<", "")
        key = (name, r[gi], r[bi]) if by_shape else (name,)
        us = float(r[vi].replace(",", "")) / 1e3
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += us
        tot += us
    print(f"# launches {sum(a[0] for a in agg.values())}, summed device time {tot / 1e3:.3f} ms")
    print(f"# {'kernel':70s} {'launches':>8s} {'ms':>10s} {'share':>7s}")
    for key, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{' '.join(key)[:70]:70s} {n:8d} {us / 1e3:10.3f} {100 * us / tot:6.1f}%")


if __name__ == "__main__":
    main()
