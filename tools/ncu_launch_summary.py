#!/usr/bin/env python
"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> per-kernel table (markdown).  Usage: ncu_launch_summary.py launches.csv [skip_first_n]"""
import csv
import sys
from collections import OrderedDict


def main(path, skip=0):
    rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[0].isdigit()]
    rows = rows[skip:]
    agg = OrderedDict()
    for r in rows:
        name = r[4].split("(")[0].replace("void ", "")
        name = name[name.rindex("::") + 2:] if "::" in name else name          # lcx::<unnamed>::k_name<...>
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[14])
    tot = sum(v[1] for v in agg.values())
    print("| kernel | launches | total ms | ms / launch | share |\n|---|---|---|---|---|")
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| %s | %d | %.3f | %.3f | %.3f |" % (k, n, ns / 1e6, ns / 1e6 / n, ns / tot))
    print("\ntotal %.2f ms over %d launches" % (tot / 1e6, len(rows)))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
