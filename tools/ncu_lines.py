#!/usr/bin/env python
"""Source-line hot spots of one kernel in an ncu report: samples and executed instructions per CUDA source line.
Usage: ncu_lines.py file.ncu-rep kernel_regex [top_n]"""
import csv
import subprocess
import sys


def main(path, kernel, top=25):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel, "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
    hdr = rows[hdr_i]
    cs, ci, cl, csrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), 0, 1
    items, seen, cur = [], set(), ""
    for r in rows[:hdr_i] + rows[hdr_i + 1:]:
        if not r:
            continue
        if r[0] in ("File Path", "File Name"):
            cur = r[1].split("/")[-1]
            if cur in seen:     # next launch of the same kernel: first one only
                break
            seen.add(cur)
            continue
        if r[0] in ("Function Name", "Line No", ""):   # "" = SASS row belonging to the source line above
            continue
        try:
            items.append((int(r[cs] or 0), int(r[ci] or 0), cur + ":" + r[cl], r[csrc].strip()[:150]))
        except (ValueError, IndexError):
            pass
    ts, ti = sum(i[0] for i in items) or 1, sum(i[1] for i in items) or 1
    print("total samples %d, warp instructions %d" % (ts, ti))
    for s, n, line, src in sorted(items, reverse=True)[:top]:
        print("%5.1f%% smp %5.1f%% inst  %-22s %s" % (100.0 * s / ts, 100.0 * n / ti, line, src))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 25)
