#!/usr/bin/env python
"""DRAM traffic per super-droplet and launch of every kernel in an `ncu --set full` report -> profiles/traffic.json (read by bench.py
for roofline.traffic).  Keys are spelled the way the engine's live profile names the kernels, e.g. "(k_cond_range<M, true>)".
Usage: ncu_traffic.py report.ncu-rep n_sd [source note]"""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def live_name(kernel):
    """demangled ncu name -> the spelling of the engine's profile table"""
    k = kernel.split("(")[0].replace("void ", "")
    k = k[k.rindex("::") + 2:] if "::" in k else k          # lcx::<unnamed>::k_name<...>
    m = re.match(r"k_cond_range<(\d+), *(\d+)>", k)
    if m:
        return "(k_cond_range<M, %s>)" % ("true" if m.group(2) == "1" else "false")
    m = re.match(r"(k_transport|k_cond_staged)<(\d+)>", k)
    if m:
        return "%s<%s>" % (m.group(1), "true" if m.group(2) == "1" else "false")
    m = re.match(r"k_vterm_beard77<(\d+)>", k)
    if m:
        return "(k_vterm_beard77<%s>)" % ("true" if m.group(1) == "1" else "false")
    return k.split("<")[0]


def main(path, n_sd, note=""):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    acc = {}
    for r in rows[2:]:
        b = float(r[ri]) * scale[units[ri]] + float(r[wi]) * scale[units[wi]]
        acc.setdefault(live_name(r[ki]), []).append(b / n_sd)
    table = {k: sum(v) / len(v) for k, v in acc.items()}
    dst = os.path.join(ROOT, "profiles", "traffic.json")
    doc = json.load(open(dst)) if os.path.exists(dst) else {"source": "", "dram_bytes_per_sd": {}}
    doc["dram_bytes_per_sd"].update(table)          # kernels absent from this capture keep their earlier figures
    this = note or ("%s (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch, %d SDs)" % (os.path.basename(path), n_sd))
    doc["source"] = (doc["source"] + "; " if doc["source"] and this not in doc["source"] else "") + (this if this not in doc["source"] else "")
    json.dump(doc, open(dst, "w"), indent=1)
    for k, v in sorted(table.items(), key=lambda kv: -kv[1]):
        print("%-34s %8.2f B/SD per launch (%d launches)" % (k, v, len(acc[k])))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), sys.argv[3] if len(sys.argv) > 3 else "")
