#!/usr/bin/env python
"""Extracts the tabulated collision efficiencies used by the `hall*` / `vohl*` coalescence kernels.

The tables are physical data (Hall 1980; Davis 1972; Pinsky et al. 2001; Vohl et al. 2007), stored by the
reference as packed lower-triangular matrices on a (1 um up to 100 um, 10 um above) radius grid
(reference src/detail/kernel_definitions/*_efficiencies.hpp, indexing src/detail/kernel_utils.hpp:12-29).
A drop-in back-end must interpolate the very same numbers, so they are shipped as data:
    libcloudphxx_b200/data/<name>.f64 = float64[1 + N]: r_max [um] followed by the N table values.
Run here (needs /root/reference); the outputs are committed.
"""
import os
import re
import sys

import numpy as np

REF = os.environ.get("LCX_REFERENCE_ROOT", "/root/reference")
SRC = os.path.join(REF, "src", "detail", "kernel_definitions")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "libcloudphxx_b200", "data")

NAMES = ["hall", "hall_davis_no_waals", "vohl_davis_no_waals", "hall_pinsky_stratocumulus",
         "hall_pinsky_cumulonimbus", "hall_pinsky_1000mb_grav"]


def main():
    os.makedirs(OUT, exist_ok=True)
    for name in NAMES:
        text = open(os.path.join(SRC, name + "_efficiencies.hpp")).read()
        r_max = float(re.search(r"%s_r_max\(\)\s*\{\s*return\s+([0-9.eE+-]+)\s*;" % name, text).group(1))
        body = text[text.index("arr[] = {") + len("arr[] = {"): text.index("};", text.index("arr[] = {"))]
        vals = np.array([float(v) for v in body.replace("\n", " ").split(",") if v.strip()], dtype=np.float64)
        out = np.concatenate([[r_max], vals])
        out.tofile(os.path.join(OUT, name + ".f64"))
        print("%-28s r_max = %6.0f um, %d values" % (name, r_max, vals.size))


if __name__ == "__main__":
    sys.exit(main())
