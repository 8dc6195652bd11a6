#!/usr/bin/env python
"""Generates the committed fixtures under tests/golden/ (run in the build container, where /root/reference exists).

  lgrngn_cond_substepping_percell.csv   the 56 per-cell-substepping rows (exact_sstp = False) of the reference's own fixture
  lgrngn_cond_substepping_perparticle.csv   its 112 per-particle rows without adaptation (exact_sstp = True, mixing on / off)
  lgrngn_cond_substepping_adaptive.csv      its 112 adaptive per-particle rows (sstp_cond_act 1 and 8)
                                        tests/python/physics/refdata/lgrngn_cond_substepping_refdata.csv
  bott1800.npy                          the 149-value Bott bin-model mass-density array embedded in the reference's
                                        tests/python/physics/coalescence_hall_davis_no_waals.py:82
  ref_golovin_box.npz                   per-step state of a 0-D Golovin box (2^10 SDs, 12 steps) from oracle/_ref
  ref_box3d.npz                         per-step state of a 4x3x4-cell full-microphysics box (3 steps) from oracle/_ref
The last two are outputs of the reference's own serial back-end built from the unmodified sources (oracle/build_ref.py).
"""
import csv
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("LCX_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")

from libcloudphxx_b200 import lgrngn as L   # noqa: E402
from tests import support as S              # noqa: E402


def main():
    os.makedirs(OUT, exist_ok=True)
    src = os.path.join(REF, "tests", "python", "physics", "refdata", "lgrngn_cond_substepping_refdata.csv")
    with open(src) as fh:
        rows = [r for r in csv.DictReader(fh) if r["exact_sstp"] == "False"]
    with open(os.path.join(OUT, "lgrngn_cond_substepping_percell.csv"), "w", newline="") as fh:
        w = csv.DictWriter(fh, fieldnames=list(rows[0].keys()))
        w.writeheader()
        w.writerows(rows)
    print("cond substepping rows:", len(rows))
    with open(src) as fh:
        rows = [r for r in csv.DictReader(fh) if r["exact_sstp"] == "True" and r["adaptive"] == "False"]
    with open(os.path.join(OUT, "lgrngn_cond_substepping_perparticle.csv"), "w", newline="") as fh:
        w = csv.DictWriter(fh, fieldnames=list(rows[0].keys()))
        w.writeheader()
        w.writerows(rows)
    print("per-particle (non-adaptive) cond substepping rows:", len(rows))
    with open(src) as fh:
        rows = [r for r in csv.DictReader(fh) if r["adaptive"] == "True"]
    with open(os.path.join(OUT, "lgrngn_cond_substepping_adaptive.csv"), "w", newline="") as fh:
        w = csv.DictWriter(fh, fieldnames=list(rows[0].keys()))
        w.writeheader()
        w.writerows(rows)
    print("adaptive per-particle cond substepping rows:", len(rows))

    text = open(os.path.join(REF, "tests", "python", "physics", "coalescence_hall_davis_no_waals.py")).read()
    arr = re.search(r"bott1800 = np.array\(\[(.*?)\]\)", text, re.S).group(1)
    bott = np.array([float(v) for v in arr.split(",")])
    np.save(os.path.join(OUT, "bott1800.npy"), bott)
    print("bott1800 values:", bott.size)

    ref = S.oracle_library()
    rec = {}

    def grab(tag):
        def f(step, p_r, *_):
            rec["%s_n_%d" % (tag, step + 1)] = p_r.get_n()
            for k in ("rw2", "rd3", "x", "y", "z"):
                a = p_r.get_attr(k)
                if a.size:
                    rec["%s_%s_%d" % (tag, k, step + 1)] = a
        return f

    oi, o, f = S.box_golovin(ref, n_sd=2 ** 10)
    p = ref.factory(L.backend_t.serial, oi)
    p.init(f["th"], f["rv"], f["rhod"])
    g = grab("g")
    g(-1, p)
    for step in range(12):
        p.step_sync(o, f["th"], f["rv"], f["rhod"]); p.step_async(o)
        g(step, p)
    np.savez_compressed(os.path.join(OUT, "ref_golovin_box.npz"), **rec)

    rec.clear()
    oi, o, f = S.box_3d(ref, nx=4, ny=3, nz=4, sd_conc=8, rain_mode=True)
    p = ref.factory(L.backend_t.serial, oi)
    p.init(f["th"], f["rv"], f["rhod"], None, f["Cx"], f["Cy"], f["Cz"])
    g = grab("b")
    g(-1, p)
    for step in range(3):
        p.step_sync(o, f["th"], f["rv"], f["rhod"], f["Cx"], f["Cy"], f["Cz"]); p.step_async(o)
        g(step, p)
        rec["b_th_%d" % (step + 1)] = f["th"].copy()
        rec["b_rv_%d" % (step + 1)] = f["rv"].copy()
    np.savez_compressed(os.path.join(OUT, "ref_box3d.npz"), **rec)
    print("golden state vectors written")


if __name__ == "__main__":
    main()
