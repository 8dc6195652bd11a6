#!/usr/bin/env python
"""Diagnostic (GPU box): how far the single-precision engine is from the reference's float instantiation, scenario by scenario.
Prints one line per scenario and step; used to set the bars of tests/test_gpu_f32.py."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libcloudphxx_b200 import lgrngn as L      # noqa: E402
from tests import support as S                 # noqa: E402

ref, new = S.oracle_library("f32"), S.b200_library("f32")


def report(tag):
    def f(step, p_r, p_n, f_r, f_n):
        n_r, n_n = p_r.get_n(), p_n.get_n()
        line = "%s step %2d  n_part %d/%d  n equal %s" % (tag, step, n_r.size, n_n.size, n_r.size == n_n.size and bool(np.array_equal(n_r, n_n)))
        if n_r.size == n_n.size:
            for a in ("rd3", "rw2", "kappa", "x", "y", "z"):
                a_r, a_n = p_r.get_attr(a), p_n.get_attr(a)
                if a_r.size and a_r.size == a_n.size:
                    line += "  %s %.2e%s" % (a, S.rel_err(a_r, a_n), "=" if np.array_equal(a_r, a_n) else "")
            if step == -1 and not np.array_equal(n_r, n_n):
                rd_r, rd_n = p_r.get_attr("rd3"), p_n.get_attr("rd3")
                bad = np.nonzero(rd_r != rd_n)[0]
                line += "\n    multisets equal: rd3 %s n %s; first mismatch at %d of %d: ref rd3 %r n %d kappa %r | new rd3 %r n %d kappa %r" % (
                    np.array_equal(np.sort(rd_r), np.sort(rd_n)), np.array_equal(np.sort(n_r), np.sort(n_n)), bad[0], bad.size,
                    rd_r[bad[0]], n_r[bad[0]], p_r.get_attr("kappa")[bad[0]], rd_n[bad[0]], n_n[bad[0]], p_n.get_attr("kappa")[bad[0]])
                ijk = None
        for k in ("th", "rv"):
            line += "  %s %.2e" % (k, S.rel_err(f_r[k], f_n[k]))
        print(line, flush=True)
    return f


S.run_pair(ref, new, S.box_golovin, 5, on_step=report("golovin "), n_sd=2 ** 12)
S.run_pair(ref, new, S.parcel, 5, on_step=report("parcel  "), n_sd=2000)
S.run_pair(ref, new, S.box_3d, 4, on_step=report("box3d   "), nx=4, ny=4, nz=6, sd_conc=24, rain_mode=True)
S.run_pair(ref, new, S.box_3d, 3, on_step=report("box3d pc"), nx=6, ny=4, nz=6, sd_conc=16, adve=L.as_t.pred_corr)
S.run_pair(ref, new, S.kinematic_2d, 3, on_step=report("kin2d   "))
