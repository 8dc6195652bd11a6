#!/usr/bin/env python
"""Times the fused per-cell condensation kernel under each work distribution (lcx_set_cond_layout) at the bench size:
one engine, one initialisation, the layout switched between profiled steps.  Prints one line per layout with the mean
duration of the condensation kernel and of the whole resident step (CUDA events around every launch)."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402


def main():
    import torch
    from libcloudphxx_b200 import lgrngn as L, distributed as D, engine as E
    nx, ny, nz = [int(v) for v in (sys.argv[1:4] if len(sys.argv) >= 4 else (64, 256, 128))]
    layouts = [int(v) for v in sys.argv[4:]] or [-1, 4, 8, 16]
    torch.cuda.set_device(0)
    lib = L.b200()
    lib.lib.lgrngn_b200_set_rng_mode.argtypes = [C.c_int]
    lib.lib.lgrngn_b200_set_rng_mode(0)
    oi, o, f = bench.make_case(lib, nx, ny, nz, 40, pin=True)
    p = lib.factory(L.backend_t.CUDA, oi)
    p.init(f["th"], f["rv"], f["rhod"], None, f["Cx"], f["Cy"], f["Cz"])
    eng = D.engine_of(lib, p)
    lib.lib.lgc_proto.restype = C.c_void_p
    lib.lib.lgc_proto.argtypes = [C.c_void_p]
    lib.lib.lgrngn_b200_step_resident.argtypes = [C.c_void_p, C.c_int]
    proto = lib.lib.lgc_proto(p._h)

    def step():
        if lib.lib.lgrngn_b200_step_resident(proto, 0b1111) != 0:
            raise RuntimeError("resident step failed")

    bench.api_step(p, o, f)
    for _ in range(3):
        step()
    out = []
    for rep in range(2):
        for lay in layouts:
            E.set_cond_layout(lay)
            step()
            eng.sync()
            eng.profile(True)
            eng.timer_start()
            for _ in range(3):
                step()
            ms = eng.timer_stop()
            r = eng.profile_report()
            eng.profile(False)
            cond = {k: v for k, v in r.items() if "k_cond" in k}
            name, (n_l, t) = max(cond.items(), key=lambda kv: kv[1][1])
            rec = {"layout": lay, "rep": rep, "kernel": name, "cond_ms": round(t / n_l, 3), "step_ms_profiled": round(ms / 3, 3),
                   "n_part": eng.n_part(), "max_count": eng.cell_stats()[1]}
            out.append(rec)
            print(json.dumps(rec), flush=True)
    E.set_cond_layout(0)
    return out


if __name__ == "__main__":
    main()
