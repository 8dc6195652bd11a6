#!/usr/bin/env python
"""Kinematic 2-D (x, z) prescribed-flow driver around the Lagrangian microphysics: the role `icicle` plays for the reference
(models/kinematic_2D/src/kin_cloud_2d_lgrngn.hpp:128-295, cases/icmw8_case1.hpp:84-218), without libmpdata++.

Per time step, like icicle's hook_post_step: the Eulerian fields th, rv are advected by the stationary single-eddy flow
(psi = -sin(pi z/Z) cos(2 pi x/X), divided by rhod; here with a first-order donor-cell scheme in flux form - mass-conserving, enough
for a driver whose job is to exercise the coupling), then `step_sync(opts, th, rv)` lets the super-droplets condense / evaporate and
writes th, rv back, then `step_async(opts)` moves and collides them - optionally on its own thread while the next Eulerian step
runs, which is how icicle overlaps the two (kin_cloud_2d_lgrngn.hpp:242-277).

The fields live either in host memory (numpy) or in GPU memory (torch CUDA tensors handed over as device pointers: the
device-pointer fast path of arrinfo_t, no PCIe traffic per step).  A caller, not part of the product: it only uses the public
lgrngn API.  `python tools/kinematic_2d.py --steps 200 --device-fields` prints one JSON line with the timing and the budgets.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class Kinematic2D:
    def __init__(self, lib, nx=76, nz=76, sd_conc=64, dt=1.0, w_max=0.6, device_fields=False, async_step=False, backend=None, seed=44):
        from libcloudphxx_b200 import lgrngn as L
        from tests import support as S
        self.L, self.lib = L, lib
        self.device_fields, self.async_step = device_fields, async_step
        # icicle's case: 1.5 km x 1.5 km, one eddy, hydrostatic profile with th_std = 289 K, rv = 7.5 g/kg, bimodal aerosol
        # (icmw8_case1.hpp:84-136, opts_common.hpp:48-62).  Grid: nx x nz full cells here (icicle insets the Lagrangian domain by half
        # a cell because libmpdata++ counts grid points; the physics is the same)
        X, Z = 1500.0, 1500.0
        dx, dz = X / nx, Z / nz
        oi = lib.opts_init_t()
        oi.nx, oi.nz, oi.dx, oi.dz = nx, nz, dx, dz
        oi.x1, oi.z1 = X, Z
        oi.dt = dt
        oi.sd_conc = sd_conc
        oi.n_sd_max = int(nx * nz * sd_conc * 1.3)
        oi.kernel = L.kernel_t.hall_davis_no_waals
        oi.terminal_velocity = L.vt_t.beard77fast
        oi.adve_scheme = L.as_t.implicit
        oi.dry_distros = [L.lognormal(0.61, S.AEROSOL_ICICLE)]
        oi.rng_seed = seed
        o = lib.opts_t()
        self.oi, self.o = oi, o
        self.nx, self.nz, self.dx, self.dz, self.dt = nx, nz, dx, dz, dt
        th_dry, rhod_col, _ = S.hydrostatic_column(nz, dz)
        f = {"th": np.full((nx, nz), th_dry), "rv": np.full((nx, nz), 7.5e-3), "rhod": np.broadcast_to(rhod_col, (nx, nz)).copy()}
        self.rhod = f["rhod"]
        # stream function of the mass flux at the cell corners: psi = -A sin(pi z / Z) cos(2 pi x / X) (icmw8_case1.hpp:84-88); its
        # differences give discretely non-divergent mass-flux Courant numbers GC (closed at z = 0, Z; periodic in x); the particles
        # move with GC / rhod (kin_cloud_2d_lgrngn.hpp:181-196)
        rho_ref = float(rhod_col.mean())
        A = w_max * X / (2 * np.pi) * rho_ref
        xc, zc = np.arange(nx + 1) * dx, np.arange(nz + 1) * dz
        psi = -A * np.sin(np.pi * zc[None, :] / Z) * np.cos(2 * np.pi * xc[:, None] / X)
        self.GCx = -(psi[:, 1:] - psi[:, :-1]) / dz * dt / dx                     # (nx + 1, nz): rho u dt / dx
        self.GCz = (psi[1:, :] - psi[:-1, :]) / dx * dt / dz                      # (nx, nz + 1): rho w dt / dz, zero at the walls
        rho_zface = np.concatenate([[rhod_col[0]], 0.5 * (rhod_col[1:] + rhod_col[:-1]), [rhod_col[-1]]])
        f["Cx"] = np.ascontiguousarray(self.GCx / rhod_col[None, :])
        f["Cz"] = np.ascontiguousarray(self.GCz / rho_zface[None, :])
        self.fields = f
        if device_fields:
            import torch
            self.xp = torch
            dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
            self.th, self.rv, self.rho_d = dev(f["th"]), dev(f["rv"]), dev(f["rhod"])
            self.GCx_d, self.GCz_d = dev(self.GCx), dev(self.GCz)
        else:
            self.xp = np
            self.th, self.rv, self.rho_d = f["th"].copy(), f["rv"].copy(), f["rhod"]
            self.GCx_d, self.GCz_d = self.GCx, self.GCz
        self.p = lib.factory(backend if backend is not None else L.backend_t.CUDA, oi)
        self.p.init(self.th, self.rv, self.rho_d, None, f["Cx"], None, f["Cz"])
        self.worker = None
        self.t_euler = self.t_sync = self.t_async = 0.0

    # ---- Eulerian part: donor-cell advection of a scalar mixing ratio in flux form, periodic in x, closed in z ----------------
    def advect(self, psi):
        xp = self.xp
        GCx, GCz, rho = self.GCx_d, self.GCz_d, self.rho_d
        if xp is np:
            roll = lambda a, s: np.roll(a, s, axis=0)
            pos, neg = (lambda a: np.maximum(a, 0.0)), (lambda a: np.minimum(a, 0.0))
            zeros = lambda n: np.zeros((n, 1))
            cat = lambda parts: np.concatenate(parts, axis=1)
        else:
            roll = lambda a, s: xp.roll(a, s, 0)
            pos, neg = (lambda a: xp.clamp(a, min=0.0)), (lambda a: xp.clamp(a, max=0.0))
            zeros = lambda n: xp.zeros((n, 1), dtype=psi.dtype, device=psi.device)
            cat = lambda parts: xp.cat(parts, 1)
        # x faces 0..nx-1 (face i lies left of cell i; face nx = face 0 by periodicity)
        gx = GCx[:-1]
        fx = pos(gx) * roll(psi, 1) + neg(gx) * psi
        div = roll(fx, -1) - fx
        # z faces 1..nz-1 carry a flux, 0 and nz are walls
        gz = GCz[:, 1:-1]
        fz_in = pos(gz) * psi[:, :-1] + neg(gz) * psi[:, 1:]
        fz = cat([zeros(psi.shape[0]), fz_in, zeros(psi.shape[0])])
        div = div + (fz[:, 1:] - fz[:, :-1])
        return psi - div / rho

    def wait(self):
        if self.worker is not None:
            self.worker.join()
            self.worker = None

    def step(self):
        t0 = time.time()
        th_new, rv_new = self.advect(self.th), self.advect(self.rv)      # may overlap the particles' step_async of the previous step
        if self.device_fields:
            self.xp.cuda.synchronize()
        self.wait()
        if self.xp is np:
            self.th[...], self.rv[...] = th_new, rv_new
        else:
            self.th.copy_(th_new); self.rv.copy_(rv_new)
            self.xp.cuda.synchronize()                        # the library reads the fields on its own stream: hand them over complete
        t1 = time.time()
        self.p.step_sync(self.o, self.th, self.rv)            # th, rv are updated in place by condensation
        t2 = time.time()
        if self.async_step:
            self.worker = threading.Thread(target=self.p.step_async, args=(self.o,))
            self.worker.start()
        else:
            self.p.step_async(self.o)
        t3 = time.time()
        self.t_euler += t1 - t0; self.t_sync += t2 - t1; self.t_async += t3 - t2

    def host(self, a):
        return a if isinstance(a, np.ndarray) else a.cpu().numpy()

    def diagnostics(self):
        """what icicle's diag() records (kin_cloud_2d_lgrngn.hpp:42-118), reduced to the budgets a test can pin"""
        self.wait()
        p, rho = self.p, self.rhod
        shape = (self.nx, self.nz)
        p.diag_wet_rng(0.5e-6, 25e-6); p.diag_wet_mom(3)
        rc = p.outbuf().reshape(shape) * 4.0 / 3.0 * np.pi * 1e3                # cloud water mixing ratio [kg/kg]
        p.diag_wet_rng(25e-6, 1.0); p.diag_wet_mom(3)
        rr = p.outbuf().reshape(shape) * 4.0 / 3.0 * np.pi * 1e3                # rain water
        p.diag_all(); p.diag_wet_mom(3)
        rl = p.outbuf().reshape(shape) * 4.0 / 3.0 * np.pi * 1e3                # all liquid incl. aerosol water
        p.diag_all(); p.diag_sd_conc()
        sd = p.outbuf().reshape(shape)
        p.diag_RH()
        RH = p.outbuf().reshape(shape)
        rv, th = self.host(self.rv), self.host(self.th)
        # dv of the boundary cells is halved by the Lagrangian domain starting half a cell inside; the budget uses the Eulerian cell mass
        mass = rho * self.dx * self.dz
        return {"rc_max": float(rc.max()), "rr_max": float(rr.max()), "RH_max": float(RH.max()), "sd_min": float(sd.min()), "sd_mean": float(sd.mean()),
                "total_water": float(((rv + rl) * mass).sum()), "vapour": float((rv * mass).sum()), "liquid": float((rl * mass).sum()),
                "th_min": float(th.min()), "th_max": float(th.max()), "cloudy_cells": int((rc > 1e-5).sum()),
                "puddle_liquid_volume": float(p.diag_puddle()["liquid_volume"])}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=76)
    ap.add_argument("--nz", type=int, default=76)
    ap.add_argument("--sd-conc", type=int, default=128)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--device-fields", action="store_true")
    ap.add_argument("--async-step", action="store_true")
    args = ap.parse_args()
    from libcloudphxx_b200 import lgrngn as L
    lib = L.b200()
    m = Kinematic2D(lib, args.nx, args.nz, args.sd_conc, device_fields=args.device_fields, async_step=args.async_step)
    d0 = m.diagnostics()
    for _ in range(5):
        m.step()
    m.wait()
    m.t_euler = m.t_sync = m.t_async = 0.0
    t0 = time.time()
    for _ in range(args.steps):
        m.step()
    m.wait()
    wall = time.time() - t0
    d1 = m.diagnostics()
    n_sd = args.nx * args.nz * args.sd_conc
    print(json.dumps({"model": "kinematic_2d (icicle set-up, donor-cell Eulerian part)", "grid": [args.nx, args.nz], "sd_conc": args.sd_conc,
                      "steps": args.steps, "device_fields": args.device_fields, "async_step": args.async_step,
                      "ms_per_step": 1e3 * wall / args.steps, "sd_updates_per_s": n_sd * args.steps / wall,
                      "ms_euler": 1e3 * m.t_euler / args.steps, "ms_step_sync": 1e3 * m.t_sync / args.steps, "ms_step_async": 1e3 * m.t_async / args.steps,
                      "start": d0, "end": d1, "water_budget_rel_err": abs(d1["total_water"] - d0["total_water"]) / d0["total_water"]}))


if __name__ == "__main__":
    main()
