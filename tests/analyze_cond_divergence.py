#!/usr/bin/env python
"""How much lock-step work does the condensation kernel lose to droplets of one warp needing different numbers of
growth-law evaluations?  CPU-only analysis with the oracle restatement (oracle/sdm_port.py): a bench-shaped column
(hydrostatic, supersaturated upper half, two-mode aerosol, 40 SDs per cell, Cx = 0.1, Cy = 0.05) is stepped a few times,
the evaluations of drw2/dt are counted per droplet in the last condensation step, and the cost of a warp round
(max over its 32 lanes) is compared between orderings of the droplets:
  storage    - cells contiguous, inside a cell the order the re-layout leaves (stayers in old order, then arrivals)
  size class - the same windows of `run` cells, droplets ordered by a coarse size class first (what a warp could do itself)
  ideal      - ordered by the evaluation count itself (lower bound)
  previous   - ordered by the count the droplet needed in the previous step (a 1-byte record a kernel could carry along)
  driving force - ordered by RH - a_w(rw, rd, kappa) exp(A / rw), which a kernel can compute itself before the solve
Test infrastructure (lives under tests/ because it uses oracle/); never shipped, not collected by pytest."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))      # tests/ is one of the places allowed to use oracle/
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import sdm_port as port  # noqa: E402
from tests import support as S  # noqa: E402
from tests.test_cpu_oracle import lognormal_as_capi  # noqa: E402


def main(nx=2, ny=2, nz=128, sd_conc=40, steps=8, run=16):
    eff = np.fromfile(os.path.join(ROOT, "libcloudphxx_b200", "data", "hall_davis_no_waals.f64"))
    th_dry, rhod_col, _ = S.hydrostatic_column(nz, 20.0)
    f = {"th": np.full((nx, ny, nz), th_dry), "rv": np.full((nx, ny, nz), 6e-3), "rhod": np.broadcast_to(rhod_col, (nx, ny, nz)).copy(),
         "Cx": np.full((nx + 1, ny, nz), 0.1), "Cy": np.full((nx, ny + 1, nz), 0.05), "Cz": np.zeros((nx, ny, nz + 1))}
    f["rv"][:, :, nz // 2:] = 8.2e-3
    p = port.Particles(nx=nx, ny=ny, nz=nz, dx=20., dy=20., dz=20., dt=1., x1=nx * 20., y1=ny * 20., z1=nz * 20., sd_conc=sd_conc,
                       n_sd_max=int(nx * ny * nz * sd_conc * 1.5), kernel="efficiencies", kernel_params={"eff": eff[1:], "r_max": eff[0]},
                       dry_distros=[(0.61, lognormal_as_capi(S.AEROSOL_ICICLE))])
    p.init(f["th"], f["rv"], f["rhod"], f["Cx"], f["Cy"], f["Cz"])

    # physical order of the B200 engine: grouped by cell; stayers keep their order, arrivals are appended in old physical order.
    # (no SD is removed in this set-up, so the port's storage index identifies an SD throughout)
    n0 = p.n_part
    phys = np.argsort(p.ijk[:n0], kind="stable")       # initial grouping: storage order inside each cell (sorted by dry size)
    cell_of = p.ijk[:n0].copy()

    counter = {"n": 0}
    orig = port.drw2_dt

    def counted(*a):
        counter["n"] += 1
        return orig(*a)
    port.drw2_dt = counted

    evals = prev_evals = None
    for step in range(steps):
        last = step == steps - 1
        if last:                                         # what a kernel could compute before it orders its run: the driving force RH - a_w * Kelvin
            p.hskpng_Tpr()
            n_ = p.n_part
            rw_ = np.sqrt(p.rw2[:n_])
            a_w_ = (rw_ ** 3 - p.rd3[:n_]) / (rw_ ** 3 - p.rd3[:n_] * (1 - p.kpa[:n_]))
            diseq = p.RH[p.ijk[:n_]] - a_w_ * np.exp(np.array([port.kelvin_A(t) for t in p.T[p.ijk[:n_]]]) / rw_)
        if last or step == steps - 2:
            # per-droplet evaluation counts: wrap advance_rw2
            per = np.zeros(p.n_part, dtype=np.int64)
            adv = port.advance_rw2
            idx = {"s": 0}

            def adv_counted(*a, **k):
                c0 = counter["n"]
                r = adv(*a, **k)
                per[idx["s"]] = counter["n"] - c0
                idx["s"] += 1
                return r
            port.advance_rw2 = adv_counted
        th, rv = p.step_sync(f["th"], f["rv"], f["rhod"])
        if last:
            port.advance_rw2 = adv
            evals = per
            break
        if step == steps - 2:
            port.advance_rw2 = adv
            prev_evals = per                             # storage order is stable (no SD is removed in this set-up)
        f["th"][:], f["rv"][:] = th.reshape(f["th"].shape), rv.reshape(f["rv"].shape)
        n_before = p.n_part
        p.step_async()
        assert p.n_part == n_before, "an SD was removed: the layout emulation below assumes none are (no rain in this set-up)"
        new_cell = p.ijk[:p.n_part].copy()
        stay = new_cell[phys] == cell_of[phys]
        order_stay = phys[stay]
        order_move = phys[~stay]
        # per cell: stayers (old order) then arrivals (old physical order)
        key = np.concatenate([new_cell[order_stay] * 2, new_cell[order_move] * 2 + 1])
        allp = np.concatenate([order_stay, order_move])
        phys = allp[np.argsort(key, kind="stable")]
        cell_of = new_cell
        print("step %d: %.1f %% changed cell" % (step, 100.0 * (~stay).mean()), flush=True)

    n = p.n_part
    ev = evals[phys].astype(float)                     # evaluation counts in physical order
    rw2 = p.rw2[:n][phys]
    cells = cell_of[phys]
    print("droplets %d, evaluations per droplet: mean %.2f, max %d; histogram %s" % (n, ev.mean(), ev.max(), np.bincount(evals)[:16]))

    def cost(order):                                   # sum over warp rounds of the max evaluation count, per droplet
        e = ev[order]
        pad = (-len(e)) % 32
        e = np.concatenate([e, np.zeros(pad)])
        return 32.0 * e.reshape(-1, 32).max(axis=1).sum() / len(order)

    ident = np.arange(n)
    print("lock-step cost (evaluations per droplet, max over the 32 lanes of a round):")
    print("  mean per droplet (no divergence)         %.2f" % ev.mean())
    print("  storage order, 32 consecutive            %.2f" % cost(ident))
    # old kernel: 8 lanes per cell, 4 cells per warp (round r: in-cell positions 8r..8r+7 of four consecutive cells)
    n_cell = nx * ny * nz
    off = np.searchsorted(cells, np.arange(n_cell + 1))
    tot = 0.0
    for c0 in range(0, n_cell, 4):
        cnts = [off[c + 1] - off[c] for c in range(c0, min(c0 + 4, n_cell))]
        for r in range(max((k + 7) // 8 for k in cnts)):
            m = 0.0
            for q, c in enumerate(range(c0, min(c0 + 4, n_cell))):
                seg = ev[off[c] + 8 * r: min(off[c] + 8 * r + 8, off[c + 1])]
                if len(seg):
                    m = max(m, seg.max())
            tot += m * 32
    print("  8 lanes per cell, 4 cells per warp        %.2f  (idle lanes of short cells included)" % (tot / n))
    for bits, name in ((3, "factor 4 in radius"), (5, "factor 2"), (8, "factor 1.19")):
        e2 = np.floor(np.log2(np.maximum(rw2, 1e-30)) * {3: 0.25, 5: 0.5, 8: 2.0}[bits]).astype(np.int64)   # class by exponent of rw2
        order = []
        for c0 in range(0, n_cell, run):
            a, b = off[c0], off[min(c0 + run, n_cell)]
            order.append(a + np.argsort(e2[a:b], kind="stable"))
        print("  runs of %d cells ordered by size class (%s) %.2f" % (run, name, cost(np.concatenate(order))))
    order = []
    for c0 in range(0, n_cell, run):
        a, b = off[c0], off[min(c0 + run, n_cell)]
        order.append(a + np.argsort(ev[a:b], kind="stable"))
    print("  runs of %d cells ordered by the count itself %.2f" % (run, cost(np.concatenate(order))))
    dq = diseq[phys]
    order = []
    for c0 in range(0, n_cell, run):
        a, b = off[c0], off[min(c0 + run, n_cell)]
        order.append(a + np.argsort(dq[a:b], kind="stable"))
    print("  runs of %d cells ordered by the driving force RH - a_w exp(A/rw) (stateless, ~1/5 of one growth-law evaluation) %.2f"
          % (run, cost(np.concatenate(order))))
    if prev_evals is not None:
        pe = prev_evals[phys]
        print("  the previous step's count predicts this step's exactly for %.1f %% of the droplets, within one for %.1f %%"
              % (100. * (pe == ev).mean(), 100. * (np.abs(pe - ev) <= 1).mean()))
        for r in (run, 8):
            order = []
            for c0 in range(0, n_cell, r):
                a, b = off[c0], off[min(c0 + r, n_cell)]
                order.append(a + np.argsort(pe[a:b], kind="stable"))
            print("  runs of %d cells ordered by the PREVIOUS step's count %.2f" % (r, cost(np.concatenate(order))))


if __name__ == "__main__":
    main(*[int(v) for v in sys.argv[1:]])
