"""Class-ordered condensation (k_cond_classed, csrc/lcx_cond.cu): the droplets of a warp's run are walked class by class - drizzle / rain
drops (rw > 40 um) apart from the rest - instead of in storage order, so that the expensive branches of the ventilation factors run on
full warps.  Same arithmetic per droplet: one step from the same state gives bit-identical wet radii; only the order in which a cell's droplets are
summed differs, so th / rv agree to rounding (and later steps follow within that).  The automatic mode must pick the variant from the
previous step's count alone (same sequence in every run, chunked or not)."""
import numpy as np
import pytest

from libcloudphxx_b200 import distributed as D
from libcloudphxx_b200 import engine as E
from libcloudphxx_b200 import lgrngn as L
from tests import support as S

pytestmark = pytest.mark.gpu


def run(b200, classed, layout, monkeypatch, lazy="1", steps=6, chunks=1, **box):
    monkeypatch.setenv("LCX_LAZY_GATHER", lazy)
    monkeypatch.setenv("LCX_SYNC_CHUNKS", str(chunks))
    monkeypatch.setenv("LCX_SYNC_CHUNK_MIN_CELLS", "1")
    E.set_cond_layout(layout)
    E.set_cond_classed(classed)
    try:
        kw = dict(nx=6, ny=5, nz=8, sd_conc=40, rain_mode=True)
        kw.update(box)
        oi, o, f = S.box_3d(b200, **kw)
        p = b200.factory(L.backend_t.CUDA, oi)
        p.init(f["th"], f["rv"], f["rhod"], None, f["Cx"], f["Cy"], f["Cz"])
        eng = D.engine_of(b200, p)
        out, names = [], []
        for _ in range(steps):
            eng.profile(True)
            p.step_sync(o, f["th"], f["rv"], f["rhod"], f["Cx"], f["Cy"], f["Cz"])
            names.append(sorted(k for k in eng.profile_report() if "k_cond" in k))
            eng.profile(False)
            p.step_async(o)
            out.append((p.get_n(), p.get_attr("rw2"), p.get_attr("rd3"), f["th"].copy(), f["rv"].copy()))
        return out, names
    finally:
        E.set_cond_layout(0)
        E.set_cond_classed(-1)


@pytest.mark.parametrize("lazy", ["1", "0"])
@pytest.mark.parametrize("layout", [16, 5, 1])
def test_classed_equals_storage_order(b200, monkeypatch, layout, lazy):
    (a, na), (b, nb) = run(b200, 0, layout, monkeypatch, lazy), run(b200, 1, layout, monkeypatch, lazy)
    assert all("classed" not in k for names in na for k in names) and all(any("classed" in k for k in names) for names in nb)
    for step, (x, y) in enumerate(zip(a, b)):
        assert np.array_equal(x[0], y[0]) and np.array_equal(x[2], y[2]), step
        # one step from the same state: the same wet radii bit for bit; afterwards th / rv differ in their last bits and the radii follow
        # (a droplet whose last trial point hits the root to rounding accuracy may land on the other side of it: the condensation tolerance)
        assert np.array_equal(x[1], y[1]) if step == 0 else (S.rel_err(x[1], y[1]) < (step + 1) * 2.0 ** -15 and np.median(np.abs(x[1] / y[1] - 1)) < 1e-13), step
        # th, rv: rounding of the per-cell sums after one step; later the few droplets that landed on the other side of their root show
        assert S.rel_err(x[3], y[3]) < (1e-13 if step == 0 else 1e-10) and S.rel_err(x[4], y[4]) < (1e-12 if step == 0 else 1e-8), step


def test_classed_with_populous_cells_and_substeps(b200, monkeypatch):
    """300 droplets per cell: runs of 16 cells exceed the order list (1024) and keep the storage order, runs of 3 cells fit"""
    for layout in (16, 3):
        kw = dict(nx=3, ny=3, nz=8, sd_conc=300, steps=3, sstp_cond=2)
        (a, _), (b, _) = run(b200, 0, layout, monkeypatch, **kw), run(b200, 1, layout, monkeypatch, **kw)
        for x, y in zip(a, b):
            assert np.array_equal(x[0], y[0]) and S.rel_err(x[1], y[1]) < 4 * 2.0 ** -15 and np.median(np.abs(x[1] / y[1] - 1)) < 1e-13
            assert S.rel_err(x[3], y[3]) < 1e-10 and S.rel_err(x[4], y[4]) < 1e-8


def test_automatic_mode_switches_on_the_previous_steps_count_and_is_reproducible(b200, monkeypatch):
    (a, na), (b, nb) = run(b200, -1, 16, monkeypatch), run(b200, -1, 16, monkeypatch)
    assert na == nb
    assert all("classed" not in k for k in na[0]), "nothing has been counted before the first step"
    assert any("classed" in k for k in na[-1]), "a third of this box's droplets are rain-mode drops: the class order must have been chosen"
    for x, y in zip(a, b):
        for u, v in zip(x, y):
            assert np.array_equal(u, v)
    # chunked step_sync sees the same count, takes the same variants, gives the same bits
    (c, nc) = run(b200, -1, 3, monkeypatch, chunks=3)
    (d, nd) = run(b200, -1, 3, monkeypatch, chunks=1)
    for x, y in zip(c, d):
        for u, v in zip(x, y):
            assert np.array_equal(u, v)


def test_automatic_mode_keeps_storage_order_without_large_drops(b200, monkeypatch):
    _, names = run(b200, -1, 16, monkeypatch, rain_mode=False, steps=4)
    assert all("classed" not in k for step in names for k in step)


def test_classed_against_reference(ref, b200, monkeypatch):
    E.set_cond_layout(16)
    E.set_cond_classed(1)
    try:
        def check(step, p_r, p_n, f_r, f_n):
            assert np.array_equal(p_r.get_n(), p_n.get_n()), step
            assert S.rel_err(p_r.get_attr("rw2"), p_n.get_attr("rw2")) < (step + 2) * 2.0 ** -15, step
            assert S.rel_err(f_r["th"], f_n["th"]) < 1e-9 and S.rel_err(f_r["rv"], f_n["rv"]) < 1e-7, step
        S.run_pair(ref, b200, S.box_3d, 5, on_step=check, nx=6, ny=5, nz=8, sd_conc=32, rain_mode=True)
    finally:
        E.set_cond_layout(0)
        E.set_cond_classed(-1)
