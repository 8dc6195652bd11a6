"""The phase-grouped condensation kernel (k_cond_staged, csrc/lcx_cond.cu) re-orders WHEN a droplet's growth-law evaluations
happen (parked in shared memory between the 3rd, 4th and 5th one), never WHAT is evaluated: every result must be bit-identical
to the plain run-per-warp kernel, with and without the gather-on-read re-layout, for every run length, with sub-stepping, and
for cells from empty to several hundred droplets (queues wrapping many times inside one warp's run)."""
import numpy as np
import pytest

from libcloudphxx_b200 import engine as E
from libcloudphxx_b200 import lgrngn as L
from tests import support as S

pytestmark = pytest.mark.gpu


def run(b200, staged, layout, monkeypatch, lazy="1", steps=6, **box):
    monkeypatch.setenv("LCX_LAZY_GATHER", lazy)
    E.set_cond_layout(layout)
    E.set_cond_staged(staged)
    E.set_cond_classed(0)          # the staged kernel sums a cell's droplets in storage order: compare with the plain kernel doing the same
    try:
        kw = dict(nx=6, ny=5, nz=8, sd_conc=40, rain_mode=True)
        kw.update(box)
        oi, o, f = S.box_3d(b200, **kw)
        p = b200.factory(L.backend_t.CUDA, oi)
        p.init(f["th"], f["rv"], f["rhod"], None, f["Cx"], f["Cy"], f["Cz"])
        out = []
        for _ in range(steps):
            p.step_sync(o, f["th"], f["rv"], f["rhod"], f["Cx"], f["Cy"], f["Cz"])
            p.step_async(o)
            out.append((p.get_n(), p.get_attr("rw2"), p.get_attr("rd3"), f["th"].copy(), f["rv"].copy()))
        return out
    finally:
        E.set_cond_layout(0)
        E.set_cond_staged(False)
        E.set_cond_classed(-1)


def same(a, b):
    for step, (x, y) in enumerate(zip(a, b)):
        for u, v, name in zip(x, y, ("n", "rw2", "rd3", "th", "rv")):
            assert np.array_equal(u, v), (name, step)


@pytest.mark.parametrize("lazy", ["1", "0"])
@pytest.mark.parametrize("layout", [16, 5, 1])
def test_staged_equals_plain(b200, monkeypatch, layout, lazy):
    same(run(b200, False, layout, monkeypatch, lazy), run(b200, True, layout, monkeypatch, lazy))


def test_staged_equals_plain_with_substeps(b200, monkeypatch):
    same(run(b200, False, 16, monkeypatch, sstp_cond=3), run(b200, True, 16, monkeypatch, sstp_cond=3))


def test_staged_equals_plain_with_populous_cells(b200, monkeypatch):
    """300 droplets per cell, runs of 16 cells: ~150 rounds per warp, the queues fill and drain dozens of times"""
    kw = dict(nx=3, ny=3, nz=8, sd_conc=300, steps=3)
    same(run(b200, False, 16, monkeypatch, **kw), run(b200, True, 16, monkeypatch, **kw))


def test_staged_kernel_is_the_one_that_runs(b200, monkeypatch):
    from libcloudphxx_b200 import distributed as D
    monkeypatch.setenv("LCX_LAZY_GATHER", "1")
    E.set_cond_layout(16)
    E.set_cond_staged(True)
    try:
        oi, o, f = S.box_3d(b200, nx=4, ny=4, nz=6, sd_conc=24)
        p = b200.factory(L.backend_t.CUDA, oi)
        p.init(f["th"], f["rv"], f["rhod"], None, f["Cx"], f["Cy"], f["Cz"])
        eng = D.engine_of(b200, p)
        eng.profile(True)
        for _ in range(2):
            p.step_sync(o, f["th"], f["rv"], f["rhod"], f["Cx"], f["Cy"], f["Cz"])
            p.step_async(o)
        rep = eng.profile_report()
        eng.profile(False)
        assert any("k_cond_staged" in k for k in rep), sorted(rep)
    finally:
        E.set_cond_layout(0)
        E.set_cond_staged(False)
