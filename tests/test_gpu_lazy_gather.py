"""Gather-on-read re-layout (LCX_LAZY_GATHER=1): after step_async only positions and storage indices are moved at once;
n, rd3, rw2, kpa, vt stay in the old buffer set until the condensation kernel reads them through the permutation (or any
other consumer completes the gather first).  Data movement only - every result must be bit-identical to the eager path."""
import numpy as np
import pytest

from libcloudphxx_b200 import engine as E
from libcloudphxx_b200 import lgrngn as L
from tests import support as S

pytestmark = pytest.mark.gpu


def run(b200, lazy, monkeypatch, layout, steps=6, poke=False, **kw):
    monkeypatch.setenv("LCX_LAZY_GATHER", "1" if lazy else "0")
    E.set_cond_layout(layout)
    try:
        oi, o, f = S.box_3d(b200, nx=6, ny=5, nz=8, sd_conc=32, rain_mode=True, **kw.pop("box", {}))
        for k, v in kw.pop("opts", {}).items():
            setattr(o, k, v)
        p = b200.factory(L.backend_t.CUDA, oi)
        p.init(f["th"], f["rv"], f["rhod"], None, f["Cx"], f["Cy"], f["Cz"])
        out = []
        for step in range(steps):
            if kw.get("alternate_cond"):
                o.cond = step % 2          # steps without condensation: the fall-speed pass is the first consumer
            p.step_sync(o, f["th"], f["rv"], f["rhod"], f["Cx"], f["Cy"], f["Cz"])
            p.step_async(o)
            if poke and step % 2 == 0:     # a diagnostic between the steps has to complete the pending gather itself
                p.diag_all(); p.diag_wet_mom(3)
                out.append(p.outbuf().copy())
            out.append((p.get_n(), p.get_attr("rw2"), p.get_attr("rd3"), p.get_attr("kappa"), p.get_attr("x"), p.get_attr("z"),
                        f["th"].copy(), f["rv"].copy()) if step == steps - 1 or poke else None)
        return out
    finally:
        E.set_cond_layout(0)


def same(a, b):
    assert len(a) == len(b)
    for x, y in zip(a, b):
        if x is None:
            continue
        if isinstance(x, tuple):
            for u, v in zip(x, y):
                assert np.array_equal(u, v)
        else:
            assert np.array_equal(x, y)


@pytest.mark.parametrize("layout", [16, 3, -1])
@pytest.mark.parametrize("sstp_cond", [1, 3])
def test_lazy_equals_eager(b200, monkeypatch, layout, sstp_cond):
    kw = dict(box=dict(sstp_cond=sstp_cond))
    same(run(b200, False, monkeypatch, layout, **kw), run(b200, True, monkeypatch, layout, **kw))


def test_lazy_with_diagnostics_between_steps(b200, monkeypatch):
    same(run(b200, False, monkeypatch, 16, poke=True), run(b200, True, monkeypatch, 16, poke=True))


def test_lazy_with_steps_without_condensation(b200, monkeypatch):
    same(run(b200, False, monkeypatch, 16, alternate_cond=True), run(b200, True, monkeypatch, 16, alternate_cond=True))


def test_lazy_with_recycling(b200, monkeypatch):
    kw = dict(box=dict(dt=2.0, sstp_coal=2), opts=dict(rcyc=1))
    same(run(b200, False, monkeypatch, 16, **kw), run(b200, True, monkeypatch, 16, **kw))


def test_lazy_against_reference(ref, b200, monkeypatch):
    monkeypatch.setenv("LCX_LAZY_GATHER", "1")
    E.set_cond_layout(16)
    try:
        def check(step, p_r, p_n, f_r, f_n):
            assert np.array_equal(p_r.get_n(), p_n.get_n()), step
            assert np.array_equal(p_r.get_attr("rd3"), p_n.get_attr("rd3")), step
            assert S.rel_err(p_r.get_attr("rw2"), p_n.get_attr("rw2")) < (step + 2) * 2.0 ** -15, step
            assert S.rel_err(f_r["th"], f_n["th"]) < 1e-9 and S.rel_err(f_r["rv"], f_n["rv"]) < 1e-7, step
        S.run_pair(ref, b200, S.box_3d, 5, on_step=check, nx=6, ny=5, nz=8, sd_conc=32, rain_mode=True)
    finally:
        E.set_cond_layout(0)


def test_lazy_paths_are_taken(b200, monkeypatch):
    """the per-kernel profile names the variants that ran: with the opt-in both consumers read through the permutation"""
    from libcloudphxx_b200 import distributed as D
    monkeypatch.setenv("LCX_LAZY_GATHER", "1")
    E.set_cond_layout(16)
    try:
        oi, o, f = S.box_3d(b200, nx=6, ny=5, nz=8, sd_conc=32, rain_mode=True)
        p = b200.factory(L.backend_t.CUDA, oi)
        p.init(f["th"], f["rv"], f["rhod"], None, f["Cx"], f["Cy"], f["Cz"])
        eng = D.engine_of(b200, p)
        p.step_sync(o, f["th"], f["rv"], f["rhod"], f["Cx"], f["Cy"], f["Cz"]); p.step_async(o)
        eng.profile(True)
        for _ in range(2):
            p.step_sync(o, f["th"], f["rv"], f["rhod"], f["Cx"], f["Cy"], f["Cz"]); p.step_async(o)
        names = " ".join(eng.profile_report())
        eng.profile(False)
        assert "k_transport<true>" in names and "true>)" in names.replace("k_transport<true>", ""), names
    finally:
        E.set_cond_layout(0)
