"""Chunked step_sync (host layer: slab::step_sync_chunked; engine: lcx_set_cell_window): th / rv / rhod travel to the device chunk by
chunk, hskpng_Tpr and the run-per-warp condensation kernel work on one chunk of cells while the next one is uploaded and the
previous one's th / rv are read back.  Scheduling only - every result must be bit-identical to the un-chunked step, and the
reference parity of the chunked step is checked once more directly."""
import numpy as np
import pytest

from libcloudphxx_b200 import distributed as D
from libcloudphxx_b200 import engine as E
from libcloudphxx_b200 import lgrngn as L
from tests import support as S

pytestmark = pytest.mark.gpu


def run(lib, monkeypatch, chunks, layout, lazy=True, steps=5, real=np.float64, strided=False, **box):
    monkeypatch.setenv("LCX_SYNC_CHUNKS", str(chunks))
    monkeypatch.setenv("LCX_SYNC_CHUNK_MIN_CELLS", "1")
    monkeypatch.setenv("LCX_LAZY_GATHER", "1" if lazy else "0")
    E.set_cond_layout(layout, "f32" if real == np.float32 else "f64")
    try:
        oi, o, f = S.box_3d(lib, nx=6, ny=5, nz=8, sd_conc=32, rain_mode=True, **box)
        f = {k: np.ascontiguousarray(v, dtype=real) for k, v in f.items()}
        if strided:      # rows padded in z: more than 16 runs, the host layer gathers through its page-locked buffer and must not chunk
            for k in ("th", "rv", "rhod"):
                wide = np.zeros(f[k].shape[:-1] + (f[k].shape[-1] + 3,), dtype=real)
                wide[..., :-3] = f[k]
                f[k] = wide[..., :-3]
        p = lib.factory(L.backend_t.CUDA, oi)
        p.init(f["th"], f["rv"], f["rhod"], None, f["Cx"], f["Cy"], f["Cz"])
        eng = D.engine_of(lib, p)
        l0 = eng.launches()
        for _ in range(steps):
            p.step_sync(o, f["th"], f["rv"], f["rhod"], f["Cx"], f["Cy"], f["Cz"])
            p.step_async(o)
        launches = eng.launches() - l0
        return (p.get_n(), p.get_attr("rw2"), p.get_attr("rd3"), p.get_attr("x"), p.get_attr("z"), np.array(f["th"]), np.array(f["rv"])), launches
    finally:
        E.set_cond_layout(0, "f32" if real == np.float32 else "f64")


def same(a, b):
    for u, v in zip(a, b):
        assert np.array_equal(u, v)


@pytest.mark.parametrize("lazy", [True, False])
@pytest.mark.parametrize("layout,chunks", [(16, 2), (3, 3), (3, 7), (1, 5)])
def test_chunked_equals_whole(b200, monkeypatch, layout, chunks, lazy):
    whole, l_whole = run(b200, monkeypatch, 1, layout, lazy)
    parts, l_parts = run(b200, monkeypatch, chunks, layout, lazy)
    same(whole, parts)
    assert l_parts > l_whole          # the chunked path really ran: hskpng_Tpr + condensation once per chunk


def test_chunking_is_declined_where_it_cannot_apply(b200, monkeypatch):
    """eight lanes per cell (no run-per-warp kernel), sub-stepped condensation, padded arrays: the plain step runs, same results and launches"""
    for kw in (dict(layout=-1), dict(layout=16, sstp_cond=3), dict(layout=16, strided=True)):
        whole, l_whole = run(b200, monkeypatch, 1, **kw)
        parts, l_parts = run(b200, monkeypatch, 4, **kw)
        same(whole, parts)
        assert l_parts == l_whole, kw


def test_chunked_f32_equals_whole(monkeypatch):
    lib = S.b200_library("f32")
    whole, l_whole = run(lib, monkeypatch, 1, 3, real=np.float32)
    parts, l_parts = run(lib, monkeypatch, 3, 3, real=np.float32)
    same(whole, parts)
    assert l_parts > l_whole


def test_chunked_against_reference(ref, b200, monkeypatch):
    monkeypatch.setenv("LCX_SYNC_CHUNKS", "3")
    monkeypatch.setenv("LCX_SYNC_CHUNK_MIN_CELLS", "1")
    E.set_cond_layout(3)
    try:
        def check(step, p_r, p_n, f_r, f_n):
            assert np.array_equal(p_r.get_n(), p_n.get_n()), step
            assert np.array_equal(p_r.get_attr("rd3"), p_n.get_attr("rd3")), step
            assert S.rel_err(p_r.get_attr("rw2"), p_n.get_attr("rw2")) < (step + 2) * 2.0 ** -15, step
            assert S.rel_err(f_r["th"], f_n["th"]) < 1e-9 and S.rel_err(f_r["rv"], f_n["rv"]) < 1e-7, step
        S.run_pair(ref, b200, S.box_3d, 5, on_step=check, nx=6, ny=5, nz=8, sd_conc=32, rain_mode=True)
    finally:
        E.set_cond_layout(0)
