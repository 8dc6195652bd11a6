"""The configuration users get by default and bench.py times: Philox4x32-10 evaluated inside the coalescence kernels,
storage indices re-numbered lazily, and the device-resident step (lgrngn_b200_step_resident).

The reference has no counter-based generator, so the tie to the oracle goes in two links:
  1. the injected-stream path is bit-exact against the reference under its own mt19937 stream (tests/test_gpu_parity.py);
  2. here the Philox stream is restated on the host (numpy, tests/support.py philox_streams_for_layout), fed through that
     same injected path, and must give bit-identical multiplicities, dry and wet radii as the in-kernel generator - for the
     per-cell kernel (k_coal_small) and the global-sort kernel (k_coal_big).
Plus: resident step == step_sync + step_async, lazy == dense storage indices, and the reference's own statistical checks
(Golovin's analytic solution, Bott's spectrum) under Philox, conservation, uniqueness of storage indices."""
import os

import numpy as np
import pytest

from libcloudphxx_b200 import lgrngn as L
from tests import support as S

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def state(p):
    return p.get_n(), p.get_attr("rd3"), p.get_attr("rw2"), p.get_attr("kappa")


def make(lib, setup, mode, dense, **kw):
    with S.rng_mode(lib, mode, dense):
        oi, o, f = setup(lib, **kw)
        p = lib.factory(L.backend_t.CUDA, oi)
    p.init(f["th"], f["rv"], f["rhod"], None, f.get("Cx"), f.get("Cy"), f.get("Cz"))
    return oi, o, f, p


def api_step(p, o, f, rhod=True):
    p.step_sync(o, f["th"], f["rv"], f["rhod"] if rhod else None, f.get("Cx"), f.get("Cy"), f.get("Cz"))
    p.step_async(o)


@pytest.mark.parametrize("case", ["small_cells_3d", "small_cells_3d_substeps", "big_cell_0d"])
def test_in_kernel_philox_equals_host_philox_through_the_injected_path(b200, case, monkeypatch):
    monkeypatch.setenv("LCX_DEVICE_INIT", "0")                 # both runs start from the host-made (mt19937) initial state
    if case == "big_cell_0d":
        setup, kw, small, steps = S.box_golovin, dict(n_sd=2 ** 12, dt=20.0), False, 12
    else:
        setup, small, steps = S.box_3d, True, 6
        kw = dict(nx=5, ny=4, nz=6, sd_conc=48, rain_mode=True, sstp_coal=3 if case.endswith("substeps") else 1)
    oi_a, o_a, f_a, A = make(b200, setup, 0, 1, **kw)          # Philox in the kernels, dense storage indices
    oi_b, o_b, f_b, B = make(b200, setup, 1, -1, **kw)         # injected streams
    if setup is S.box_3d:
        for o in (o_a, o_b):
            o.cond = 0                                         # keeps the two runs on one code path besides the generator
    cap = int(oi_a.n_sd_max)
    collisions = 0
    for step in range(steps):
        sid, ijk = S.physical_layout(b200, B, cap)
        assert np.array_equal(np.sort(sid), np.arange(sid.size)), "storage indices are not a permutation"
        sa, _ = S.physical_layout(b200, A, cap)
        assert np.array_equal(sa, sid), "the two runs lie differently in memory at step %d" % step
        call = S.philox_call(b200, A)
        assert call == S.philox_call(b200, B)
        for sub in range(int(oi_a.sstp_coal)):
            un, u01 = S.philox_streams_for_layout(sid, ijk, int(oi_a.rng_seed), call + sub, small)
            S.inject_rng(b200, B, un, u01)
        n_before = A.get_n()
        api_step(A, o_a, f_a)
        api_step(B, o_b, f_b)
        for x, y, name in zip(state(A), state(B), ("n", "rd3", "rw2", "kappa")):
            assert np.array_equal(x, y), (name, step)
        na = A.get_n()
        collisions += int(na.size != n_before.size or not np.array_equal(na, n_before))
    assert collisions > steps // 2, "hardly any collision happened - the comparison would be vacuous"


@pytest.mark.parametrize("sstp_cond", [1, 2])
def test_resident_step_equals_api_step(b200, sstp_cond):
    """lgrngn_b200_step_resident (no field traffic) == step_sync + step_async on the fields the previous step left"""
    kw = dict(nx=6, ny=5, nz=8, sd_conc=32, rain_mode=True, sstp_cond=sstp_cond)
    _, o_a, f_a, A = make(b200, S.box_3d, 0, -1, **kw)
    _, o_b, f_b, B = make(b200, S.box_3d, 0, -1, **kw)
    api_step(A, o_a, f_a, rhod=False)
    api_step(B, o_b, f_b, rhod=False)                          # the first step uploads the fields
    for step in range(5):
        api_step(A, o_a, f_a, rhod=False)                      # th, rv written back by the previous step: the same values again
        S.step_resident(b200, B)
    for x, y in zip(state(A) + (A.get_attr("x"), A.get_attr("z")), state(B) + (B.get_attr("x"), B.get_attr("z"))):
        assert np.array_equal(x, y)
    B.diag_all(); B.diag_wet_mom(3); m_b = B.outbuf().copy()
    A.diag_all(); A.diag_wet_mom(3); m_a = A.outbuf().copy()
    assert np.array_equal(m_a, m_b)


def test_lazy_and_dense_storage_indices_agree(b200):
    """the per-cell kernel only uses storage indices to break ties of equal random keys; re-numbering keeps their order, so
    postponing it (what Philox runs do) must not change anything - with removals, rain-out and recycling of memory slots"""
    kw = dict(nx=6, ny=5, nz=8, sd_conc=40, rain_mode=True, cx=0.4)
    res = []
    for dense in (1, 0):
        oi, o, f, p = make(b200, S.box_3d, 0, dense, **kw)
        n0 = p.get_n().size
        for _ in range(8):
            api_step(p, o, f)
        sid, _ = S.physical_layout(b200, p, int(oi.n_sd_max))
        assert np.unique(sid).size == sid.size, "storage indices must stay unique"
        res.append(state(p) + (f["th"].copy(), f["rv"].copy()))
        assert res[-1][0].size < n0, "nothing was removed - the test would be vacuous"
    for x, y in zip(*res):
        assert np.array_equal(x, y)


def test_philox_conserves_and_keeps_counts(b200):
    """full step under Philox: SD count + dry volume accounted for by the puddle, multiplicities positive"""
    oi, o, f, p = make(b200, S.box_3d, 0, -1, nx=6, ny=5, nz=8, sd_conc=40, rain_mode=True)
    dv_rhod = 20.0 ** 3 * f["rhod"]

    def dry_volume():
        p.diag_all(); p.diag_dry_mom(3)
        return float((p.outbuf().reshape(6, 5, 8) * dv_rhod).sum()) * 4. / 3 * np.pi
    v0 = dry_volume()
    for _ in range(10):
        api_step(p, o, f)
    assert abs(dry_volume() + p.diag_puddle()["dry_volume"] - v0) <= 1e-10 * v0
    assert (p.get_n() > 0).all()


@pytest.mark.parametrize("dense", [1, 0])
def test_golovin_analytic_under_philox(b200, dense):
    """tests/python/physics/coalescence_golovin.py:112-155 with the in-kernel generator (global-sort kernel, 800 sub-steps)"""
    from tests.test_gpu_fixtures import golovin_analytic_rmsd
    with S.rng_mode(b200, 0, dense):
        assert golovin_analytic_rmsd(b200) < 1.2e-5


@pytest.mark.parametrize("vt", [L.vt_t.beard77fast])
def test_hall_davis_coalescence_vs_bott_under_philox(b200, vt):
    from tests.test_gpu_fixtures import bott_rmsd
    with S.rng_mode(b200, 0, -1):
        assert bott_rmsd(b200, vt) < 6e-2
