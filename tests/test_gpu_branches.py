"""Branches of the CUDA kernels that round 1 only exercised through the numpy restatement: large-scale subsidence, open side
walls, periodic top / bottom walls, the non-spherical Khvorostyanov fall speed, the Pinsky and Vohl efficiency tables, and the
per-step opts.dt override.  Same bars as tests/test_gpu_parity.py: integer state and + - * / arithmetic exact under the
replayed random stream, anything through exp / log / pow within a few ulp."""
import numpy as np
import pytest

from libcloudphxx_b200 import lgrngn as L
from tests import support as S

pytestmark = pytest.mark.gpu


def exact_transport(step, p_r, p_n, tol_z=0.0):
    n_r, n_n = p_r.get_n(), p_n.get_n()
    assert n_r.size == n_n.size, (step, n_r.size, n_n.size)
    assert np.array_equal(n_r, n_n), step
    for k in ("x", "y"):
        assert np.array_equal(p_r.get_attr(k), p_n.get_attr(k)), (k, step)
    if tol_z:
        assert S.rel_err(p_r.get_attr("z"), p_n.get_attr("z")) < tol_z, step
    else:
        assert np.array_equal(p_r.get_attr("z"), p_n.get_attr("z")), step


@pytest.mark.parametrize("scheme", [L.as_t.implicit, L.as_t.euler, L.as_t.pred_corr])
def test_subsidence_and_open_side_walls(ref, b200, scheme):
    """subs.ipp:13-25 (w_LS of the level the SD was last indexed in), bcnd.ipp:126-142,202-216 (SDs leaving through x or y are
    removed): survivors and positions bit-identical"""
    sizes = []

    def setup(lib):
        oi, o, f = S.box_3d(lib, nx=5, ny=4, nz=6, sd_conc=16, adve=scheme, cx=0.4, cy=-0.3)
        oi.subs_switch, oi.w_LS, oi.open_side_walls = 1, list(0.5 + 0.3 * np.arange(6)), 1
        o.cond = o.coal = o.sedi = 0
        o.subs = 1
        return oi, o, f

    def check(step, p_r, p_n, f_r, f_n):
        exact_transport(step, p_r, p_n)
        sizes.append(p_r.get_n().size)
    S.run_pair(ref, b200, setup, 6, on_step=check)
    assert 0 < sizes[-1] < sizes[0], "nothing left through the open walls"


def test_subsidence_with_sedimentation_and_puddle(ref, b200):
    def setup(lib):
        oi, o, f = S.box_3d(lib, nx=4, ny=4, nz=6, sd_conc=24, rain_mode=True)
        oi.subs_switch, oi.w_LS = 1, list(0.2 + 0.1 * np.arange(6))
        o.cond = o.coal = 0
        o.subs = 1
        return oi, o, f

    def check(step, p_r, p_n, f_r, f_n):
        exact_transport(step, p_r, p_n, tol_z=1e-11)
        if step >= 0:
            a, b = p_r.diag_puddle(), p_n.diag_puddle()
            for k in a:
                assert abs(a[k] - b[k]) <= 1e-12 * abs(a[k]), (k, step)
    p_r, p_n, _, _ = S.run_pair(ref, b200, setup, 8, on_step=check)
    assert p_r.diag_puddle()["particle_number"] > 0, "nothing rained out"


def test_periodic_top_and_bottom_walls(ref, b200):
    """opts_init.periodic_topbot_walls (bcnd.ipp:221-245): SDs sedimenting through z0 re-enter at the top; nothing is removed"""
    def setup(lib):
        oi, o, f = S.box_3d(lib, nx=4, ny=4, nz=5, sd_conc=24, rain_mode=True)
        oi.periodic_topbot_walls = 1
        f["Cz"][:] = -0.35
        o.cond = o.coal = 0
        return oi, o, f
    sizes, wrapped = [], []

    def check(step, p_r, p_n, f_r, f_n):
        exact_transport(step, p_r, p_n, tol_z=1e-11)
        sizes.append(p_r.get_n().size)
        if step >= 0:
            wrapped.append(float(p_r.get_attr("z").max()))
    S.run_pair(ref, b200, setup, 6, on_step=check)
    assert len(set(sizes)) == 1, "periodic walls must not remove anything"
    assert max(wrapped) > 80.0, "nothing re-entered at the top"


@pytest.mark.parametrize("kernel,vt", [(L.kernel_t.geometric, L.vt_t.khvorostyanov_nonspherical),
                                       (L.kernel_t.hall_pinsky_stratocumulus, L.vt_t.beard77fast),
                                       (L.kernel_t.hall_pinsky_1000mb_grav, L.vt_t.beard76),
                                       (L.kernel_t.hall_pinsky_cumulonimbus, L.vt_t.khvorostyanov_nonspherical),
                                       (L.kernel_t.vohl_davis_no_waals, L.vt_t.beard77fast)])
def test_remaining_kernels_and_fall_speeds(ref, b200, kernel, vt):
    """vterm.hpp:170-221 (non-spherical Khvorostyanov), kernel_definitions/{hall_pinsky_*,vohl_davis_no_waals}: collision outcomes
    exact under the replayed stream, radii and heights to a few ulp"""
    seen = {}

    def setup(lib):
        oi, o, f = S.box_3d(lib, nx=5, ny=4, nz=8, sd_conc=32, rain_mode=True, kernel=kernel, vt=vt, dt=2.0, sstp_coal=2)
        o.cond = 0
        return oi, o, f

    def check(step, p_r, p_n, f_r, f_n):
        n_r, n_n = p_r.get_n(), p_n.get_n()
        assert n_r.size == n_n.size and np.array_equal(n_r, n_n), step
        assert np.array_equal(p_r.get_attr("rd3"), p_n.get_attr("rd3")), step
        assert S.rel_err(p_r.get_attr("rw2"), p_n.get_attr("rw2")) < 1e-14, step
        assert S.rel_err(p_r.get_attr("z"), p_n.get_attr("z")) < 1e-11, step
        seen.setdefault("n0", int(n_r.sum()))
        seen["n1"] = int(n_r.sum())
    S.run_pair(ref, b200, setup, 6, on_step=check)
    assert seen["n1"] < seen["n0"], "no collision happened"


def test_variable_time_step(ref, b200):
    """opts.dt with opts_init.variable_dt_switch (impl_adjust_timesteps.ipp:13-22): the step length and the numbers of sub-steps
    follow the per-step value; full microphysics, integer state exact, radii within the condensation tolerance"""
    dts = [1.0, 0.5, 2.0, 1.5, 0.25]

    def drive(lib, backend):
        oi, o, f = S.box_3d(lib, nx=4, ny=4, nz=6, sd_conc=24, rain_mode=True, sstp_cond=2, sstp_coal=2)
        oi.variable_dt_switch = 1
        p = lib.factory(backend, oi)
        p.init(f["th"], f["rv"], f["rhod"], None, f["Cx"], f["Cy"], f["Cz"])
        out = []
        for dt in dts:
            o.dt = dt
            p.step_sync(o, f["th"], f["rv"], f["rhod"], f["Cx"], f["Cy"], f["Cz"])
            p.step_async(o)
            out.append((p.get_n(), p.get_attr("rw2"), p.get_attr("x"), f["th"].copy(), f["rv"].copy()))
        return out
    a, b = drive(ref, L.backend_t.serial), drive(b200, L.backend_t.CUDA)
    for step, ((n_r, rw_r, x_r, th_r, rv_r), (n_n, rw_n, x_n, th_n, rv_n)) in enumerate(zip(a, b)):
        assert np.array_equal(n_r, n_n), step
        assert np.array_equal(x_r, x_n), step
        assert S.rel_err(rw_r, rw_n) < (step + 2) * 2.0 ** -14, step
        assert S.rel_err(th_r, th_n) < 1e-8 and S.rel_err(rv_r, rv_n) < 1e-6, step


def test_variable_time_step_needs_the_switch(b200):
    oi, o, f = S.box_3d(b200, nx=3, ny=3, nz=3, sd_conc=8)
    p = b200.factory(L.backend_t.CUDA, oi)
    p.init(f["th"], f["rv"], f["rhod"], None, f["Cx"], f["Cy"], f["Cz"])
    o.dt = 0.5
    with pytest.raises(RuntimeError, match="variable_dt_switch"):
        p.step_sync(o, f["th"], f["rv"], f["rhod"], f["Cx"], f["Cy"], f["Cz"])
