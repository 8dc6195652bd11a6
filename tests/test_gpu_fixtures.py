"""The B200 back-end against the fixtures the reference's own tests hold for this path, with the reference's own tolerances,
and against golden vectors produced by the reference build (tests/golden/, made by tools/make_golden.py)."""
import os

import numpy as np
import pytest

from libcloudphxx_b200 import lgrngn as L
from tests import support as S

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RH_NAMES = {"pv_cc": 0, "rv_cc": 1, "pv_tet": 2, "rv_tet": 3}


@pytest.mark.parametrize("row", S.load_cond_substepping_rows(), ids=lambda r: "%s-sstp%s-constp%s" % (r["RH_formula"], r["sstp_cond"], r["constp"]))
def test_cond_substepping_fixture(b200, row):
    """all 56 per-cell rows of refdata/lgrngn_cond_substepping_refdata.csv (4 RH formulae x 7 sub-step counts x const_p on/off);
    tolerances of tests/python/physics/lgrngn_cond_substepping_test.py:79-91"""
    res = S.cond_substepping_scenario(b200, L.backend_t.CUDA, RH_NAMES[row["RH_formula"]], int(row["sstp_cond"]), row["constp"] == "True")
    bad = S.check_cond_substepping(res, row)
    assert not bad, bad


@pytest.mark.parametrize("row", S.load_cond_substepping_rows("perparticle"),
                         ids=lambda r: "%s-sstp%s-constp%s-mix%s" % (r["RH_formula"], r["sstp_cond"], r["constp"], r["mixing"]))
def test_cond_perparticle_substepping_fixture(b200, row):
    """the 112 per-particle rows without adaptation (exact_sstp_cond, sstp_cond_mix on and off) of the same reference file"""
    res = S.cond_substepping_scenario(b200, L.backend_t.CUDA, RH_NAMES[row["RH_formula"]], int(row["sstp_cond"]), row["constp"] == "True",
                                      exact_sstp=True, mixing=row["mixing"] == "True")
    bad = S.check_cond_substepping(res, row)
    assert not bad, bad


@pytest.mark.parametrize("row", S.load_cond_substepping_rows("adaptive"),
                         ids=lambda r: "%s-sstp%s-constp%s-act%s" % (r["RH_formula"], r["sstp_cond"], r["constp"], r["sstp_cond_act"]))
def test_cond_adaptive_substepping_fixture(b200, row):
    """the 112 adaptive per-particle rows (sstp_cond_act 1 and 8) of the same reference file"""
    res = S.cond_substepping_scenario(b200, L.backend_t.CUDA, RH_NAMES[row["RH_formula"]], int(row["sstp_cond"]), row["constp"] == "True",
                                      exact_sstp=True, mixing=False, adaptive=True, sstp_cond_act=int(row["sstp_cond_act"]),
                                      drw2_eps=float(row["sstp_cond_adapt_drw2_eps"]), drw2_max=float(row["sstp_cond_adapt_drw2_max"]))
    bad = S.check_cond_substepping(res, row)
    assert not bad, bad


@pytest.mark.parametrize("vt", [L.vt_t.beard76, L.vt_t.beard77, L.vt_t.beard77fast])
def test_hall_davis_coalescence_vs_bott(b200, vt):
    """tests/python/physics/coalescence_hall_davis_no_waals.py:82-105: mass-density spectrum after 1800 s vs Bott's bin model"""
    assert bott_rmsd(b200, vt) < 6e-2


def bott_rmsd(b200, vt):
    bott = np.load(os.path.join(ROOT, "tests", "golden", "bott1800.npy"))
    oi, o, f = S.hall_davis_box(b200, vt)
    p = b200.factory(L.backend_t.CUDA, oi)
    p.init(f["th"], f["rv"], f["rhod"])
    p.step_sync(o, f["th"], f["rv"], f["rhod"])
    p.step_async(o)
    return S.rmsd(S.mass_density_spectrum(p) * 1000, bott)


def test_golden_golovin_box(b200):
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_golovin_box.npz"))
    oi, o, f = S.box_golovin(b200, n_sd=2 ** 10)
    p = b200.factory(L.backend_t.CUDA, oi)
    p.init(f["th"], f["rv"], f["rhod"])
    assert np.array_equal(g["g_n_0"], p.get_n()) and np.array_equal(g["g_rw2_0"], p.get_attr("rw2"))
    for step in range(1, 13):
        p.step_sync(o, f["th"], f["rv"], f["rhod"]); p.step_async(o)
        assert np.array_equal(g["g_n_%d" % step], p.get_n()), step
        assert np.array_equal(g["g_rd3_%d" % step], p.get_attr("rd3")), step
        assert S.rel_err(g["g_rw2_%d" % step], p.get_attr("rw2")) < 1e-14, step


def test_golden_box3d(b200):
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_box3d.npz"))
    oi, o, f = S.box_3d(b200, nx=4, ny=3, nz=4, sd_conc=8, rain_mode=True)
    p = b200.factory(L.backend_t.CUDA, oi)
    p.init(f["th"], f["rv"], f["rhod"], None, f["Cx"], f["Cy"], f["Cz"])
    for k in ("rd3", "rw2", "x", "y", "z"):
        assert np.array_equal(g["b_%s_0" % k], p.get_attr(k)), k
    for step in range(1, 4):
        p.step_sync(o, f["th"], f["rv"], f["rhod"], f["Cx"], f["Cy"], f["Cz"]); p.step_async(o)
        assert np.array_equal(g["b_n_%d" % step], p.get_n()), step
        assert np.array_equal(g["b_rd3_%d" % step], p.get_attr("rd3")), step
        for k in ("x", "y"):
            assert np.array_equal(g["b_%s_%d" % (k, step)], p.get_attr(k)), (k, step)
        assert S.rel_err(g["b_rw2_%d" % step], p.get_attr("rw2")) < (step + 1) * 2.0 ** -15, step
        assert S.rel_err(g["b_th_%d" % step], f["th"]) < 1e-9 and S.rel_err(g["b_rv_%d" % step], f["rv"]) < 1e-7, step


def test_golovin_analytic(b200):
    """tests/python/physics/coalescence_golovin.py:112-155 (sd_conc branch): RMSD of the mass density vs Golovin's solution < 1.2e-5"""
    assert golovin_analytic_rmsd(b200) < 1.2e-5


def golovin_analytic_rmsd(b200):
    from scipy import special
    n_zero, r_zero, b, t_sim = 2.0 ** 23, 30.084e-6, 1500., 800
    oi, o, f = S.box_golovin(b200, n_sd=2 ** 14, dt=float(t_sim), sstp_coal=t_sim)
    oi.terminal_velocity = L.vt_t.beard77
    p = b200.factory(L.backend_t.CUDA, oi)
    p.init(f["th"], f["rv"], f["rhod"])
    p.diag_all(); p.diag_wet_mom(0)
    n0 = p.outbuf()[0]
    p.step_sync(o, f["th"], f["rv"], f["rhod"]); p.step_async(o)
    bins = 10 ** (-6 + np.arange(150) / 50.)
    res = S.mass_density_spectrum(p, scale=1.0)
    vol = lambda r: 4. / 3. * r ** 3 * np.pi
    v0 = vol(r_zero)
    gol = np.zeros_like(res)
    for i in range(res.size):
        v = vol((bins[i] + bins[i + 1]) / 2.)
        x, T = v / v0, b * n0 * v0 * t_sim
        tau = 1 - np.exp(-T)
        bessel = special.iv(1, 2 * x * np.sqrt(tau))
        val = 0. if np.isinf(bessel) else n0 / v0 * bessel * (1 - tau) * np.exp(-x * (tau + 1)) / x / np.sqrt(tau)
        gol[i] = (0. if np.isnan(val) else val) * v * v * 3000.
    return S.rmsd(res, gol)
