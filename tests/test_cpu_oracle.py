"""Pins the oracle (no GPU needed).

  1. oracle/_ref - the reference's own serial/OpenMP back-ends built from the unmodified sources - reproduces the fixtures the
     reference's tests hold for this path: refdata/lgrngn_cond_substepping_refdata.csv (per-cell rows, tolerances of
     lgrngn_cond_substepping_test.py:79-91), the Bott array of coalescence_hall_davis_no_waals.py:82 (RMSD < 6e-2),
     the known-answer p_vs(273.16 K) = 611.73 Pa (tests/common/test_common_pvs.cpp:7), mass conservation of test_coal.py.
  2. oracle/sdm_port.py - the numpy restatement - agrees with oracle/_ref bit for bit on the same seeded inputs, and with
     the committed golden vectors made from it (tools/make_golden.py), so it stays pinned where the reference is absent.
"""
import importlib.util
import math
import os

import numpy as np
import pytest

from libcloudphxx_b200 import lgrngn as L
from tests import support as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_port():
    spec = importlib.util.spec_from_file_location("sdm_port", os.path.join(ROOT, "oracle", "sdm_port.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


port = load_port()


def test_known_answer_saturation_pressure():
    assert port.p_vs_cc(273.16) == 611.73


def test_host_random_stream_is_libstdcxx_mt19937():
    """10000th draw of std::mt19937 (default seed 5489) is 4123659995 by the C++ standard [rand.predef]"""
    rng = port.HostRNG(5489)
    assert int(rng.raw(10000)[-1]) == 4123659995


@pytest.mark.parametrize("sstp_cond", [1, 2, 3, 4, 6, 8, 32])
@pytest.mark.parametrize("rhf", [L.RH_formula_t.pv_cc, L.RH_formula_t.rv_tet])
def test_reference_build_reproduces_cond_fixture(ref, sstp_cond, rhf):
    names = {0: "pv_cc", 1: "rv_cc", 2: "pv_tet", 3: "rv_tet"}
    rows = [r for r in S.load_cond_substepping_rows() if r["constp"] == "False" and r["RH_formula"] == names[rhf] and int(r["sstp_cond"]) == sstp_cond]
    assert len(rows) == 1
    res = S.cond_substepping_scenario(ref, L.backend_t.OpenMP, rhf, sstp_cond, False)
    assert not S.check_cond_substepping(res, rows[0]), S.check_cond_substepping(res, rows[0])


def test_reference_build_reproduces_cond_fixture_const_p(ref):
    rows = [r for r in S.load_cond_substepping_rows() if r["constp"] == "True" and r["RH_formula"] == "pv_cc" and int(r["sstp_cond"]) == 4]
    res = S.cond_substepping_scenario(ref, L.backend_t.OpenMP, L.RH_formula_t.pv_cc, 4, True)
    assert not S.check_cond_substepping(res, rows[0]), S.check_cond_substepping(res, rows[0])


def test_th_diff_of_the_sstp32_rows_depends_on_the_build_flags(ref):
    """why tests/support.py allows 2e-5 on th_diff for the sstp_cond = 32 rows: the reference's fixture was produced with its -Ofast
    release flags; the same sources built that way (oracle/_ref/liblgrngn_ref_fast.so) reproduce the row to 1e-6, the IEEE-strict
    -O2 build used for parity lands 1.0e-5 .. 1.5e-5 away - a property of the build flags, not of the restated algorithm"""
    fast_path = os.path.join(ROOT, "oracle", "_ref", "liblgrngn_ref_fast.so")
    if not os.path.exists(fast_path):
        pytest.skip("the -Ofast build of the reference is not present")
    fast = L.Library(fast_path)
    row = [r for r in S.load_cond_substepping_rows() if r["constp"] == "False" and r["RH_formula"] == "pv_cc" and int(r["sstp_cond"]) == 32][0]
    want = float(row["th_diff"])
    got_fast = S.cond_substepping_scenario(fast, L.backend_t.OpenMP, L.RH_formula_t.pv_cc, 32, False)["th_diff"]
    got_strict = S.cond_substepping_scenario(ref, L.backend_t.OpenMP, L.RH_formula_t.pv_cc, 32, False)["th_diff"]
    assert abs(got_fast - want) < 1e-6, (got_fast, want)
    assert 5e-6 < abs(got_strict - want) < S.TH_DIFF_ATOL_SSTP32, (got_strict, want)
    # and a row with fewer sub-steps holds the reference's own 1e-5 on the strict build with room to spare
    row8 = [r for r in S.load_cond_substepping_rows() if r["constp"] == "False" and r["RH_formula"] == "pv_cc" and int(r["sstp_cond"]) == 8][0]
    got8 = S.cond_substepping_scenario(ref, L.backend_t.OpenMP, L.RH_formula_t.pv_cc, 8, False)["th_diff"]
    assert abs(got8 - float(row8["th_diff"])) < 5e-6


def test_reference_build_reproduces_bott_spectrum(ref):
    bott = np.load(os.path.join(ROOT, "tests", "golden", "bott1800.npy"))
    oi, o, f = S.hall_davis_box(ref, L.vt_t.beard77fast)
    p = ref.factory(L.backend_t.OpenMP, oi)
    p.init(f["th"], f["rv"], f["rhod"])
    p.step_sync(o, f["th"], f["rv"], f["rhod"])
    p.step_async(o)
    assert S.rmsd(S.mass_density_spectrum(p) * 1000, bott) < 6e-2


def make_port(case, **kw):
    if case == "golovin":
        n_sd = kw["n_sd"]
        r0, n0 = 30.084e-6, 2.0 ** 23
        fun = lambda lnr: n0 * 3. * (math.exp(lnr) / r0) ** 3 * math.exp(-(math.exp(lnr) / r0) ** 3)
        return port.Particles(dt=1., sd_conc=n_sd, n_sd_max=n_sd, kernel="golovin", kernel_params={"b": 1500.}, dry_distros=[(1e-10, fun)])
    raise ValueError(case)


def expvol_as_capi(lnr, r0=30.084e-6, n0=2.0 ** 23):
    # same arithmetic as the binding's functor (bindings/lgrngn_capi.cpp: expvolume::funval)
    r = math.exp(lnr)
    q = r / r0
    q3 = q * q * q
    return n0 * 3. * q3 * math.exp(-q3)


def lognormal_as_capi(modes):
    def f(lnr):
        res = 0.0
        for mean_r, stdev, n_tot in modes:
            lns, d = math.log(stdev), lnr - math.log(mean_r)
            res += n_tot * math.exp(-(d * d) / 2. / (lns * lns)) / lns / math.sqrt(2 * math.pi)
        return res
    return f


def test_port_matches_reference_golovin_box_exactly(ref):
    n_sd = 512
    oi, o, f = S.box_golovin(ref, n_sd=n_sd)
    p_r = ref.factory(L.backend_t.serial, oi)
    p_r.init(f["th"], f["rv"], f["rhod"])
    p_p = port.Particles(dt=1., sd_conc=n_sd, n_sd_max=n_sd, kernel="golovin", kernel_params={"b": 1500.}, dry_distros=[(1e-10, expvol_as_capi)],
                         sedi_switch=False)
    p_p.init(f["th"], f["rv"], f["rhod"])
    assert np.array_equal(p_r.get_n(), p_p.n) and np.array_equal(p_r.get_attr("rd3"), p_p.rd3) and np.array_equal(p_r.get_attr("rw2"), p_p.rw2)
    collided = 0
    for step in range(25):
        p_r.step_sync(o, f["th"], f["rv"], f["rhod"]); p_r.step_async(o)
        p_p.step_sync(f["th"], f["rv"], f["rhod"], cond=False)
        collided += p_p.step_async(adve=False, sedi=False, coal=True, cond=False)
        assert np.array_equal(p_r.get_n(), p_p.n), step
        assert np.array_equal(p_r.get_attr("rd3"), p_p.rd3), step
        assert np.array_equal(p_r.get_attr("rw2"), p_p.rw2), step
    assert collided > 0


def test_port_matches_golden_golovin_box():
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_golovin_box.npz"))
    n_sd = 2 ** 10
    p = port.Particles(dt=1., sd_conc=n_sd, n_sd_max=n_sd, kernel="golovin", kernel_params={"b": 1500.}, dry_distros=[(1e-10, expvol_as_capi)])
    th, rv, rhod = np.array([300.]), np.array([0.01]), np.array([1.])
    p.init(th, rv, rhod)
    assert np.array_equal(g["g_n_0"], p.n) and np.array_equal(g["g_rw2_0"], p.rw2)
    for step in range(1, 13):
        p.step_sync(th, rv, rhod, cond=False)
        p.step_async(adve=False, sedi=False, coal=True, cond=False)
        assert np.array_equal(g["g_n_%d" % step], p.n), step
        assert np.array_equal(g["g_rd3_%d" % step], p.rd3) and np.array_equal(g["g_rw2_%d" % step], p.rw2), step


def port_box3d(f, nx, ny, nz, sd_conc, eff):
    return port.Particles(nx=nx, ny=ny, nz=nz, dx=20., dy=20., dz=20., dt=1., x1=nx * 20., y1=ny * 20., z1=nz * 20., sd_conc=sd_conc,
                          n_sd_max=int(nx * ny * nz * sd_conc * 1.5), kernel="efficiencies", kernel_params={"eff": eff[1:], "r_max": eff[0]},
                          dry_distros=[(0.61, lognormal_as_capi(S.AEROSOL_ICICLE)), (1.28, lognormal_as_capi([(30e-6, 1.2, 1e5)]))])


def test_port_matches_reference_full_step_3d(ref):
    """cond + coal + sedi + adve in a 3-D box: the restatement follows the reference to the last bit"""
    nx, ny, nz, sd_conc = 3, 2, 4, 8
    eff = np.fromfile(os.path.join(ROOT, "libcloudphxx_b200", "data", "hall_davis_no_waals.f64"))
    oi, o, f = S.box_3d(ref, nx=nx, ny=ny, nz=nz, sd_conc=sd_conc, rain_mode=True)
    p_r = ref.factory(L.backend_t.serial, oi)
    p_r.init(f["th"], f["rv"], f["rhod"], None, f["Cx"], f["Cy"], f["Cz"])
    fp = {k: v.copy() for k, v in f.items()}
    p_p = port_box3d(fp, nx, ny, nz, sd_conc, eff)
    p_p.init(fp["th"], fp["rv"], fp["rhod"], fp["Cx"], fp["Cy"], fp["Cz"])
    for k, a in (("rd3", p_p.rd3), ("rw2", p_p.rw2), ("x", p_p.x), ("y", p_p.y), ("z", p_p.z)):
        assert np.array_equal(p_r.get_attr(k), a), k
    assert np.array_equal(p_r.get_n(), p_p.n)
    for step in range(3):
        p_r.step_sync(o, f["th"], f["rv"], f["rhod"], f["Cx"], f["Cy"], f["Cz"]); p_r.step_async(o)
        th, rv = p_p.step_sync(fp["th"], fp["rv"], fp["rhod"])
        fp["th"][:], fp["rv"][:] = th.reshape(fp["th"].shape), rv.reshape(fp["rv"].shape)
        p_p.step_async()
        assert np.array_equal(p_r.get_n(), p_p.n), step
        for k, a in (("rd3", p_p.rd3), ("rw2", p_p.rw2), ("x", p_p.x), ("y", p_p.y), ("z", p_p.z)):
            assert np.array_equal(p_r.get_attr(k), a), (k, step, S.rel_err(p_r.get_attr(k), a))
        assert np.array_equal(f["th"], fp["th"]) and np.array_equal(f["rv"], fp["rv"]), step


@pytest.mark.parametrize("scheme", ["implicit", "euler", "pred_corr"])
def test_port_advection_schemes_match_reference_3d(ref, scheme):
    """all three advection schemes on a sheared, non-uniform Courant field (pred_corr reads the 2-column Courant halo, filled
    the way init_e2l wraps it): positions bit-identical to the reference after every step"""
    nx, ny, nz, sd_conc = 6, 5, 7, 6
    eff = np.fromfile(os.path.join(ROOT, "libcloudphxx_b200", "data", "hall_davis_no_waals.f64"))
    oi, o, f = S.box_3d(ref, nx=nx, ny=ny, nz=nz, sd_conc=sd_conc, adve=getattr(L.as_t, scheme))
    rng = np.random.default_rng(7)
    f["Cx"] = 0.3 + 0.4 * rng.random(f["Cx"].shape)
    f["Cy"] = -0.2 + 0.4 * rng.random(f["Cy"].shape)
    f["Cz"] = -0.1 + 0.2 * rng.random(f["Cz"].shape)
    f["Cz"][:, :, 0] = 0.0
    f["Cz"][:, :, -1] = 0.0
    o.cond = o.coal = o.sedi = 0
    p_r = ref.factory(L.backend_t.serial, oi)
    p_r.init(f["th"], f["rv"], f["rhod"], None, f["Cx"], f["Cy"], f["Cz"])
    p_p = port.Particles(nx=nx, ny=ny, nz=nz, dx=20., dy=20., dz=20., dt=1., x1=nx * 20., y1=ny * 20., z1=nz * 20., sd_conc=sd_conc,
                         n_sd_max=int(nx * ny * nz * sd_conc * 1.5), kernel="efficiencies", kernel_params={"eff": eff[1:], "r_max": eff[0]},
                         dry_distros=[(0.61, lognormal_as_capi(S.AEROSOL_ICICLE))], adve_scheme=scheme)
    p_p.init(f["th"], f["rv"], f["rhod"], f["Cx"], f["Cy"], f["Cz"])
    moved = False
    x0 = p_p.x.copy()
    for step in range(8):
        p_r.step_sync(o, f["th"], f["rv"], f["rhod"], f["Cx"], f["Cy"], f["Cz"]); p_r.step_async(o)
        p_p.step_sync(f["th"], f["rv"], f["rhod"], cond=False)
        p_p.step_async(adve=True, sedi=False, coal=False, cond=False)
        for k, a in (("x", p_p.x), ("y", p_p.y), ("z", p_p.z)):
            b = p_r.get_attr(k)
            assert a.size == b.size, (k, step, a.size, b.size)
            assert np.array_equal(b, a), (k, step, S.rel_err(b, a))
        moved |= p_p.x.size != x0.size or not np.array_equal(p_p.x, x0)
    assert moved


@pytest.mark.parametrize("rhf", ["pv_cc", "rv_cc", "pv_tet", "rv_tet"])
@pytest.mark.parametrize("sstp_cond", [1, 3])
def test_port_parcel_condensation_matches_reference(ref, rhf, sstp_cond):
    """0-D parcel, condensation only, each of the four RH formulae (hskpng_Tpr.ipp:71-103), with and without per-cell
    sub-stepping: wet radii, th and rv bit-identical to the reference after every step"""
    oi, o, f = S.parcel(ref, n_sd=96, dt=1.0, sstp_cond=sstp_cond, RH_formula=getattr(L.RH_formula_t, rhf))
    p_r = ref.factory(L.backend_t.serial, oi)
    p_r.init(f["th"], f["rv"], f["rhod"])
    fp = {k: v.copy() for k, v in f.items()}
    p_p = port.Particles(dt=1., sd_conc=96, n_sd_max=96, sstp_cond=sstp_cond, dry_distros=[(0.61, lognormal_as_capi([(0.04e-6, 2.0, 566e6)]))],
                         sedi_switch=False, coal_switch=False, RH_formula=rhf, vt="undefined")
    p_p.init(fp["th"], fp["rv"], fp["rhod"])
    assert np.array_equal(p_r.get_attr("rw2"), p_p.rw2)
    for step in range(12):
        f["rhod"] *= 0.999
        fp["rhod"] *= 0.999
        p_r.step_sync(o, f["th"], f["rv"], f["rhod"]); p_r.step_async(o)
        th, rv = p_p.step_sync(fp["th"], fp["rv"], fp["rhod"])
        fp["th"][:], fp["rv"][:] = th, rv
        p_p.step_async(adve=False, sedi=False, coal=False, cond=True)
        assert np.array_equal(p_r.get_attr("rw2"), p_p.rw2), (step, S.rel_err(p_r.get_attr("rw2"), p_p.rw2))
        assert np.array_equal(f["th"], fp["th"]) and np.array_equal(f["rv"], fp["rv"]), step
    assert p_p.rw2.max() > 1e-12, "nothing grew"


@pytest.mark.parametrize("vt", ["beard76", "beard77", "beard77fast", "khvorostyanov_spherical", "khvorostyanov_nonspherical"])
def test_port_fall_speed_formulae_match_reference(ref, vt):
    """sedimentation with each terminal-velocity formula (common/vterm.hpp:38-221), droplets from haze to millimetre drops so
    that every branch of the piecewise fits is taken: z - dt * vt bit-identical to the reference, step after step"""
    nx, ny, nz, sd_conc = 3, 2, 8, 12
    eff = np.fromfile(os.path.join(ROOT, "libcloudphxx_b200", "data", "hall_davis_no_waals.f64"))
    oi, o, f = S.box_3d(ref, nx=nx, ny=ny, nz=nz, sd_conc=sd_conc, vt=getattr(L.vt_t, vt), cx=0.0, cy=0.0)
    big = [(30e-6, 1.3, 1e5), (200e-6, 1.25, 1e2)]
    oi.dry_distros = [L.lognormal(0.61, S.AEROSOL_ICICLE), L.lognormal(1.28, big)]
    o.cond = o.coal = 0
    p_r = ref.factory(L.backend_t.serial, oi)
    p_r.init(f["th"], f["rv"], f["rhod"], None, f["Cx"], f["Cy"], f["Cz"])
    p_p = port.Particles(nx=nx, ny=ny, nz=nz, dx=20., dy=20., dz=20., dt=1., x1=nx * 20., y1=ny * 20., z1=nz * 20., sd_conc=sd_conc,
                         n_sd_max=int(nx * ny * nz * sd_conc * 1.5), kernel="efficiencies", kernel_params={"eff": eff[1:], "r_max": eff[0]},
                         dry_distros=[(0.61, lognormal_as_capi(S.AEROSOL_ICICLE)), (1.28, lognormal_as_capi(big))], vt=vt)
    p_p.init(f["th"], f["rv"], f["rhod"], f["Cx"], f["Cy"], f["Cz"])
    r = np.sqrt(p_p.rw2)
    assert r.min() < 9.5e-6 and ((r > 20e-6) & (r < 5.035e-4)).any() and r.max() > 5.035e-4, (r.min(), r.max())
    z0 = p_p.z.copy()
    for step in range(4):
        p_r.step_sync(o, f["th"], f["rv"], f["rhod"], f["Cx"], f["Cy"], f["Cz"]); p_r.step_async(o)
        p_p.step_sync(f["th"], f["rv"], f["rhod"], cond=False)
        p_p.step_async(adve=True, sedi=True, coal=False, cond=False)
        a, b = p_r.get_attr("z"), p_p.z
        assert a.size == b.size, (step, a.size, b.size)
        assert np.array_equal(a, b), (step, S.rel_err(a, b))
    assert p_p.z.size < z0.size, "no drop reached the ground"


@pytest.mark.parametrize("kernel", ["geometric", "geometric_mult", "long", "golovin"])
def test_port_collision_kernels_match_reference_3d(ref, kernel):
    """coalescence + sedimentation + advection with the analytic collision kernels (kernels.hpp:40-176): multiplicities and radii
    bit-identical to the reference, collision by collision (the tabulated-efficiency kernels are covered by the tests above)"""
    nx, ny, nz, sd_conc = 3, 2, 4, 16
    oi, o, f = S.box_3d(ref, nx=nx, ny=ny, nz=nz, sd_conc=sd_conc, rain_mode=True, dt=4.0)
    kparams = {}
    if kernel == "geometric":
        oi.kernel = L.kernel_t.geometric
    elif kernel == "geometric_mult":
        oi.kernel, oi.kernel_parameters, kparams = L.kernel_t.geometric, [3.5], {"mult": 3.5}
    elif kernel == "long":
        oi.kernel = L.kernel_t.Long
    else:
        oi.kernel, oi.kernel_parameters, kparams = L.kernel_t.golovin, [1500.], {"b": 1500.}
    o.cond = 0
    p_r = ref.factory(L.backend_t.serial, oi)
    p_r.init(f["th"], f["rv"], f["rhod"], None, f["Cx"], f["Cy"], f["Cz"])
    p_p = port.Particles(nx=nx, ny=ny, nz=nz, dx=20., dy=20., dz=20., dt=4., x1=nx * 20., y1=ny * 20., z1=nz * 20., sd_conc=sd_conc,
                         n_sd_max=int(nx * ny * nz * sd_conc * 1.5), kernel=kernel.split("_")[0], kernel_params=kparams,
                         dry_distros=[(0.61, lognormal_as_capi(S.AEROSOL_ICICLE)), (1.28, lognormal_as_capi([(30e-6, 1.2, 1e5)]))])
    p_p.init(f["th"], f["rv"], f["rhod"], f["Cx"], f["Cy"], f["Cz"])
    collided = 0
    for step in range(6):
        p_r.step_sync(o, f["th"], f["rv"], f["rhod"], f["Cx"], f["Cy"], f["Cz"]); p_r.step_async(o)
        p_p.step_sync(f["th"], f["rv"], f["rhod"], cond=False)
        collided += p_p.step_async(cond=False)
        assert np.array_equal(p_r.get_n(), p_p.n), step
        for k, a in (("rd3", p_p.rd3), ("rw2", p_p.rw2), ("x", p_p.x), ("z", p_p.z)):
            assert np.array_equal(p_r.get_attr(k), a), (k, step, S.rel_err(p_r.get_attr(k), a))
    assert collided > 0, "no collision happened - vacuous"


def test_port_diagnostics_match_reference(ref):
    """selectors, moments, SD concentration, precipitation flux, largest radius (particles_diag.ipp:148-656, moms.ipp:50-387)
    after a few full steps: the restatement sums in the reference's order, so the per-cell arrays are bit-identical"""
    nx, ny, nz, sd_conc = 3, 2, 5, 12
    eff = np.fromfile(os.path.join(ROOT, "libcloudphxx_b200", "data", "hall_davis_no_waals.f64"))
    oi, o, f = S.box_3d(ref, nx=nx, ny=ny, nz=nz, sd_conc=sd_conc, rain_mode=True)
    f["rv"][:, :, nz // 2:] = 1.25e-2                    # supersaturated at the temperature of this shallow column: droplets activate
    p_r = ref.factory(L.backend_t.serial, oi)
    p_r.init(f["th"], f["rv"], f["rhod"], None, f["Cx"], f["Cy"], f["Cz"])
    fp = {k: v.copy() for k, v in f.items()}
    p_p = port_box3d(fp, nx, ny, nz, sd_conc, eff)
    p_p.init(fp["th"], fp["rv"], fp["rhod"], fp["Cx"], fp["Cy"], fp["Cz"])
    for step in range(3):
        p_r.step_sync(o, f["th"], f["rv"], f["rhod"], f["Cx"], f["Cy"], f["Cz"]); p_r.step_async(o)
        th, rv = p_p.step_sync(fp["th"], fp["rv"], fp["rhod"])
        fp["th"][:], fp["rv"][:] = th.reshape(fp["th"].shape), rv.reshape(fp["rv"].shape)
        p_p.step_async()
        assert np.array_equal(p_r.get_attr("rw2"), p_p.rw2), step

    def same(name, a):
        b = p_r.outbuf()
        assert np.array_equal(a.reshape(-1), b), (name, S.rel_err(b, a.reshape(-1)))
        return b
    seen = 0.
    p_r.diag_all(); p_p.diag_all()
    for k in range(4):
        p_r.diag_wet_mom(k); seen += same("wet_mom %d" % k, p_p.diag_wet_mom(k)).sum()
        p_r.diag_dry_mom(k); same("dry_mom %d" % k, p_p.diag_dry_mom(k))
    p_r.diag_kappa_mom(1); same("kappa_mom", p_p.diag_kappa_mom(1))
    p_r.diag_sd_conc(); assert same("sd_conc", p_p.diag_sd_conc()).sum() == p_p.n_part
    p_r.diag_precip_rate(); assert same("precip_rate", p_p.diag_precip_rate()).sum() > 0
    p_r.diag_max_rw(); same("max_rw", p_p.diag_max_rw())
    p_r.diag_wet_rng(1e-6, 25e-6); p_p.diag_wet_rng(1e-6, 25e-6)
    p_r.diag_wet_mom(3); same("cloud water", p_p.diag_wet_mom(3))
    p_r.diag_sd_conc(); part = same("sd_conc of a range", p_p.diag_sd_conc()).sum()
    assert 0 < part < p_p.n_part
    p_r.diag_dry_rng_cons(0.05e-6, 1.); p_p.diag_dry_rng(0.05e-6, 1., cons=True)
    p_r.diag_wet_mom(0); assert 0 < same("consecutive selection", p_p.diag_wet_mom(0)).sum()
    p_r.diag_kappa_rng(1.0, 2.0); p_p.diag_kappa_rng(1.0, 2.0)
    p_r.diag_dry_mom(3); assert same("second aerosol type", p_p.diag_dry_mom(3)).sum() > 0
    p_r.diag_rw_ge_rc(); p_p.diag_rw_ge_rc()
    p_r.diag_wet_mom(0); act = same("activated", p_p.diag_wet_mom(0)).sum()
    p_r.diag_RH_ge_Sc(); p_p.diag_RH_ge_Sc()
    p_r.diag_wet_mom(0); same("RH above critical", p_p.diag_wet_mom(0))
    assert act > 0 and seen > 0


@pytest.mark.parametrize("scheme", ["implicit", "pred_corr"])
def test_port_subsidence_and_open_side_walls_match_reference(ref, scheme):
    """large-scale subsidence (subs.ipp:13-25: w_LS at the level the SD was last indexed in) and open side walls (bcnd.ipp:126-142,
    202-216: SDs leaving through x or y are removed): surviving SDs and their positions bit-identical to the reference"""
    nx, ny, nz, sd_conc = 4, 3, 6, 8
    eff = np.fromfile(os.path.join(ROOT, "libcloudphxx_b200", "data", "hall_davis_no_waals.f64"))
    oi, o, f = S.box_3d(ref, nx=nx, ny=ny, nz=nz, sd_conc=sd_conc, adve=getattr(L.as_t, scheme), cx=0.4, cy=-0.3)
    w_LS = 0.5 + 0.3 * np.arange(nz)
    oi.subs_switch, oi.w_LS, oi.open_side_walls = 1, list(w_LS), 1
    o.cond = o.coal = 0
    o.subs = 1
    p_r = ref.factory(L.backend_t.serial, oi)
    p_r.init(f["th"], f["rv"], f["rhod"], None, f["Cx"], f["Cy"], f["Cz"])
    p_p = port.Particles(nx=nx, ny=ny, nz=nz, dx=20., dy=20., dz=20., dt=1., x1=nx * 20., y1=ny * 20., z1=nz * 20., sd_conc=sd_conc,
                         n_sd_max=int(nx * ny * nz * sd_conc * 1.5), kernel="efficiencies", kernel_params={"eff": eff[1:], "r_max": eff[0]},
                         dry_distros=[(0.61, lognormal_as_capi(S.AEROSOL_ICICLE))], adve_scheme=scheme, open_side_walls=True, w_LS=w_LS)
    p_p.init(f["th"], f["rv"], f["rhod"], f["Cx"], f["Cy"], f["Cz"])
    n0 = p_p.n_part
    for step in range(6):
        p_r.step_sync(o, f["th"], f["rv"], f["rhod"], f["Cx"], f["Cy"], f["Cz"]); p_r.step_async(o)
        p_p.step_sync(f["th"], f["rv"], f["rhod"], cond=False)
        p_p.step_async(adve=True, sedi=True, coal=False, cond=False, subs=True)
        for k, a in (("x", p_p.x), ("y", p_p.y), ("z", p_p.z)):
            b = p_r.get_attr(k)
            assert a.size == b.size, (k, step, a.size, b.size)
            assert np.array_equal(b, a), (k, step, S.rel_err(b, a))
    assert 0 < p_p.n_part < n0, "nothing left the domain"


@pytest.mark.parametrize("sstp_cond", [1, 4])
def test_port_const_p_standard_theta_matches_reference(ref, sstp_cond):
    """opts_init.const_p with the 'standard' potential temperature (hskpng_Tpr.ipp:231-260, theta_std.hpp:36-41): prescribed pressure,
    T = th * exner(p); parcel condensation bit-identical to the reference"""
    oi, o, f = S.parcel(ref, n_sd=80, dt=1.0, sstp_cond=sstp_cond)
    R_d, R_v, c_pd = 8.3144621 / 0.02897, 8.3144621 / 0.018, 1005.0
    T0 = (f["th"][0] * (f["rhod"][0] * R_d / 1e5) ** (R_d / c_pd)) ** (c_pd / (c_pd - R_d))
    p_prof = np.array([f["rhod"][0] * (R_d + f["rv"][0] * R_v) * T0])
    f["th"][0] = S.th_dry2std(f["th"][0], f["rv"][0])
    oi.const_p, oi.th_dry = 1, 0
    p_r = ref.factory(L.backend_t.serial, oi)
    p_r.init(f["th"], f["rv"], f["rhod"], p_prof)
    fp = {k: v.copy() for k, v in f.items()}
    p_p = port.Particles(dt=1., sd_conc=80, n_sd_max=80, sstp_cond=sstp_cond, dry_distros=[(0.61, lognormal_as_capi([(0.04e-6, 2.0, 566e6)]))],
                         sedi_switch=False, coal_switch=False, vt="undefined", th_dry=False, const_p=True)
    p_p.init(fp["th"], fp["rv"], fp["rhod"], p=p_prof)
    assert np.array_equal(p_r.get_attr("rw2"), p_p.rw2)
    for step in range(10):
        f["th"] -= 0.02                                   # cooling at constant pressure drives the condensation
        fp["th"] -= 0.02
        p_r.step_sync(o, f["th"], f["rv"], f["rhod"]); p_r.step_async(o)
        th, rv = p_p.step_sync(fp["th"], fp["rv"], fp["rhod"])
        fp["th"][:], fp["rv"][:] = th, rv
        p_p.step_async(adve=False, sedi=False, coal=False, cond=True)
        assert np.array_equal(p_r.get_attr("rw2"), p_p.rw2), (step, S.rel_err(p_r.get_attr("rw2"), p_p.rw2))
        assert np.array_equal(f["th"], fp["th"]) and np.array_equal(f["rv"], fp["rv"]), step
    assert p_p.rw2.max() > 1e-12, "nothing grew"


def test_port_kinematic_2d_matches_reference(ref):
    """cfg3-shaped case (BASELINE configs[2]): 2-D single-eddy flow with partial boundary cells (x0 = dx/2 ...), geometric kernel
    with the icicle multiplier 0.5, Khvorostyanov fall speeds, cond + coal sub-stepping: bit-identical to the reference"""
    nx, nz, sd_conc = 7, 6, 10
    oi, o, f = S.kinematic_2d(ref, nx=nx, nz=nz, sd_conc=sd_conc, w_max=6.0)
    f["rv"][:, nz // 2:] = 1.2e-2                        # supersaturated aloft so that droplets activate and collide
    p_r = ref.factory(L.backend_t.serial, oi)
    p_r.init(f["th"], f["rv"], f["rhod"], None, f["Cx"], None, f["Cz"])
    fp = {k: v.copy() for k, v in f.items()}
    p_p = port.Particles(nx=nx, nz=nz, dx=oi.dx, dz=oi.dz, dt=1., x0=oi.x0, z0=oi.z0, x1=oi.x1, z1=oi.z1, sd_conc=sd_conc,
                         n_sd_max=nx * nz * sd_conc, sstp_cond=2, sstp_coal=2, kernel="geometric", kernel_params={"mult": 0.5},
                         vt="khvorostyanov_spherical", dry_distros=[(0.61, lognormal_as_capi(S.AEROSOL_ICICLE))])
    p_p.init(fp["th"], fp["rv"], fp["rhod"], fp["Cx"], None, fp["Cz"])
    assert np.array_equal(p_r.get_n(), p_p.n)
    for k, a in (("rd3", p_p.rd3), ("rw2", p_p.rw2), ("x", p_p.x), ("z", p_p.z)):
        assert np.array_equal(p_r.get_attr(k), a), k
    for step in range(4):
        p_r.step_sync(o, f["th"], f["rv"], f["rhod"], f["Cx"], None, f["Cz"]); p_r.step_async(o)
        th, rv = p_p.step_sync(fp["th"], fp["rv"], fp["rhod"])
        fp["th"][:], fp["rv"][:] = th.reshape(fp["th"].shape), rv.reshape(fp["rv"].shape)
        p_p.step_async()
        assert np.array_equal(p_r.get_n(), p_p.n), step
        for k, a in (("rd3", p_p.rd3), ("rw2", p_p.rw2), ("x", p_p.x), ("z", p_p.z)):
            assert np.array_equal(p_r.get_attr(k), a), (k, step, S.rel_err(p_r.get_attr(k), a))
        assert np.array_equal(f["th"], fp["th"]) and np.array_equal(f["rv"], fp["rv"]), step


def test_port_recycling_matches_reference(ref):
    """opts.rcyc (rcyc.ipp:44-139): who is split, who is re-created, and the storage order afterwards"""
    nx, ny, nz, sd_conc = 4, 3, 6, 16
    eff = np.fromfile(os.path.join(ROOT, "libcloudphxx_b200", "data", "hall_davis_no_waals.f64"))
    oi, o, f = S.box_3d(ref, nx=nx, ny=ny, nz=nz, sd_conc=sd_conc, rain_mode=True, dt=2.0, sstp_coal=2)
    o.cond, o.rcyc = 0, 1
    p_r = ref.factory(L.backend_t.serial, oi)
    p_r.init(f["th"], f["rv"], f["rhod"], None, f["Cx"], f["Cy"], f["Cz"])
    fp = {k: v.copy() for k, v in f.items()}
    p_p = port_box3d(fp, nx, ny, nz, sd_conc, eff)
    p_p.dt, p_p.sstp_coal = 2.0, 2
    p_p.init(fp["th"], fp["rv"], fp["rhod"], fp["Cx"], fp["Cy"], fp["Cz"])
    size0, recycled = p_p.n.size, 0
    for step in range(8):
        p_r.step_sync(o, f["th"], f["rv"], f["rhod"], f["Cx"], f["Cy"], f["Cz"]); p_r.step_async(o)
        p_p.step_sync(fp["th"], fp["rv"], fp["rhod"], cond=False)
        dead_before = p_p.n.copy()
        p_p.step_async(cond=False, rcyc=True)
        assert np.array_equal(p_r.get_n(), p_p.n), step
        for k, a in (("rd3", p_p.rd3), ("rw2", p_p.rw2), ("x", p_p.x), ("y", p_p.y), ("z", p_p.z)):
            assert np.array_equal(p_r.get_attr(k), a), (k, step)
        assert p_p.n.size == size0                        # recycled, not removed
        recycled += p_p.n_recycled
    assert recycled > 0


@pytest.mark.parametrize("variant", ["mix", "nomix", "adaptive", "adaptive_act"])
def test_port_perparticle_substepping_matches_reference(ref, variant):
    """exact_sstp_cond in all its flavours (condensation/perparticle/*.ipp): the restatement follows the reference bit for bit,
    including the per-SD records of rv / th / rhod surviving advection, coalescence and removal"""
    nx, ny, nz, sd_conc = 3, 2, 4, 8
    eff = np.fromfile(os.path.join(ROOT, "libcloudphxx_b200", "data", "hall_davis_no_waals.f64"))
    kw = dict(mix=dict(sstp_cond=3, sstp_cond_mix=True), nomix=dict(sstp_cond=3, sstp_cond_mix=False),
              adaptive=dict(sstp_cond=4, sstp_cond_mix=False, adaptive_sstp_cond=True),
              adaptive_act=dict(sstp_cond=4, sstp_cond_mix=False, adaptive_sstp_cond=True, sstp_cond_act=8))[variant]
    oi, o, f = S.box_3d(ref, nx=nx, ny=ny, nz=nz, sd_conc=sd_conc, rain_mode=True, sstp_cond=kw["sstp_cond"])
    oi.exact_sstp_cond = 1
    oi.sstp_cond_mix = int(kw["sstp_cond_mix"])
    oi.adaptive_sstp_cond = int(kw.get("adaptive_sstp_cond", False))
    oi.sstp_cond_act = kw.get("sstp_cond_act", 1)
    oi.sstp_cond_adapt_drw2_eps, oi.sstp_cond_adapt_drw2_max = 1e-3, 2.0
    p_r = ref.factory(L.backend_t.serial, oi)
    p_r.init(f["th"], f["rv"], f["rhod"], None, f["Cx"], f["Cy"], f["Cz"])
    fp = {k: v.copy() for k, v in f.items()}
    p_p = port.Particles(nx=nx, ny=ny, nz=nz, dx=20., dy=20., dz=20., dt=1., x1=nx * 20., y1=ny * 20., z1=nz * 20., sd_conc=sd_conc,
                         n_sd_max=int(nx * ny * nz * sd_conc * 1.5), kernel="efficiencies", kernel_params={"eff": eff[1:], "r_max": eff[0]},
                         dry_distros=[(0.61, lognormal_as_capi(S.AEROSOL_ICICLE)), (1.28, lognormal_as_capi([(30e-6, 1.2, 1e5)]))],
                         exact_sstp_cond=True, sstp_cond_adapt_drw2_eps=1e-3, sstp_cond_adapt_drw2_max=2.0, **kw)
    p_p.init(fp["th"], fp["rv"], fp["rhod"], fp["Cx"], fp["Cy"], fp["Cz"])
    for step in range(3):
        p_r.step_sync(o, f["th"], f["rv"], f["rhod"], f["Cx"], f["Cy"], f["Cz"]); p_r.step_async(o)
        th, rv = p_p.step_sync(fp["th"], fp["rv"], fp["rhod"])
        fp["th"][:], fp["rv"][:] = th.reshape(fp["th"].shape), rv.reshape(fp["rv"].shape)
        p_p.step_async()
        assert np.array_equal(p_r.get_n(), p_p.n), step
        for k, a in (("rd3", p_p.rd3), ("rw2", p_p.rw2), ("x", p_p.x), ("y", p_p.y), ("z", p_p.z)):
            assert np.array_equal(p_r.get_attr(k), a), (k, step, S.rel_err(p_r.get_attr(k), a))
        assert np.array_equal(f["th"], fp["th"]) and np.array_equal(f["rv"], fp["rv"]), (step, S.rel_err(f["th"], fp["th"]), S.rel_err(f["rv"], fp["rv"]))


@pytest.mark.parametrize("kind", ["const_multi", "const_multi_user_range", "large_tail", "dry_sizes", "conc_factor", "sd_conc_user_range"])
def test_port_initialisation_flavours_match_reference(ref, kind):
    """every way of creating SDs at t = 0 (init_SD_with_distros*.ipp, init_SD_with_sizes.ipp): same attributes to the last bit"""
    nx, ny, nz = 3, 2, 5
    oi, o, f = S.box_3d(ref, nx=nx, ny=ny, nz=nz, sd_conc=16, n_sd_max=200000)
    kw = dict(sd_conc=16, dry_distros=[(0.61, lognormal_as_capi(S.AEROSOL_ICICLE))])
    if kind == "const_multi":
        oi.sd_conc, oi.sd_const_multi = 0, int(2e9)
        kw.update(sd_conc=0, sd_const_multi=int(2e9))
    elif kind == "const_multi_user_range":
        oi.sd_conc, oi.sd_const_multi, oi.rd_min, oi.rd_max = 0, int(1e9), 5e-9, 2e-6
        kw.update(sd_conc=0, sd_const_multi=int(1e9), rd_min=5e-9, rd_max=2e-6)
    elif kind == "large_tail":
        oi.sd_conc_large_tail, oi.sd_conc = 1, 256
        oi.dry_distros = [L.lognormal(0.61, S.AEROSOL_ICICLE), L.lognormal(1.28, [(0.5e-6, 1.6, 2e3)])]
        kw.update(sd_conc=256, sd_conc_large_tail=True,
                  dry_distros=[(0.61, lognormal_as_capi(S.AEROSOL_ICICLE)), (1.28, lognormal_as_capi([(0.5e-6, 1.6, 2e3)]))])
    elif kind == "dry_sizes":
        oi.dry_sizes = {0.3: {0.1e-6: (30e6, 5), 0.5e-6: (1e6, 2)}, (0.9, 0.0): {1e-6: (2e5, 3)}}
        kw.update(dry_sizes=[(0.3, {0.1e-6: (30e6, 5), 0.5e-6: (1e6, 2)}), (0.9, {1e-6: (2e5, 3)})])
    elif kind == "conc_factor":
        oi.aerosol_independent_of_rhod, oi.aerosol_conc_factor = 1, [1.0, 0.5, 2.0, 0.25, 1.5]
        kw.update(aerosol_independent_of_rhod=True, aerosol_conc_factor=[1.0, 0.5, 2.0, 0.25, 1.5])
    elif kind == "sd_conc_user_range":
        oi.rd_min, oi.rd_max = 1e-9, 5e-6
        kw.update(rd_min=1e-9, rd_max=5e-6)
    p_r = ref.factory(L.backend_t.serial, oi)
    p_r.init(f["th"], f["rv"], f["rhod"], None, f["Cx"], f["Cy"], f["Cz"])
    p_p = port.Particles(nx=nx, ny=ny, nz=nz, dx=20., dy=20., dz=20., dt=1., x1=nx * 20., y1=ny * 20., z1=nz * 20., n_sd_max=200000,
                         kernel="efficiencies", kernel_params={}, **kw)
    p_p.init(f["th"], f["rv"], f["rhod"], f["Cx"], f["Cy"], f["Cz"])
    assert p_p.n.size > 0 and np.array_equal(p_r.get_n(), p_p.n)
    for k, a in (("rd3", p_p.rd3), ("rw2", p_p.rw2), ("kappa", p_p.kpa), ("x", p_p.x), ("y", p_p.y), ("z", p_p.z)):
        assert np.array_equal(p_r.get_attr(k), a), (kind, k)


def slab_system(size, nx, ny, nz, sd_conc, cx, **kw):
    f = {"th": np.full((nx, ny, nz), 289.), "rv": np.full((nx, ny, nz), 7.5e-3), "rhod": np.full((nx, ny, nz), 1.1),
         "Cx": np.full((nx + 1, ny, nz), cx), "Cy": np.zeros((nx, ny + 1, nz)), "Cz": np.zeros((nx, ny, nz + 1))}
    eff = np.fromfile(os.path.join(ROOT, "libcloudphxx_b200", "data", "hall_davis_no_waals.f64"))
    p = port.SlabParticles(size, nx=nx, ny=ny, nz=nz, dx=20., dy=20., dz=20., dt=1., x1=nx * 20., y1=ny * 20., z1=nz * 20., sd_conc=sd_conc,
                           n_sd_max=int(nx * ny * nz * sd_conc * 2), kernel="efficiencies", kernel_params={"eff": eff[1:], "r_max": eff[0]},
                           dry_distros=[(0.61, lognormal_as_capi(S.AEROSOL_ICICLE))], **kw)
    p.init(f["th"], f["rv"], f["rhod"], f["Cx"], f["Cy"], f["Cz"])
    return p, f


def cell_contents(p, shape):
    """per cell of the global grid: the sorted dry volumes of the SDs in it (attributes travel verbatim, so these compare exactly)"""
    out = np.empty(shape, dtype=object)
    nx0 = 0
    for q in p.slabs:
        i, j, k = q.unravel(q.ijk)
        for c in np.ndindex(q.nx, shape[1], shape[2]):
            sel = (i == c[0]) & (j == c[1]) & (k == c[2])
            out[nx0 + c[0], c[1], c[2]] = tuple(np.sort(q.rd3[sel]))
        nx0 += q.nx
    return out


@pytest.mark.parametrize("size,nx", [(2, 5), (3, 7)])
def test_port_slabs_advect_a_pattern_once_around_the_ring(size, nx):
    """what the reference's own distributed test demands (tests/mpi/mpi_adve_test.cpp:138-254): with a Courant number of one the
    content of every cell moves one cell per step across the slab faces and is back after nx steps; slabs of unequal width
    (distmem_opts.hpp:10-18), every SD inside its slab after each exchange"""
    shape = (nx, 2, 3)
    p, f = slab_system(size, *shape, sd_conc=6, cx=1.0)
    assert [q.nx for q in p.slabs] == [port.slab_nx(nx, r, size) for r in range(size)] and sum(q.nx for q in p.slabs) == nx
    before = cell_contents(p, shape)
    n0 = sum(q.n_part for q in p.slabs)
    for step in range(nx):
        p.step_sync(f["th"], f["rv"], f["rhod"], cond=False)
        p.step_async(adve=True, sedi=False, coal=False, cond=False)
        assert sum(q.n_part for q in p.slabs) == n0
        for q in p.slabs:
            assert ((q.x >= q.x0) & (q.x < q.x1)).all()
        now = cell_contents(p, shape)
        assert (np.roll(before, step + 1, axis=0) == now).all(), step
    assert (before == cell_contents(p, shape)).all()


def test_port_slabs_order_of_arrival():
    """every slab appends the batch from its right neighbour first, then the one from its left neighbour, each in the sender's
    ascending storage order (particles_multi_gpu_impl_step_async_and_copy.ipp:104,134,161,190)"""
    p, f = slab_system(3, 6, 1, 2, sd_conc=4, cx=0.0)
    f["Cx"][:] = np.array([0.9, 0.9, 0.9, 0., -0.9, -0.9, -0.9])[:, None, None]    # both outer slabs push SDs into the middle one
    for q, r in zip(p.slabs, range(3)):
        q.Cx = p._cut(f["Cx"], r, 1)
    mid = p.slabs[1]
    keep = mid.rd3.copy()
    p.step_sync(f["th"], f["rv"], f["rhod"], cond=False)
    lft_src, rgt_src = p.slabs[0], p.slabs[2]
    # run the pre-copy part by hand to see who leaves
    for q in p.slabs:
        q.step_async(adve=True, sedi=False, coal=False, cond=False)
    from_right = rgt_src.rd3[rgt_src.lft_id].copy()
    from_left = lft_src.rd3[lft_src.rgt_id].copy()
    stay = np.delete(keep, np.concatenate([mid.lft_id, mid.rgt_id]))
    assert from_right.size and from_left.size, "both neighbours must send something"
    # now the real thing on a fresh copy of the same system (same seeds: identical state)
    p2, _ = slab_system(3, 6, 1, 2, sd_conc=4, cx=0.0)
    for q, r in zip(p2.slabs, range(3)):
        q.Cx = p2._cut(f["Cx"], r, 1)
    p2.step_sync(f["th"], f["rv"], f["rhod"], cond=False)
    p2.step_async(adve=True, sedi=False, coal=False, cond=False)
    assert np.array_equal(p2.slabs[1].rd3, np.concatenate([stay, from_right, from_left]))


def test_port_slabs_full_microphysics_conserves_dry_volume():
    """cond + coal + sedi + adve over three slabs: the dry aerosol volume is conserved up to what rains out"""
    p, f = slab_system(3, 6, 2, 4, sd_conc=8, cx=0.5)
    vol = lambda: sum(float((q.n.astype(np.float64) * q.rd3).sum()) for q in p.slabs)
    v0 = vol()
    for step in range(4):
        th, rv = p.step_sync(f["th"], f["rv"], f["rhod"])
        f["th"][:], f["rv"][:] = th, rv
        p.step_async()
    fallen = sum(q.puddle["dry_volume"] for q in p.slabs) / (4. / 3. * np.pi)
    assert abs(vol() + fallen - v0) <= 1e-12 * v0


def test_reference_float_instantiation_tracks_its_double_one():
    """the oracle's single-precision arm (factory<float> of the reference build, lgcf_* binding): same case, same seed; the float
    run's root search stops on a 2^-7 bracket (8 bits), so wet radii agree with the double run to a few per cent, fields to 1e-3"""
    out = {}
    for real in ("f64", "f32"):
        ref = S.oracle_library(real)
        oi, o, f = S.parcel(ref, n_sd=500)
        f = {k: np.ascontiguousarray(v, dtype=ref.dtype) for k, v in f.items()}
        p = ref.factory(L.backend_t.serial, oi)
        p.init(f["th"], f["rv"], f["rhod"])
        for _ in range(5):
            p.step_sync(o, f["th"], f["rv"], f["rhod"])
            p.step_async(o)
        out[real] = (p.get_n(), p.get_attr("rw2"), float(f["th"][0]), float(f["rv"][0]))
        assert out[real][1].dtype == ref.dtype
    assert out["f32"][0].size == out["f64"][0].size == 500
    r64, r32 = np.sort(out["f64"][1]), np.sort(out["f32"][1].astype(np.float64))
    assert np.median(np.abs(r32 - r64) / r64) < 2e-2
    assert abs(out["f32"][2] - out["f64"][2]) < 1e-3 * out["f64"][2] and abs(out["f32"][3] - out["f64"][3]) < 1e-2 * out["f64"][3]
