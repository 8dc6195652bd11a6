"""The reference-style Python surface (`from libcloudphxx import lgrngn, common`) served by the B200 back-end.

Scenarios are this repository's own restatements of what the reference's Python tests pin (tests/python/unit/uniform_init.py,
lgrngn_adve.py:105, multiple_kappas.py, lgrngn_subsidence.py, physics/test_coal.py:96-103, physics/puddle.py,
physics/lgrngn_cond.py:131-187), written against the compat package only, with scripts' habit of asking for the serial /
OpenMP back-end redirected to CUDA (LIBCLOUDPHXX_COMPAT_REDIRECT=1)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def pkg():
    os.environ.pop("LIBCLOUDPHXX_COMPAT_LIBRARY", None)
    os.environ["LIBCLOUDPHXX_COMPAT_REDIRECT"] = "1"
    sys.path.insert(0, os.path.join(ROOT, "libcloudphxx_b200", "compat"))
    import libcloudphxx
    yield libcloudphxx
    sys.path.remove(os.path.join(ROOT, "libcloudphxx_b200", "compat"))


def lognormal(mean_r, stdev, n_tot):
    def n_of_lnr(lnr):
        return n_tot * np.exp(-(lnr - np.log(mean_r)) ** 2 / 2 / np.log(stdev) ** 2) / np.log(stdev) / np.sqrt(2 * np.pi)
    return n_of_lnr


def base(lgrngn, nx=0, ny=0, nz=0, sd_conc=64, d=10.0):
    oi = lgrngn.opts_init_t()
    oi.dt = 1
    oi.nx, oi.ny, oi.nz = nx, ny, nz
    oi.dx = oi.dy = oi.dz = d
    oi.x1, oi.y1, oi.z1 = max(nx, 1) * d, max(ny, 1) * d, max(nz, 1) * d
    oi.sd_conc = sd_conc
    oi.n_sd_max = max(nx, 1) * max(ny, 1) * max(nz, 1) * sd_conc
    oi.dry_distros = {0.61: lognormal(0.04e-6, 1.4, 60e6)}
    oi.kernel = lgrngn.kernel_t.geometric
    oi.terminal_velocity = lgrngn.vt_t.beard77fast
    oi.coal_switch = oi.sedi_switch = False
    return oi


def fields(shape, th=300., rv=0.01, rhod=1.1):
    return np.full(shape, th), np.full(shape, rv), np.full(shape, rhod)


def test_uniform_init_puts_sd_conc_droplets_in_every_cell(pkg):
    lgrngn = pkg.lgrngn
    oi = base(lgrngn, 4, 3, 5, sd_conc=32)
    p = lgrngn.factory(lgrngn.backend_t.serial, oi)            # redirected to CUDA
    p.init(*fields((4, 3, 5)))
    p.diag_all()
    p.diag_sd_conc()
    conc = np.frombuffer(p.outbuf()).reshape(4, 3, 5)
    assert (conc == 32).all()
    p.diag_all()
    p.diag_wet_mom(0)
    n0 = np.frombuffer(p.outbuf())
    assert n0.std() / n0.mean() < 0.06          # same spectrum everywhere: number concentration uniform up to sampling noise


def test_advection_moves_the_pattern_by_one_cell_per_step(pkg):
    lgrngn = pkg.lgrngn
    oi = base(lgrngn, nx=6, nz=1, sd_conc=16)
    oi.adve_scheme = lgrngn.as_t.euler
    p = lgrngn.factory(lgrngn.backend_t.OpenMP, oi)
    th, rv, rhod = fields((6, 1))
    rhod[:] = 1.0
    Cx, Cz = np.ones((7, 1)), np.zeros((6, 2))
    # different air density column by column -> different number concentration per column at init (aerosol scales with rhod)
    rhod[:, 0] = [1.0, 0.9, 0.8, 1.1, 1.2, 0.7]
    p.init(th, rv, rhod, Cx=Cx, Cz=Cz)
    o = lgrngn.opts_t()
    o.cond = o.coal = o.sedi = False
    p.diag_all(); p.diag_wet_mom(0)
    before = np.frombuffer(p.outbuf()).copy() * rhod[:, 0]        # moments are per mass of dry air
    for step in range(1, 7):
        p.step_sync(o, th, rv, rhod, Cx=Cx, Cz=Cz)
        p.step_async(o)
        p.diag_all(); p.diag_wet_mom(0)
        now = np.frombuffer(p.outbuf()) * rhod[:, 0]
        assert np.allclose(now, np.roll(before, step), rtol=1e-12), step     # Courant number 1, periodic domain


def test_two_species_keep_their_kappas(pkg):
    lgrngn = pkg.lgrngn
    oi = base(lgrngn, sd_conc=1000)
    oi.dry_distros = {0.3: lognormal(0.04e-6, 1.4, 60e6), (1.2, 0.0): lognormal(0.1e-6, 1.6, 10e6)}
    oi.n_sd_max = 1000
    p = lgrngn.factory(lgrngn.backend_t.CUDA, oi)
    p.init(*fields(1))
    kappa = np.asarray(p.get_attr("kappa"))
    assert set(np.unique(kappa)) == {0.3, 1.2}
    p.diag_kappa_rng(0.25, 0.35); p.diag_dry_mom(0)
    # per kg of dry air; with several spectra the reference scales multiplicities by the INTEGER quotient
    # sd_conc / int(fraction * sd_conc + 0.5) (init_SD_with_distros_sd_conc.ipp:29) - kept, hence the factor
    n_this = int((kappa == 0.3).sum())
    quirk = (1000 // n_this) / (1000. / n_this)
    assert np.frombuffer(p.outbuf())[0] == pytest.approx(60e6 / pkg.common.rho_stp * quirk, rel=0.03)


def test_coalescence_conserves_dry_and_wet_volume(pkg):
    lgrngn = pkg.lgrngn
    oi = base(lgrngn, sd_conc=2 ** 12)
    oi.n_sd_max = 2 ** 12
    oi.coal_switch = True
    oi.kernel = lgrngn.kernel_t.golovin
    oi.kernel_parameters = [1500.]
    oi.dry_distros = {1e-10: lambda lnr: 2 ** 23 * 3 * (np.exp(lnr) / 30.084e-6) ** 3 * np.exp(-(np.exp(lnr) / 30.084e-6) ** 3)}
    p = lgrngn.factory(lgrngn.backend_t.CUDA, oi)
    th, rv, rhod = fields(1, rhod=1.0)
    p.init(th, rv, rhod)
    o = lgrngn.opts_t()
    o.adve = o.sedi = o.cond = False

    def moments():
        out = []
        for diag, k in ((p.diag_wet_mom, 3), (p.diag_dry_mom, 3), (p.diag_wet_mom, 0)):
            p.diag_all(); diag(k)
            out.append(np.frombuffer(p.outbuf())[0])
        return out
    w3, d3, n0 = moments()
    for _ in range(100):
        p.step_sync(o, th, rv, rhod)
        p.step_async(o)
    w3b, d3b, n0b = moments()
    assert w3b == pytest.approx(w3, rel=1e-10) and d3b == pytest.approx(d3, rel=1e-10)      # test_coal.py:96-103
    assert n0b < 0.9 * n0                                                                    # and it did coalesce


def test_rain_leaves_through_the_bottom_into_the_puddle(pkg):
    lgrngn = pkg.lgrngn
    oi = base(lgrngn, nx=3, nz=2, sd_conc=8, d=100.)
    oi.sedi_switch = True
    oi.dry_sizes = {0.61: {200e-6: [1e3, 8]}}          # monodisperse 200 um particles, 8 SDs per cell
    oi.dry_distros = {}
    oi.sd_conc = 0
    oi.n_sd_max = 3 * 2 * 8
    p = lgrngn.factory(lgrngn.backend_t.CUDA, oi)
    th, rv, rhod = fields((3, 2), rv=0.002, rhod=1.0)
    p.init(th, rv, rhod, Cx=np.zeros((4, 2)), Cz=np.zeros((3, 3)))
    o = lgrngn.opts_t()
    o.cond = o.coal = o.adve = False
    n_before = np.asarray(p._p.get_n()).sum()
    for _ in range(400):
        p.step_sync(o, th, rv, rhod)
        p.step_async(o)
    puddle = p.diag_puddle()
    assert len(p.get_attr("rw2")) == 0                                   # everything fell out (vt ~ 1.6 m/s, 200 m, 400 s)
    assert puddle["particle_number"] == pytest.approx(float(n_before), rel=1e-12)
    assert puddle["dry_volume"] == pytest.approx(n_before * 4. / 3 * np.pi * 200e-6 ** 3, rel=1e-10)
    assert puddle["liquid_volume"] >= puddle["dry_volume"]


def test_condensation_closes_the_water_budget(pkg):
    lgrngn, common = pkg.lgrngn, pkg.common
    oi = base(lgrngn, sd_conc=256)
    oi.n_sd_max = 256
    oi.sstp_cond = 10
    oi.dt = 0.5
    oi.RH_max = 0.9999
    p = lgrngn.factory(lgrngn.backend_t.CUDA, oi)
    rhod = np.array([1.0])
    th_std = 300.
    T0 = 283.
    # start at RH = 99 %, then cool by lowering theta each step (an updraft seen by the parcel)
    p0 = common.p_hydro(0., th_std, 0., 0., 95000.)
    rv = np.array([0.99 * common.r_vs(T0, p0)])
    th = np.array([T0 / common.exner(p0)])
    th[0] = common.th_std2dry(th[0], rv[0])
    rhod[0] = common.rhod(p0, common.th_dry2std(th[0], rv[0]), rv[0])
    p.init(th, rv, rhod)
    o = lgrngn.opts_t()
    o.adve = o.sedi = o.coal = False

    def liquid():
        p.diag_all(); p.diag_wet_mom(3)
        return np.frombuffer(p.outbuf())[0] * 4. / 3 * np.pi * common.rho_w
    total0 = rv[0] + liquid()
    for _ in range(200):
        th[0] -= 0.01
        p.step_sync(o, th, rv, rhod)
        p.step_async(o)
    assert liquid() > 1e-4                                             # a cloud formed
    assert rv[0] + liquid() == pytest.approx(total0, rel=1e-9)         # lgrngn_cond.py:158-187 (water budget)
    T = common.T(th[0], rhod[0])
    S = common.p_v(common.p(rhod[0], rv[0], T), rv[0]) / common.p_vs(T) - 1
    assert abs(S) < 0.02                                               # supersaturation relaxed by the droplets


def test_step_order_errors_have_the_reference_texts(pkg):
    lgrngn = pkg.lgrngn
    oi = base(lgrngn, sd_conc=8)
    p = lgrngn.factory(lgrngn.backend_t.CUDA, oi)
    th, rv, rhod = fields(1)
    o = lgrngn.opts_t()
    with pytest.raises(RuntimeError, match="please call init\\(\\) before calling step_sync\\(\\)"):
        p.step_sync(o, th, rv, rhod)
    p.init(th, rv, rhod)
    with pytest.raises(RuntimeError, match="init\\(\\) may be called just once"):
        p.init(th, rv, rhod)
    with pytest.raises(RuntimeError, match="please call step_sync\\(\\) before calling step_async\\(\\) again"):
        p.step_async(o)
    p.step_sync(o, th, rv, rhod)
    with pytest.raises(RuntimeError, match="please call step_async\\(\\) before calling step_sync\\(\\) again"):
        p.step_sync(o, th, rv, rhod)


# ---- the reference's own test scripts, unchanged, on the CUDA back-end ----------------------------------------------------------
REF_SCRIPTS = ["unit/uniform_init.py", "unit/lgrngn_adve.py", "unit/terminal_velocities.py", "unit/multiple_kappas.py",
               "unit/adve_scheme.py", "unit/lgrngn_subsidence.py", "physics/test_coal.py", "physics/lgrngn_cond.py", "physics/puddle.py"]


@pytest.mark.parametrize("rng", ["mt19937", "philox"])
@pytest.mark.parametrize("script", REF_SCRIPTS)
def test_reference_script_runs_unchanged_on_the_cuda_backend(script, rng):
    """tests/golden/ref_scripts/ holds verbatim copies of the reference's tests/python/{unit,physics} scripts; each runs in its own
    interpreter against liblgrngn_b200.so, the serial / OpenMP back-ends they ask for redirected to CUDA (puddle.py asks for
    multi_CUDA itself), under the replayed mt19937 stream and under the default in-kernel Philox stream"""
    import subprocess
    path = os.path.join(ROOT, "tests", "golden", "ref_scripts", script)
    env = dict(os.environ)
    env.pop("LIBCLOUDPHXX_COMPAT_LIBRARY", None)
    env["PYTHONPATH"] = ROOT + os.pathsep + os.path.join(ROOT, "libcloudphxx_b200", "compat") + os.pathsep + env.get("PYTHONPATH", "")
    env["LIBCLOUDPHXX_COMPAT_REDIRECT"] = "1"
    env["LCX_RNG"] = rng
    r = subprocess.run([sys.executable, os.path.basename(path)], cwd=os.path.dirname(path), env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-2500:]
