"""BASELINE.json configs[0..2] at their full sizes, as parity cases against the reference's CPU back-end (configs[3] is the
bench workload: tests/test_gpu_fullsize.py checks its invariants, bench.py measures it):

  cfg1  0-D box, Golovin kernel, 2^17 SDs                      -> exact (multiplicities, dry radii), rw2 to cbrt ulps
  cfg2  adiabatic parcel, 10^6 SDs, condensation / activation  -> tolerance class of the root solve, th / rv follow
  cfg3  kinematic 2-D (ICMW case 1 shape) 76x76 cells x 128 SD, cond + coal + sedi + adve, 10 / 10 sub-steps
                                                                -> integer state exact, floating-point state to tolerance
A few steps each: the oracle is serial."""
import numpy as np
import pytest

from libcloudphxx_b200 import lgrngn as L
from tests import support as S

pytestmark = pytest.mark.gpu


def test_cfg1_golovin_box_2p17(ref, b200):
    hist = []

    def check(step, p_r, p_n, f_r, f_n):
        n_r, n_n = p_r.get_n(), p_n.get_n()
        assert np.array_equal(n_r, n_n), "multiplicities differ at step %d" % step
        assert np.array_equal(p_r.get_attr("rd3"), p_n.get_attr("rd3")), step
        assert S.rel_err(p_r.get_attr("rw2"), p_n.get_attr("rw2")) < 1e-14, step
        hist.append((n_r.size, int(n_r.sum())))
    S.run_pair(ref, b200, S.box_golovin, 12, on_step=check, n_sd=2 ** 17)
    assert hist[0][0] == 2 ** 17 and hist[-1][1] < hist[0][1]


def test_cfg2_parcel_1e6(ref, b200):
    def drive(lib):
        oi, o, f = S.parcel(lib, n_sd=10 ** 6, dt=0.1)
        p = lib.factory(L.backend_t.serial if lib.name == "reference" else L.backend_t.CUDA, oi)
        p.init(f["th"], f["rv"], f["rhod"])
        out = []
        for step in range(6):
            f["rhod"] *= 0.9992            # expansion fast enough to activate droplets within the few steps
            p.step_sync(o, f["th"], f["rv"], f["rhod"])
            p.step_async(o)
            out.append((f["th"][0], f["rv"][0], p.get_attr("rw2")))
        return out
    a, b = drive(ref), drive(b200)
    for step, ((th_r, rv_r, rw_r), (th_n, rv_n, rw_n)) in enumerate(zip(a, b)):
        assert rw_r.size == 10 ** 6 == rw_n.size
        err = np.abs(rw_r - rw_n) / rw_r
        if step == 0:
            assert err.max() < 2.0 ** -15, err.max()
        assert np.quantile(err, 0.999) < (step + 1) * 2.0 ** -15, (step, np.quantile(err, 0.999))
        assert np.median(err) < 1e-7, (step, np.median(err))
        assert abs(th_r - th_n) / th_r < 1e-9 and abs(rv_r - rv_n) / rv_r < 1e-7, (step, th_r, th_n, rv_r, rv_n)


def test_cfg3_kinematic_2d_76x76x128(ref, b200):
    def check(step, p_r, p_n, f_r, f_n):
        n_r, n_n = p_r.get_n(), p_n.get_n()
        assert n_r.size == n_n.size, (step, n_r.size, n_n.size)
        if step == -1:
            assert n_r.size == 76 * 76 * 128
        assert np.array_equal(n_r, n_n), "multiplicities differ at step %d" % step
        assert np.array_equal(p_r.get_attr("rd3"), p_n.get_attr("rd3")), step
        assert np.array_equal(p_r.get_attr("x"), p_n.get_attr("x")), step
        err = np.abs(p_r.get_attr("rw2") - p_n.get_attr("rw2")) / p_r.get_attr("rw2")
        assert np.quantile(err, 0.999) < (step + 2) * 10 * 2.0 ** -15, (step, np.quantile(err, 0.999))
        assert S.rel_err(p_r.get_attr("z"), p_n.get_attr("z")) < 1e-7, step
        assert S.rel_err(f_r["th"], f_n["th"]) < 1e-8 and S.rel_err(f_r["rv"], f_n["rv"]) < 1e-6, step
    S.run_pair(ref, b200, S.kinematic_2d, 3, on_step=check, nx=76, nz=76, sd_conc=128, sstp_cond=10, sstp_coal=10,
               kernel=L.kernel_t.hall_davis_no_waals, kparams=())
