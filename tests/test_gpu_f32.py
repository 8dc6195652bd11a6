"""Single-precision engine (factory<float>, liblcx_b200_f32.so) against the reference's OWN float instantiation
(src/lib.cpp:43, oracle/_ref through the lgcf_* binding) on the same seeded cases.

Bars: everything integer / decided by comparisons is exact (multiplicities, the set of SDs, dry radii, kappa, positions after
advection); wet radii after condensation agree within the bracket the reference's float root search stops at - 2^-7 relative
(sizeof(float) * 8 / 4 = 8 bits, src/detail/config.hpp:39, common/detail/toms748.hpp:262-286), a few float ulp otherwise;
th / rv within 1e-5 relative.  Measured on B200 (tools/diag_f32.py): rw2 <= 7.7e-3, z <= 1.2e-5, th <= 4e-7, rv <= 7.5e-6.
"""
import numpy as np
import pytest

from libcloudphxx_b200 import lgrngn as L
from tests import support as S

pytestmark = pytest.mark.gpu

TOL_COND = 2.0 ** -7 * 1.05      # float TOMS 748 stops on a bracket of relative width 2^-7


@pytest.fixture(scope="module")
def ref32():
    return S.oracle_library("f32")


@pytest.fixture(scope="module")
def b200_32():
    return S.b200_library("f32")


def test_f32_is_served_by_the_single_precision_engine(b200_32):
    """the float particle system owns an engine of liblcx_b200_f32.so whose kernels ran (launch counter of THAT library)"""
    from libcloudphxx_b200 import distributed as D
    oi, o, f = S.box_3d(b200_32, nx=4, ny=4, nz=4, sd_conc=16)
    f = {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in f.items()}
    p = b200_32.factory(L.backend_t.CUDA, oi)
    p.init(f["th"], f["rv"], f["rhod"], None, f["Cx"], f["Cy"], f["Cz"])
    eng = D.engine_of(b200_32, p)
    assert eng.real == "f32"
    l0 = eng.launches()
    p.step_sync(o, f["th"], f["rv"], f["rhod"], f["Cx"], f["Cy"], f["Cz"])
    p.step_async(o)
    assert eng.launches() > l0 + 10
    assert eng.n_part() == 4 * 4 * 4 * 16
    assert p.get_attr("rw2").dtype == np.float32 and f["th"].dtype == np.float32


def test_f32_golovin_box_exact(ref32, b200_32):
    def check(step, p_r, p_n, f_r, f_n):
        assert np.array_equal(p_r.get_n(), p_n.get_n()), step
        assert np.array_equal(p_r.get_attr("rd3"), p_n.get_attr("rd3")), step
        assert S.rel_err(p_r.get_attr("rw2"), p_n.get_attr("rw2")) < 1e-6, step        # cbrt: a few float ulp
    S.run_pair(ref32, b200_32, S.box_golovin, 6, on_step=check, n_sd=2 ** 12)


@pytest.mark.parametrize("sstp_cond", [1, 3])
def test_f32_parcel_condensation(ref32, b200_32, sstp_cond):
    def check(step, p_r, p_n, f_r, f_n):
        assert np.array_equal(p_r.get_n(), p_n.get_n()), step
        assert S.rel_err(p_r.get_attr("rw2"), p_n.get_attr("rw2")) < TOL_COND, step
        assert S.rel_err(f_r["th"], f_n["th"]) < 1e-5 and S.rel_err(f_r["rv"], f_n["rv"]) < 1e-4, step
    S.run_pair(ref32, b200_32, S.parcel, 6, on_step=check, n_sd=2000, sstp_cond=sstp_cond)


@pytest.mark.parametrize("adve", ["implicit", "euler", "pred_corr"])
def test_f32_full_step_3d(ref32, b200_32, adve):
    """cond + coal + sedi + adve in 3-D, two aerosol modes (kappa mixing on collision), every advection scheme"""
    scheme = getattr(L.as_t, adve)

    def check(step, p_r, p_n, f_r, f_n):
        n_r, n_n = p_r.get_n(), p_n.get_n()
        assert n_r.size == n_n.size and np.array_equal(n_r, n_n), step
        for a in ("rd3", "kappa", "x", "y"):
            assert np.array_equal(p_r.get_attr(a), p_n.get_attr(a)), (step, a)
        assert S.rel_err(p_r.get_attr("z"), p_n.get_attr("z")) < 1e-4, step             # sedimentation: vt to float accuracy
        assert S.rel_err(p_r.get_attr("rw2"), p_n.get_attr("rw2")) < TOL_COND, step
        assert S.rel_err(f_r["th"], f_n["th"]) < 1e-5 and S.rel_err(f_r["rv"], f_n["rv"]) < 1e-4, step
    S.run_pair(ref32, b200_32, S.box_3d, 4, on_step=check, nx=4, ny=4, nz=6, sd_conc=24, rain_mode=True, adve=scheme)


def test_f32_kinematic_2d(ref32, b200_32):
    def check(step, p_r, p_n, f_r, f_n):
        n_r, n_n = p_r.get_n(), p_n.get_n()
        assert n_r.size == n_n.size and np.array_equal(n_r, n_n), step
        assert np.array_equal(p_r.get_attr("x"), p_n.get_attr("x")), step
        assert S.rel_err(p_r.get_attr("rw2"), p_n.get_attr("rw2")) < TOL_COND, step
        assert S.rel_err(f_r["th"], f_n["th"]) < 1e-5 and S.rel_err(f_r["rv"], f_n["rv"]) < 1e-4, step
    S.run_pair(ref32, b200_32, S.kinematic_2d, 4, on_step=check)


def test_f32_diagnostics(ref32, b200_32):
    """selectors and moments through outbuf() in single precision"""
    p_r, p_n, f_r, f_n = S.run_pair(ref32, b200_32, S.box_3d, 2, nx=4, ny=4, nz=6, sd_conc=24)
    for p in (p_r, p_n):
        p.diag_all()
        p.diag_sd_conc()
    assert np.array_equal(p_r.outbuf(), p_n.outbuf())
    for sel, mom, bar in ((("diag_all",), ("diag_dry_mom", 3), 1e-5), (("diag_wet_rng", 0.5e-6, 25e-6), ("diag_wet_mom", 0), 2e-2),
                          (("diag_all",), ("diag_wet_mom", 3), 3e-2)):
        out = []
        for p in (p_r, p_n):
            getattr(p, sel[0])(*sel[1:])
            getattr(p, mom[0])(*mom[1:])
            out.append(p.outbuf().astype(np.float64).sum())
        assert abs(out[0] - out[1]) <= bar * abs(out[0]), (sel, mom, out)


def test_f32_multi_cuda_roundtrip_on_one_device(b200_32, monkeypatch):
    """the single-precision engine behind factory<float>(multi_CUDA): three unequal x-slabs (folded onto one GPU), migration through
    the float inboxes; advecting once round the periodic domain rolls every per-cell statistic exactly and returns it unchanged"""
    monkeypatch.setenv("LCX_SLABS_ON_ONE_DEVICE", "1")
    nx = 7
    oi, o, f = S.box_3d(b200_32, nx=nx, ny=3, nz=4, sd_conc=16, rain_mode=True)
    oi.dev_count = 3
    oi.n_sd_max = int(oi.n_sd_max * 2)
    f = {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in f.items()}
    f["Cx"][:] = 1.0
    f["Cy"][:] = 0.0
    o.cond = o.coal = o.sedi = 0
    p = b200_32.factory(L.backend_t.multi_CUDA, oi)
    p.init(f["th"], f["rv"], f["rhod"], None, f["Cx"], f["Cy"], f["Cz"])

    def per_cell():
        out = []
        for mom in (p.diag_sd_conc, lambda: p.diag_dry_mom(1), lambda: p.diag_kappa_mom(1)):
            p.diag_all(); mom(); out.append(p.outbuf().reshape(nx, 3, 4).copy())
        return out
    before = per_cell()
    # two spectra share sd_conc = 16: int(fraction * 16) each (init_count_num.ipp:32-35), 15 in all - the same in every cell
    assert before[0][0, 0, 0] in (15, 16) and (before[0] == before[0][0, 0, 0]).all()
    for step in range(nx):
        p.step_sync(o, f["th"], f["rv"], f["rhod"], f["Cx"], f["Cy"], f["Cz"])
        p.step_async(o)
        for a, b in zip(before, per_cell()):
            assert np.array_equal(np.roll(a, step + 1, axis=0), b), step
    for a, b in zip(before, per_cell()):
        assert np.array_equal(a, b)
