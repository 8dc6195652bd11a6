"""x-slab decomposition: the in-process multi_CUDA back-end and the one-process-per-GPU (torchrun + NCCL) mode.

No CPU oracle exists for partitioned runs (the reference needs MPI for that, absent here), so the checks are the ones the
reference's own distributed test makes (tests/mpi/mpi_adve_test.cpp:143-256) plus transport independence:
  * advecting once round the periodic domain with C = +1 rolls every per-cell statistic exactly and returns it unchanged,
    with unequal slabs; nothing is lost or duplicated;
  * the same slabs on one device and on several devices give bit-identical results;
  * one slab of multi_CUDA (dev_count = 1) is the CUDA back-end.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from libcloudphxx_b200 import lgrngn as L
from tests import support as S

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def n_devices():
    import torch
    return torch.cuda.device_count()


def per_cell(p, shape):
    out = []
    for mom in (p.diag_sd_conc, lambda: p.diag_dry_mom(1), lambda: p.diag_wet_mom(1), lambda: p.diag_kappa_mom(1), lambda: p.diag_dry_mom(0)):
        p.diag_all(); mom(); out.append(p.outbuf().reshape(shape).copy())
    return out


def roundtrip_case(b200, dev_count, nx=5, scheme=L.as_t.implicit):
    oi, o, f = S.box_3d(b200, nx=nx, ny=3, nz=4, sd_conc=16, rain_mode=True, adve=scheme)
    oi.dev_count = dev_count
    oi.n_sd_max = int(oi.n_sd_max * 2)       # the per-slab capacity n_sd_max / G + 1 must cover the widest slab
    f["Cx"][:] = 1.0
    f["Cy"][:] = 0.0
    o.cond = o.coal = o.sedi = 0
    return oi, o, f


@pytest.mark.parametrize("scheme", [L.as_t.implicit, L.as_t.pred_corr])
@pytest.mark.parametrize("slabs", [2, 3])
def test_multi_cuda_roundtrip_on_one_device(b200, monkeypatch, slabs, scheme):
    """unequal x-slabs folded onto one GPU: migration rolls the per-cell statistics exactly, once round = identity (with the
    predictor-corrector scheme this also needs the Courant halo of every slab filled: an empty halo would stop the SDs of the
    last column half way)"""
    monkeypatch.setenv("LCX_SLABS_ON_ONE_DEVICE", "1")
    nx = 5 if slabs == 2 else 7
    oi, o, f = roundtrip_case(b200, slabs, nx=nx, scheme=scheme)
    p = b200.factory(L.backend_t.multi_CUDA, oi)
    p.init(f["th"], f["rv"], f["rhod"], None, f["Cx"], f["Cy"], f["Cz"])
    shape = (nx, 3, 4)
    before = per_cell(p, shape)
    assert before[0].sum() > 0
    for step in range(nx):
        p.step_sync(o, f["th"], f["rv"], f["rhod"], f["Cx"], f["Cy"], f["Cz"])
        p.step_async(o)
        now = per_cell(p, shape)
        for a, b in zip(before, now):
            assert np.array_equal(np.roll(a, step + 1, axis=0), b), step
    for a, b in zip(before, per_cell(p, shape)):
        assert np.array_equal(a, b)


def test_multi_cuda_single_slab_is_cuda(b200):
    """dev_count = 1: the multi_CUDA object must reproduce the CUDA back-end bit for bit (full microphysics)"""
    res = []
    for backend in (L.backend_t.CUDA, L.backend_t.multi_CUDA):
        oi, o, f = S.box_3d(b200, nx=4, ny=3, nz=6, sd_conc=24, rain_mode=True)
        oi.dev_count = 1
        p = b200.factory(backend, oi)
        p.init(f["th"], f["rv"], f["rhod"], None, f["Cx"], f["Cy"], f["Cz"])
        for _ in range(4):
            p.step_sync(o, f["th"], f["rv"], f["rhod"], f["Cx"], f["Cy"], f["Cz"])
            p.step_async(o)
        res.append(per_cell(p, (4, 3, 6)) + [f["th"].copy(), f["rv"].copy()])
    for a, b in zip(*res):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("exact_sstp", [0, 1])
def test_multi_cuda_full_microphysics_runs_and_conserves(b200, monkeypatch, exact_sstp):
    """cond + coal + sedi + adve over 3 slabs: dry aerosol volume is conserved up to what rains out; with per-particle
    sub-stepping the migrants also carry their rv / th / rhod records (particles_impl.ipp:452-459)"""
    monkeypatch.setenv("LCX_SLABS_ON_ONE_DEVICE", "1")
    oi, o, f = S.box_3d(b200, nx=6, ny=4, nz=6, sd_conc=24, rain_mode=True, cx=0.5, sstp_cond=2 if exact_sstp else 1)
    oi.exact_sstp_cond = exact_sstp
    oi.dev_count = 3
    p = b200.factory(L.backend_t.multi_CUDA, oi)
    p.init(f["th"], f["rv"], f["rhod"], None, f["Cx"], f["Cy"], f["Cz"])
    dv_rhod = 20.0 ** 3 * f["rhod"]

    def dry_volume():
        p.diag_all(); p.diag_dry_mom(3)
        return float((p.outbuf().reshape(6, 4, 6) * dv_rhod).sum()) * 4. / 3 * np.pi
    v0 = dry_volume()
    for _ in range(5):
        p.step_sync(o, f["th"], f["rv"], f["rhod"], f["Cx"], f["Cy"], f["Cz"])
        p.step_async(o)
    v1 = dry_volume()
    fallen = p.diag_puddle()["dry_volume"]
    assert abs(v1 + fallen - v0) <= 1e-10 * v0, (v0, v1, fallen)
    assert np.isfinite(f["th"]).all() and np.isfinite(f["rv"]).all() and (f["rv"] > 0).all()


@pytest.mark.skipif("n_devices() < 2")
def test_multi_cuda_devices_match_fold(b200, monkeypatch):
    """transport independence: 2 slabs on 2 GPUs (peer copies) == the same 2 slabs on one GPU, bit for bit"""
    res = []
    for fold in ("1", "0"):
        monkeypatch.setenv("LCX_SLABS_ON_ONE_DEVICE", fold)
        oi, o, f = S.box_3d(b200, nx=6, ny=4, nz=6, sd_conc=24, rain_mode=True, cx=0.5)
        oi.dev_count = 2
        p = b200.factory(L.backend_t.multi_CUDA, oi)
        p.init(f["th"], f["rv"], f["rhod"], None, f["Cx"], f["Cy"], f["Cz"])
        for _ in range(5):
            p.step_sync(o, f["th"], f["rv"], f["rhod"], f["Cx"], f["Cy"], f["Cz"])
            p.step_async(o)
        res.append(per_cell(p, (6, 4, 6)) + [f["th"].copy(), f["rv"].copy()])
    for a, b in zip(*res):
        assert np.array_equal(a, b)


def test_torchrun_ranks_roundtrip():
    """one process per slab (one per GPU where there are enough; on a single-GPU box the ranks share the device and only the
    transport differs: CUDA IPC inside one device instead of NVLink), migrants packed straight into the neighbours' inboxes"""
    n = min(max(n_devices(), 2), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tests", "dist_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0 and "DIST_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


def test_torchrun_ranks_full_microphysics_conserve():
    """process-distributed slabs, cfg5-shaped (Cx = 0.5, rain mode, cond + coal + sedi + adve under Philox): conservation of the dry
    volume over all ranks, sent == received for every face and step"""
    n = min(max(n_devices(), 2), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", "29519", os.path.join(ROOT, "tests", "dist_worker.py"), "--full"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0 and "DIST_FULL_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


@pytest.mark.parametrize("slabs", [2, 3])
def test_multi_cuda_cfg5_shaped_conserves(b200, monkeypatch, slabs):
    """the same through the in-process multi_CUDA back-end (slabs folded onto one device when there are fewer GPUs)"""
    if n_devices() < slabs:
        monkeypatch.setenv("LCX_SLABS_ON_ONE_DEVICE", "1")
    from libcloudphxx_b200 import distributed as D
    with S.rng_mode(b200, 0, -1):
        oi, o, f = S.box_3d(b200, nx=6 * slabs, ny=8, nz=10, sd_conc=40, rain_mode=True, cx=0.5)
        oi.dev_count = slabs
        p = b200.factory(L.backend_t.multi_CUDA, oi)
    p.init(f["th"], f["rv"], f["rhod"], None, f["Cx"], f["Cy"], f["Cz"])
    engines = [D.engine_of(b200, p, d) for d in range(D.n_slabs(b200, p))]
    assert len(engines) == slabs

    def total():
        p.diag_all(); p.diag_dry_mom(3)
        live = float((p.outbuf().reshape(f["rhod"].shape) * f["rhod"]).sum() * 20.0 ** 3 * 4.0 / 3.0 * np.pi)
        return live + p.diag_puddle()["dry_volume"] + sum(e.top_loss()[0] for e in engines)
    v0 = total()
    for _ in range(24):
        p.step_sync(o, f["th"], f["rv"], f["rhod"], f["Cx"], f["Cy"], f["Cz"])
        p.step_async(o)
        s = D.migr_stats(b200, p)
        assert s[0] + s[1] == s[2] + s[3] and s[1] > 0
    assert abs(total() - v0) <= 1e-10 * v0
    assert p.diag_puddle()["dry_volume"] > 0


@pytest.mark.parametrize("slabs", [2, 3])
def test_multi_cuda_pred_corr_halo_holds_the_neighbours_columns(b200, monkeypatch, slabs):
    """predictor-corrector advection reads Courant numbers two columns beyond an SD's cell (adve.ipp:183-303): across a slab face
    these must be the NEIGHBOUR's columns (the reference's MPI build exchanges them, particles_impl_xchng_courants.ipp:15-153; the
    slabs of multi_CUDA map their halo from the caller's global array, init_e2l.ipp:34-114).  Checked on the device arrays themselves:
    every slab's Cx / Cy / Cz, halo included, against an independent numpy statement of the rule, with a field whose every value is
    unique; then the scheme runs and keeps every super-droplet"""
    import ctypes as C
    from libcloudphxx_b200 import distributed as D, engine as E
    if n_devices() < slabs:
        monkeypatch.setenv("LCX_SLABS_ON_ONE_DEVICE", "1")
    nx, ny, nz, halo = 4 * slabs + 1, 3, 6, 2
    oi, o, f = S.box_3d(b200, nx=nx, ny=ny, nz=nz, sd_conc=16, adve=L.as_t.pred_corr)
    oi.dev_count = slabs
    oi.n_sd_max = int(oi.n_sd_max * 2)
    for k in ("Cx", "Cy", "Cz"):
        f[k][:] = (1e-3 + np.arange(f[k].size).reshape(f[k].shape) * 1e-4) * {"Cx": 1.0, "Cy": 0.5, "Cz": 0.1}[k]
    o.cond = o.coal = o.sedi = 0
    p = b200.factory(L.backend_t.multi_CUDA, oi)
    p.init(f["th"], f["rv"], f["rhod"], None, f["Cx"], f["Cy"], f["Cz"])
    lib = E.lib()
    lib.lcx_cells_get.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64]
    lib.lcx_field_size.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int64)]
    share = nx // slabs                         # distmem_opts.hpp:10-18: int / int, the .5 never rounds up
    for d in range(slabs):
        eng = D.engine_of(b200, p, d)
        nx_d = share if d < slabs - 1 else nx - d * share
        x_bfr = d * share
        for field, name, ext in ((4, "Cx", (1, 0, 0)), (5, "Cy", (0, 1, 0)), (6, "Cz", (0, 0, 1))):
            cnt = C.c_int64()
            E.check(lib.lcx_field_size(eng.h, field, C.byref(cnt)))
            got = np.empty(cnt.value)
            E.check(lib.lcx_cells_get(eng.h, field, got.ctypes.data, cnt.value))
            g = f[name]
            cols = nx_d + 2 * halo + ext[0]
            assert cnt.value == cols * (ny + ext[1]) * (nz + ext[2]), name
            # column q of the slab's array is global column x_bfr - halo + q, wrapped over the global array's own extent
            want = g[(x_bfr - halo + np.arange(cols)) % g.shape[0]]
            assert np.array_equal(got.reshape(want.shape), want), (d, name)
    n0 = sum(D.engine_of(b200, p, d).n_part() for d in range(slabs))
    for _ in range(4):
        p.step_sync(o, f["th"], f["rv"], f["rhod"], f["Cx"], f["Cy"], f["Cz"])
        p.step_async(o)
    n1 = sum(D.engine_of(b200, p, d).n_part() for d in range(slabs))
    assert 0.98 * n0 <= n1 <= n0          # the test field has a small upward component: a few SDs may leave through the lid


@pytest.mark.parametrize("ranks", [2, 3])
def test_torchrun_ranks_pred_corr_courant_halo(ranks):
    """process-distributed predictor-corrector advection: the two Courant halo planes per side come from the neighbour ranks each
    step (particles_impl_xchng_courants.ipp:15-153), delivered over peer memory; device arrays checked value by value"""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(ranks), "--master-addr", "127.0.0.1",
           "--master-port", str(29521 + ranks), os.path.join(ROOT, "tests", "dist_worker.py"), "--halo"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0 and "DIST_HALO_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


def test_distmem_setting_is_consumed_by_one_particle_system(b200):
    """lgrngn_b200_set_distmem applies to the next factory() only: the particle system after it is an ordinary single-device one"""
    from libcloudphxx_b200 import distributed as D
    D.configure(b200, 0, 2, lft_x1=80.0, rgt_x0=0.0, n_x_tot=8)
    oi, o, f = S.box_3d(b200, nx=4, ny=3, nz=4, sd_conc=8)
    first = b200.factory(L.backend_t.CUDA, oi)
    assert first is not None
    oi2, o2, f2 = S.box_3d(b200, nx=4, ny=3, nz=4, sd_conc=8)
    p = b200.factory(L.backend_t.CUDA, oi2)
    p.init(f2["th"], f2["rv"], f2["rhod"], None, f2["Cx"], f2["Cy"], f2["Cz"])
    p.step_sync(o2, f2["th"], f2["rv"], f2["rhod"], f2["Cx"], f2["Cy"], f2["Cz"])
    p.step_async(o2)
    assert p.get_n().size > 0
