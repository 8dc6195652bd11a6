"""torchrun worker of tests/test_gpu_multi.py: one rank per GPU, each owning an x-slab; checks the invariants the
reference's own distributed test asserts (tests/mpi/mpi_adve_test.cpp:143-256): after advecting once round the periodic
domain with C = +1 every per-cell statistic is back where it started, and nothing is lost or duplicated on the way."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from libcloudphxx_b200 import distributed as D, lgrngn as L   # noqa: E402
from tests import support as S                                # noqa: E402


def full_microphysics(lib, rank, world, local):
    """cfg5-shaped: Cx = 0.5 (half a column of super-droplets crosses every slab face each step, most SDs change cell, so the
    re-layout takes its full-sort path), a rain mode that falls out, condensation + coalescence under the Philox stream:
    the dry volume of all ranks + what left the domain is conserved, every rank keeps its SDs in step"""
    from libcloudphxx_b200 import engine as E
    nx, ny, nz, steps = [int(v) for v in os.environ.get("LCX_DIST_FULL_SHAPE", "6,8,10,24").split(",")]
    lib.lib.lgrngn_b200_set_rng_mode.argtypes = [C.c_int]
    lib.lib.lgrngn_b200_set_rng_mode(0)
    D.configure(lib, rank, world, lft_x1=nx * 20.0, rgt_x0=0.0, n_x_tot=nx * world)
    oi, o, f = S.box_3d(lib, nx=nx, ny=ny, nz=nz, sd_conc=40, rain_mode=True, cx=0.5, n_sd_max=int(nx * ny * nz * 40 * 1.25))
    oi.rng_seed = 44 + rank
    oi.dev_id = local
    p = lib.factory(L.backend_t.CUDA, oi)
    p.init(f["th"], f["rv"], f["rhod"], None, f["Cx"], f["Cy"], f["Cz"])
    D.connect(lib, p, rank, world)
    eng = D.engine_of(lib, p)

    def volume():
        p.diag_all(); p.diag_dry_mom(3)
        live = float((p.outbuf().reshape(nx, ny, nz) * f["rhod"]).sum() * 20.0 ** 3 * 4.0 / 3.0 * np.pi)
        tot = [None] * world
        dist.all_gather_object(tot, (live, p.diag_puddle()["dry_volume"], eng.top_loss()[0], eng.n_part()))
        return tot
    v0 = volume()
    total0 = sum(t[0] + t[1] + t[2] for t in v0)
    sent = []
    import time
    for step in range(steps):
        try:
            p.step_sync(o, f["th"], f["rv"], f["rhod"], f["Cx"], f["Cy"], f["Cz"])
            p.step_async(o)
        except Exception as ex:
            print("rank %d failed at step %d: %s" % (rank, step, ex), file=sys.stderr, flush=True)
            time.sleep(8)            # lets the other ranks report their own view before the launcher tears everything down
            raise
        sent.append(D.migr_stats(lib, p))
    v1 = volume()
    total1 = sum(t[0] + t[1] + t[2] for t in v1)
    assert abs(total1 - total0) <= 1e-10 * total0, (total0, total1)
    assert sum(t[1] for t in v1) > 0, "nothing rained out"
    moved = [None] * world
    dist.all_gather_object(moved, sent)
    for step in range(steps):      # what every rank sent right is what its right neighbour received from the left, and vice versa
        for r in range(world):
            assert moved[r][step][1] == moved[(r + 1) % world][step][3], (step, r)
            assert moved[r][step][0] == moved[(r - 1) % world][step][2], (step, r)
    assert sum(m[1] for m in sent) > steps * 0.3 * ny * nz * 40, "hardly anything migrated"
    if rank == 0:
        print("DIST_FULL_OK world=%d sd=%s" % (world, [t[3] for t in v1]))


def pred_corr_halo(lib, rank, world, local):
    """predictor-corrector advection between process-distributed slabs: every rank passes its OWN piece of the Courant fields, the
    two halo planes per side must arrive from the neighbours (the reference's MPI build: particles_impl_xchng_courants.ipp:26-140 -
    to the left neighbour go Cx faces 1, 2 and Cy / Cz columns 0, 1, to the right neighbour the last two faces / columns).  Checked on
    the device arrays themselves against a numpy statement of that rule on a global field whose every value is unique."""
    from libcloudphxx_b200 import engine as E
    halo, ny, nz, dx = 2, 3, 5, 20.0
    nxs = [4 + r for r in range(world)]                       # unequal slabs
    nx, x_bfr, n_x_tot = nxs[rank], sum(nxs[:rank]), sum(nxs)
    D.configure(lib, rank, world, lft_x1=nxs[(rank - 1) % world] * dx, rgt_x0=0.0, n_x_tot=n_x_tot)
    oi, o, f = S.box_3d(lib, nx=nx, ny=ny, nz=nz, sd_conc=16, adve=L.as_t.pred_corr)
    oi.n_sd_max = int(oi.n_sd_max * 2)
    oi.rng_seed = 99 + rank
    oi.dev_id = local
    shapes = {"Cx": (n_x_tot + 1, ny, nz), "Cy": (n_x_tot, ny + 1, nz), "Cz": (n_x_tot, ny, nz + 1)}
    G = {k: (1e-3 + np.arange(int(np.prod(sh))).reshape(sh) * 1e-4) * {"Cx": 1.0, "Cy": 0.5, "Cz": 0.1}[k] for k, sh in shapes.items()}
    for k in G:
        f[k] = np.ascontiguousarray(G[k][x_bfr:x_bfr + nx + (1 if k == "Cx" else 0)])
    o.cond = o.coal = o.sedi = 0
    p = lib.factory(L.backend_t.CUDA, oi)
    p.init(f["th"], f["rv"], f["rhod"], None, f["Cx"], f["Cy"], f["Cz"])
    D.connect(lib, p, rank, world)
    eng = D.engine_of(lib, p)
    n0 = [None] * world
    dist.all_gather_object(n0, eng.n_part())
    lcx = E.lib()
    lcx.lcx_cells_get.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64]
    lcx.lcx_field_size.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int64)]
    for step in range(4):
        f["Cx"] *= 1.0 + 0.01 * step                          # the fields change every step: a stale halo would show
        G["Cx"] = G["Cx"] * (1.0 + 0.01 * step)
        p.step_sync(o, f["th"], f["rv"], f["rhod"], f["Cx"], f["Cy"], f["Cz"])
        p.step_async(o)
        for field, name in ((4, "Cx"), (5, "Cy"), (6, "Cz")):
            cnt = C.c_int64()
            E.check(lcx.lcx_field_size(eng.h, field, C.byref(cnt)))
            got = np.empty(cnt.value)
            E.check(lcx.lcx_cells_get(eng.h, field, got.ctypes.data, cnt.value))
            own = nx + (1 if name == "Cx" else 0)
            first_rgt = 1 if name == "Cx" else 0               # the right neighbour's faces 1, 2 / columns 0, 1
            cols = np.concatenate([(x_bfr - halo + np.arange(halo)) % n_x_tot, x_bfr + np.arange(own),
                                   (x_bfr + nx + first_rgt + np.arange(halo)) % n_x_tot])
            want = G[name][cols]
            assert got.size == want.size, (name, got.size, want.shape)
            assert np.array_equal(got.reshape(want.shape), want), "rank %d step %d: %s halo differs" % (rank, step, name)
    n1 = [None] * world
    dist.all_gather_object(n1, eng.n_part())
    assert 0.98 * sum(n0) <= sum(n1) <= sum(n0), (n0, n1)      # the field has a small upward component: a few SDs may leave through the lid
    if rank == 0:
        print("DIST_HALO_OK world=%d sd=%s" % (world, n1))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    n_dev = torch.cuda.device_count()
    local = local % n_dev                       # fewer GPUs than ranks (the 1-GPU test box): ranks share devices, the inboxes still
    torch.cuda.set_device(local)                # travel through CUDA IPC; NCCL refuses two ranks per GPU, so the rendezvous uses gloo
    if n_dev >= world:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist.init_process_group("gloo")
    lib = L.b200()
    if "--halo" in sys.argv:
        pred_corr_halo(lib, rank, world, local)
        dist.barrier()
        dist.destroy_process_group()
        return
    if "--full" in sys.argv:
        full_microphysics(lib, rank, world, local)
        dist.barrier()
        dist.destroy_process_group()
        return
    nx_of = lambda r: r + 2                     # unequal slabs, like mpi_adve_test.cpp:88
    nx, ny, nz = nx_of(rank), 3, 4
    n_x_tot = sum(nx_of(r) for r in range(world))
    dx = 20.0
    D.configure(lib, rank, world, lft_x1=nx_of((rank - 1) % world) * dx, rgt_x0=0.0, n_x_tot=n_x_tot)
    oi, o, f = S.box_3d(lib, nx=nx, ny=ny, nz=nz, sd_conc=16, rain_mode=True)
    oi.rng_seed = 4444 + rank
    oi.dev_id = local
    f["Cx"][:] = 1.0
    f["Cy"][:] = 0.0
    o.cond = o.coal = o.sedi = 0
    p = lib.factory(L.backend_t.CUDA, oi)
    p.init(f["th"], f["rv"], f["rhod"], None, f["Cx"], f["Cy"], f["Cz"])
    D.connect(lib, p, rank, world)              # one-time exchange of the inbox handles; step_async migrates from now on

    def stats():
        out = []
        for sel, mom in ((p.diag_all, p.diag_sd_conc), (p.diag_all, lambda: p.diag_dry_mom(1)), (p.diag_all, lambda: p.diag_wet_mom(1)),
                         (p.diag_all, lambda: p.diag_kappa_mom(1)), (p.diag_all, lambda: p.diag_dry_mom(0))):
            sel(); mom(); out.append(p.outbuf().reshape(nx, ny, nz).copy())
        return out

    def gather_global(a):
        parts = [None] * world
        dist.all_gather_object(parts, a)
        return np.concatenate(parts, axis=0)

    before = [gather_global(a) for a in stats()]
    n_sd_before = before[0].sum()
    for step in range(n_x_tot):
        p.step_sync(o, f["th"], f["rv"], f["rhod"], f["Cx"], f["Cy"], f["Cz"])
        p.step_async(o)
        now = [gather_global(a) for a in stats()]
        assert now[0].sum() == n_sd_before, "super-droplets lost or duplicated at step %d" % step
        # one step with C = +1 rolls every per-cell field by exactly one cell in x
        for a, b in zip(before, now):
            assert np.array_equal(np.roll(a, step + 1, axis=0), b), "step %d: field is not a pure roll" % step
    after = [gather_global(a) for a in stats()]
    for a, b in zip(before, after):
        assert np.array_equal(a, b)
    if rank == 0:
        print("DIST_OK world=%d n_sd=%d migrants_last_step=%s" % (world, int(n_sd_before), D.migr_stats(lib, p)))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
