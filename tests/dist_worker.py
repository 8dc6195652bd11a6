"""torchrun worker of tests/test_gpu_multi.py: one rank per GPU, each owning an x-slab; checks the invariants the
reference's own distributed test asserts (tests/mpi/mpi_adve_test.cpp:143-256): after advecting once round the periodic
domain with C = +1 every per-cell statistic is back where it started, and nothing is lost or duplicated on the way."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from libcloudphxx_b200 import distributed as D, lgrngn as L   # noqa: E402
from tests import support as S                                # noqa: E402


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
