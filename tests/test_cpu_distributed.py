"""world_size-2 (and 3) gloo tests of the host-side plumbing of process-distributed runs (libcloudphxx_b200/distributed.py).

Since round 2 the migration itself lives in the library (device kernels writing into the neighbours' inboxes, sequence numbers
in device memory: include/lcx_b200.h); what Python still does is the one-time rendezvous - every rank publishes the handles of
its two inboxes and picks up those of its ring neighbours.  Tested here with stand-in blobs: each rank must end up with exactly
its left and right neighbours' pairs (with two ranks both neighbours are the same rank), whatever the world size."""
import os
import socket
import sys

import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def fake_pair(rank):
    from libcloudphxx_b200.distributed import BLOB_BYTES
    return bytes([rank]) * BLOB_BYTES + bytes([100 + rank]) * BLOB_BYTES       # inbox 0 | inbox 1, contents include zero bytes for rank 0


def worker(rank, size, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=size)
    from libcloudphxx_b200.distributed import ring_exchange, PAIR_BYTES
    for _ in range(2):
        lft, rgt = ring_exchange(fake_pair(rank), rank, size)
        assert len(lft) == PAIR_BYTES and len(rgt) == PAIR_BYTES
        assert lft == fake_pair((rank - 1) % size), "left neighbour's handles"
        assert rgt == fake_pair((rank + 1) % size), "right neighbour's handles"
    dist.barrier()
    dist.destroy_process_group()


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("size", [2, 3])
def test_inbox_handles_reach_the_ring_neighbours_gloo(size):
    mp.spawn(worker, args=(size, free_port()), nprocs=size, join=True)


def test_slab_split_matches_reference_rule():
    """x-columns per device: int(nx / G + .5) for all but the last, remainder to the last (src/detail/distmem_opts.hpp:10-18);
    nx and G are ints there, so nx / G is an INTEGER division and the .5 never rounds up"""
    def dev_nx(nx, rank, size):
        return int(nx // size + .5) if rank < size - 1 else nx - rank * int(nx // size + .5)
    for nx, size in ((512, 8), (5, 2), (7, 3), (10, 4)):
        parts = [dev_nx(nx, r, size) for r in range(size)]
        assert sum(parts) == nx and all(p > 0 for p in parts), (nx, size, parts)
    assert [dev_nx(5, r, 2) for r in range(2)] == [2, 3]
    assert [dev_nx(7, r, 3) for r in range(3)] == [2, 2, 3]
    assert [dev_nx(9, r, 2) for r in range(2)] == [4, 5]


def halo_worker(rank, size, port):
    """every rank owns a piece of a global periodic Courant field, embeds it in the halo-extended arrays of a slab, exchanges the planes
    the reference's rule names (oracle.sdm_port.xchng_courants_rule <- xchng_courants.ipp:26-140) with its ring neighbours over gloo, and
    must then hold - halo included - exactly the neighbours' columns.  The `want` below is the statement the GPU test uses for the device
    arrays (tests/dist_worker.py --halo): this test ties it to the reference's index arithmetic."""
    import numpy as np
    import torch
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=size)
    from oracle.sdm_port import xchng_courants_rule
    halo, ny, nz = 2, 3, 5
    nxs = [4 + r for r in range(size)]
    nx, x_bfr, n_x_tot = nxs[rank], sum(nxs[:rank]), sum(nxs)
    shapes = {"Cx": (n_x_tot + 1, ny, nz), "Cy": (n_x_tot, ny + 1, nz), "Cz": (n_x_tot, ny, nz + 1)}
    G = {k: np.arange(int(np.prod(sh)), dtype=np.float64).reshape(sh) + {"Cx": 0.25, "Cy": 0.5, "Cz": 0.75}[k] for k, sh in shapes.items()}
    rule = xchng_courants_rule(3, nx, ny, nz, halo)
    lft, rgt = (rank - 1) % size, (rank + 1) % size
    for name in ("Cx", "Cy", "Cz"):
        plane, s_lft, s_rgt, r_lft, r_rgt, count = rule[name]
        own = nx + (1 if name == "Cx" else 0)
        ext = np.full((own + 2 * halo) * plane, np.nan)
        ext[halo * plane:(halo + own) * plane] = G[name][x_bfr:x_bfr + own].ravel()          # the rank's own piece, halo unset
        to_lft, to_rgt = torch.from_numpy(ext[s_lft:s_lft + count].copy()), torch.from_numpy(ext[s_rgt:s_rgt + count].copy())
        from_lft, from_rgt = torch.empty(count, dtype=torch.float64), torch.empty(count, dtype=torch.float64)
        ops = [dist.P2POp(dist.isend, to_lft, lft, tag=1), dist.P2POp(dist.isend, to_rgt, rgt, tag=2),
               dist.P2POp(dist.irecv, from_rgt, rgt, tag=1), dist.P2POp(dist.irecv, from_lft, lft, tag=2)]
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        ext[r_lft:r_lft + count] = from_lft.numpy()
        ext[r_rgt:r_rgt + count] = from_rgt.numpy()
        first_rgt = 1 if name == "Cx" else 0
        cols = np.concatenate([(x_bfr - halo + np.arange(halo)) % n_x_tot, x_bfr + np.arange(own), (x_bfr + nx + first_rgt + np.arange(halo)) % n_x_tot])
        want = G[name][cols]
        assert np.array_equal(ext.reshape(want.shape), want), "rank %d: %s" % (rank, name)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("size", [2, 3])
def test_courant_halo_rule_over_gloo(size):
    mp.spawn(halo_worker, args=(size, free_port()), nprocs=size, join=True)
