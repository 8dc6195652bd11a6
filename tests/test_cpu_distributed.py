"""world_size-2 (and 3) gloo tests of the neighbour-exchange protocol used when one process drives one GPU.
The device engine is replaced by a stand-in that keeps its migrants in CPU tensors; the ordering contract is the
reference's (arrivals from the right neighbour first, then from the left, each in the sender's order:
src/impl_multi_gpu/particles_multi_gpu_impl_step_async_and_copy.ipp:104,134,161,190)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

N_REAL = 3


class FakeSlab:
    """a slab whose 'super-droplets' are (id, payload) pairs; ids encode origin rank and direction"""
    device = "cpu"

    def __init__(self, rank, size, step):
        self.rank, self.size = rank, size
        self.n_lft, self.n_rgt = (rank + step) % 3, (2 * rank + 1 + step) % 4      # includes zero-sized messages
        mk = lambda side, cnt: (torch.arange(cnt, dtype=torch.int64) + 1000 * rank + 100 * side,
                                torch.arange(cnt * N_REAL, dtype=torch.float64) + 0.5 * rank + 0.25 * side)
        self.out = {0: mk(0, self.n_lft), 1: mk(1, self.n_rgt)}
        self.inc = {0: (torch.zeros(16, dtype=torch.int64), torch.zeros(16 * N_REAL, dtype=torch.float64)),
                    1: (torch.zeros(16, dtype=torch.int64), torch.zeros(16 * N_REAL, dtype=torch.float64))}
        self.appended, self.finished = [], False

    def pack(self):
        return self.n_lft, self.n_rgt

    def tensors(self, side, incoming, count):
        n, r = self.inc[side] if incoming else self.out[side]
        return n[:count], r[:count * N_REAL]

    def received(self):
        pass

    def unpack(self, side, count):
        n, r = self.inc[side]
        self.appended.append((side, n[:count].clone(), r[:count * N_REAL].clone()))

    def post_copy(self, rcyc):
        self.finished = True


def worker(rank, size, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=size)
    from libcloudphxx_b200.distributed import SlabExchange
    for step in range(3):
        slab = FakeSlab(rank, size, step)
        SlabExchange(slab, rank, size).finish_step()
        assert slab.finished
        lft, rgt = (rank - 1) % size, (rank + 1) % size
        exp_r, exp_l = FakeSlab(rgt, size, step), FakeSlab(lft, size, step)
        (s0, n0, r0), (s1, n1, r1) = slab.appended
        assert (s0, s1) == (0, 1)                                  # right neighbour's batch is appended first
        assert torch.equal(n0, exp_r.out[0][0]) and torch.equal(r0, exp_r.out[0][1])    # its left-movers
        assert torch.equal(n1, exp_l.out[1][0]) and torch.equal(r1, exp_l.out[1][1])    # left neighbour's right-movers
    dist.barrier()
    dist.destroy_process_group()


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("size", [2, 3])
def test_neighbour_exchange_gloo(size):
    mp.spawn(worker, args=(size, free_port()), nprocs=size, join=True)


def test_slab_split_matches_reference_rule():
    """x-columns per device: int(nx/G + .5) for all but the last, remainder to the last (src/detail/distmem_opts.hpp:10-18)"""
    def dev_nx(nx, rank, size):
        return int(nx / size + .5) if rank < size - 1 else nx - rank * int(nx / size + .5)
    for nx, size in ((512, 8), (5, 2), (7, 3), (10, 4)):
        parts = [dev_nx(nx, r, size) for r in range(size)]
        assert sum(parts) == nx and all(p > 0 for p in parts), (nx, size, parts)
    assert [dev_nx(5, r, 2) for r in range(2)] == [3, 2]
    assert [dev_nx(7, r, 3) for r in range(3)] == [2, 2, 3]
