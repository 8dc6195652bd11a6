"""One process per GPU: x-slab decomposition of the Lagrangian domain with super-droplet migration over NCCL.

The reference's `multi_CUDA` back-end keeps all devices in one process and moves migrants with cudaMemcpyPeerAsync
between five host barriers (reference src/impl_multi_gpu/particles_multi_gpu_impl_step_async_and_copy.ipp:28-206); its
MPI mode gives each process a rank-local opts_init (tests/mpi/mpi_adve_test.cpp:88-141).  Here each torchrun rank owns one
slab (rank-local opts_init and rank-local Eulerian arrays), and after the local part of step_async the packed migrant
buffers - which live in the engine's device memory - are exchanged with the two ring neighbours by grouped NCCL
send/recv on zero-copy torch views of those buffers.  No all-reduce anywhere: the exchange is neighbour-only.

Order of arrival follows the reference: first the right neighbour's left-movers, then the left neighbour's
right-movers, each in the sender's storage order.
"""
import ctypes as C

import numpy as np

from . import engine as E
from . import lgrngn as L


class _CudaView:
    """exposes a raw device pointer through __cuda_array_interface__ so torch can wrap it without a copy"""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def configure(lib, rank, size, lft_x1, rgt_x0, n_x_tot):
    """must be called before factory(): the next CUDA particle system becomes slab `rank` of `size`"""
    class D(C.Structure):
        _fields_ = [("rank", C.c_int), ("size", C.c_int), ("lft_x1", C.c_double), ("rgt_x0", C.c_double), ("n_x_tot", C.c_int)]
    lib.lib.lgrngn_b200_set_distmem.argtypes = [C.POINTER(D)]
    d = D(rank, size, lft_x1, rgt_x0, n_x_tot)
    lib.lib.lgrngn_b200_set_distmem(C.byref(d))


def engine_of(lib, prt):
    lib.lib.lgc_proto.restype = C.c_void_p
    lib.lib.lgc_proto.argtypes = [C.c_void_p]
    lib.lib.lgrngn_b200_engine.restype = C.c_void_p
    lib.lib.lgrngn_b200_engine.argtypes = [C.c_void_p]
    return E.Engine(lib.lib.lgrngn_b200_engine(lib.lib.lgc_proto(prt._h)))


class EngineSlab:
    """adapter between SlabExchange and a real particle system: packed migrants live in the engine's device memory"""

    device = "cuda"

    def __init__(self, lib, prt):
        import torch
        self.torch, self.lib, self.prt = torch, lib, prt
        self.eng = engine_of(lib, prt)
        self.n_real = self.eng.migr_real_attrs()
        lib.lib.lgrngn_b200_post_copy.argtypes = [C.c_void_p, C.c_int]
        lib.lib.lgc_proto.restype = C.c_void_p

    def pack(self):
        return self.eng.migr_pack()          # synchronises the engine's stream: the outgoing buffers are complete

    def tensors(self, side, incoming, count):
        torch = self.torch
        n_ptr, r_ptr, cap = self.eng.migr_buffers(side, incoming)
        if count > cap:
            raise RuntimeError("migration buffer overflow: %d > %d" % (count, cap))
        n = torch.as_tensor(_CudaView(n_ptr, max(count, 1), "<i8"), device="cuda")[:count]
        r = torch.as_tensor(_CudaView(r_ptr, max(count * self.n_real, 1), "<f8"), device="cuda")[:count * self.n_real]
        return n, r

    def received(self):
        self.torch.cuda.current_stream().synchronize()

    def unpack(self, side, count):
        self.eng.migr_unpack(side, count)

    def post_copy(self, rcyc):
        lib = self.lib.lib
        if lib.lgrngn_b200_post_copy(lib.lgc_proto(self.prt._h), int(rcyc)) != 0:
            raise RuntimeError("post_copy failed")


class SlabExchange:
    """finishes step_async of a process-distributed slab: pack -> neighbour exchange -> unpack -> post_copy.
    `slab` is an EngineSlab (GPU, NCCL) or any object with the same five methods (the CPU/gloo tests use a stand-in)."""

    def __init__(self, slab, rank, size):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.slab, self.rank, self.size = slab, rank, size
        self.lft = (rank - 1) % size
        self.rgt = (rank + 1) % size

    def finish_step(self, adve=True, rcyc=False):
        torch, dist, slab = self.torch, self.dist, self.slab
        if adve and self.size > 1:
            n_lft, n_rgt = slab.pack()
            # how many arrive: my right neighbour's left-movers and my left neighbour's right-movers
            send = torch.tensor([n_lft, n_rgt], dtype=torch.int64, device=slab.device)
            from_rgt = torch.zeros(1, dtype=torch.int64, device=slab.device)
            from_lft = torch.zeros(1, dtype=torch.int64, device=slab.device)
            ops = [dist.P2POp(dist.isend, send[0:1], self.lft), dist.P2POp(dist.isend, send[1:2], self.rgt),
                   dist.P2POp(dist.irecv, from_rgt, self.rgt), dist.P2POp(dist.irecv, from_lft, self.lft)]
            for w in dist.batch_isend_irecv(ops):
                w.wait()
            n_from_rgt, n_from_lft = int(from_rgt.item()), int(from_lft.item())
            ops = []
            if n_lft:
                n, r = slab.tensors(0, False, n_lft)
                ops += [dist.P2POp(dist.isend, n, self.lft), dist.P2POp(dist.isend, r, self.lft)]
            if n_rgt:
                n, r = slab.tensors(1, False, n_rgt)
                ops += [dist.P2POp(dist.isend, n, self.rgt), dist.P2POp(dist.isend, r, self.rgt)]
            if n_from_rgt:
                n, r = slab.tensors(0, True, n_from_rgt)
                ops += [dist.P2POp(dist.irecv, n, self.rgt), dist.P2POp(dist.irecv, r, self.rgt)]
            if n_from_lft:
                n, r = slab.tensors(1, True, n_from_lft)
                ops += [dist.P2POp(dist.irecv, n, self.lft), dist.P2POp(dist.irecv, r, self.lft)]
            if ops:
                for w in dist.batch_isend_irecv(ops):
                    w.wait()
                slab.received()
            slab.unpack(0, n_from_rgt)      # arrivals from the right neighbour first, as the reference appends them
            slab.unpack(1, n_from_lft)
        slab.post_copy(rcyc)
