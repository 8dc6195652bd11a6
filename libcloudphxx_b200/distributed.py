"""One process per GPU: set-up plumbing for x-slab runs spread over several processes (torchrun).

The reference's MPI mode gives each process a rank-local opts_init (tests/mpi/mpi_adve_test.cpp:88-141) and migrates
super-droplets inside step_async (src/particles_step.ipp:484-490).  The same here: each rank owns one slab with rank-local
Eulerian arrays, and `step_async` of the library itself packs the leavers straight into the ring neighbours' inboxes - device
memory opened through CUDA IPC, written over NVLink, ordered by sequence numbers in device memory (include/lcx_b200.h,
"x-slab migration").  Nothing in the time step goes through Python, NCCL or the host.

What is left for this module is the ONE-TIME rendezvous: every rank exports two 96-byte blobs (the IPC handles of its two
inboxes), the blobs travel to the ring neighbours with a single `all_gather` (nccl or gloo), and each rank connects.  A C++
host model does the same with whatever transport it has (MPI_Allgather of 192 bytes): host/particles_b200.h
`lgrngn_b200_distmem_export` / `lgrngn_b200_distmem_connect`.
"""
import ctypes as C

from . import engine as E

BLOB_BYTES = 96          # LCX_IPC_BLOB_BYTES
PAIR_BYTES = 2 * BLOB_BYTES


def configure(lib, rank, size, lft_x1, rgt_x0, n_x_tot):
    """must be called right before factory(): the NEXT CUDA particle system becomes slab `rank` of `size` (the setting is
    consumed by that constructor)"""
    class D(C.Structure):
        _fields_ = [("rank", C.c_int), ("size", C.c_int), ("lft_x1", C.c_double), ("rgt_x0", C.c_double), ("n_x_tot", C.c_int)]
    lib.lib.lgrngn_b200_set_distmem.argtypes = [C.POINTER(D)]
    d = D(rank, size, lft_x1, rgt_x0, n_x_tot)
    lib.lib.lgrngn_b200_set_distmem(C.byref(d))


def proto_of(lib, prt):
    lib.lib.lgc_proto.restype = C.c_void_p
    lib.lib.lgc_proto.argtypes = [C.c_void_p]
    return C.c_void_p(lib.lib.lgc_proto(prt._h))


def _sfx(lib):
    return "_f32" if getattr(lib, "real", "f64") == "f32" else ""


def engine_of(lib, prt, slab=0):
    f = getattr(lib.lib, "lgrngn_b200_engine_of_slab" + _sfx(lib))
    f.restype = C.c_void_p
    f.argtypes = [C.c_void_p, C.c_int]
    return E.Engine(f(proto_of(lib, prt), slab), getattr(lib, "real", "f64"))


def n_slabs(lib, prt):
    f = getattr(lib.lib, "lgrngn_b200_n_slabs" + _sfx(lib))
    f.argtypes = [C.c_void_p]
    return int(f(proto_of(lib, prt)))


def migr_stats(lib, prt):
    """migrants of the last step summed over this process's slabs: (sent left, sent right, received from right, received from left)"""
    out = (C.c_longlong * 4)()
    lib.lib.lgrngn_b200_migr_stats.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
    if lib.lib.lgrngn_b200_migr_stats(proto_of(lib, prt), out) != 0:
        raise RuntimeError("migr_stats failed")
    return tuple(int(v) for v in out)


def export_blobs(lib, prt):
    buf = (C.c_ubyte * PAIR_BYTES)()
    lib.lib.lgrngn_b200_distmem_export.argtypes = [C.c_void_p, C.c_void_p]
    if lib.lib.lgrngn_b200_distmem_export(proto_of(lib, prt), buf) != 0:
        raise RuntimeError("exporting the migration inboxes failed")
    return bytes(buf)


def ring_exchange(my_pair, rank, size):
    """all_gather of every rank's blob pair; returns (left neighbour's pair, right neighbour's pair) as bytes"""
    import torch.distributed as dist
    assert len(my_pair) == PAIR_BYTES
    parts = [None] * size
    dist.all_gather_object(parts, bytes(my_pair))
    return parts[(rank - 1) % size], parts[(rank + 1) % size]


def connect(lib, prt, rank, size):
    """after init(): exchanges the inbox handles with the ring neighbours and connects; from then on step_async migrates"""
    lft, rgt = ring_exchange(export_blobs(lib, prt), rank, size)
    lib.lib.lgrngn_b200_distmem_connect.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
    if lib.lib.lgrngn_b200_distmem_connect(proto_of(lib, prt), lft, rgt) != 0:
        raise RuntimeError("connecting the neighbours' inboxes failed")
