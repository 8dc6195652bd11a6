#!/usr/bin/env python
"""Headline benchmark: super-droplet updates per second of the full cond + coal + sedi + adve step.

Workload (BASELINE.json configs[3], one x-slab of it per GPU = weak scaling): 3-D LES-like box, 64 x 256 x 128 cells
per GPU (512 x 256 x 128 on 8 GPUs), 40 super-droplets per cell (8.4e7 per GPU), Hall/Davis coalescence kernel,
beard77fast fall speeds, implicit advection with Cx = 0.1, Cy = 0.05, two-mode lognormal aerosol, supersaturated
upper half.  Double precision.  A "step" is one step_sync + step_async of every live super-droplet.

  value  - device-timed (CUDA events on the engine's stream, max over ranks) throughput with the Eulerian fields resident
           in HBM (lgrngn_b200_step_resident);
  e2e    - the same step through the reference-facing API (step_sync / step_async with HOST arrays: per step th, rv, rhod
           and three Courant fields go host->device from pinned memory, th and rv come back), wall clock incl. copies;
  roofline - dominant kernel from the live per-kernel CUDA-event profile; achieved = algorithmic bytes / mean duration;
  cpu_baseline - the reference's own OpenMP back-end (oracle/_ref, built from the unmodified reference sources) on a
           scaled-down box of the same shape, on this machine's host cores.

`--impl reference` times only that CPU reference arm and prints it in the same format.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

A_FULL_BYTES = 192.0          # algorithmic bytes per SD-update, full step, double
KERNEL_BYTES = {              # algorithmic bytes per SD per launch of the kernels that sweep all SDs
    "k_cond_cells": 48.0,      # rw2 r+w, rd3, kpa, vt, n (8 B each); the cell fields are 1/40 of that
    "k_cond_range": 52.0,      # the same + the cell index of every SD (4 B)
    "k_cond": 52.0, "k_coal_small": 76.0, "k_coal_big": 76.0,
    "k_transport": 72.0,       # x, y, z r+w (48), vt, n r+w (24), cell index (4), sort key out (4) - 8 with rain-out reads
    "k_gather": 136.0,
    "(k_cell_reduce_small<Term, IS_MAX>)": 20.0, "k_vterm": 24.0, "k_make_keys": 40.0,
    "k_mv_count": 4.0, "k_mv_list": 4.0, "k_mv_place_stayers": 12.0,
    # the radix sort and k_mv_place_arrivals / k_cell_offsets only see the SDs that changed cell: no per-SD figure
}
KERNEL_BYTES_EXACT = {        # gather-on-read variants (the default): the re-layout of n, rd3, rw2, kpa, vt rides in the condensation
    # kernel (permutation 4 + five attributes read 40 + cell index 4 + five attributes written 40), that of x, y, z in the
    # transport kernel (+ permutation 4), and k_gather is left with the storage index (permutation 4 + read 4 + write 4)
    "(k_cond_range<M, true>)": 88.0, "k_transport<true>": 76.0,
    "(k_cond_classed<M, true>)": 96.0, "(k_cond_classed<M, false>)": 60.0,      # class-ordered walk: + one more read of rw2 for the classification sweep
    "k_cond_staged<true>": 88.0, "k_cond_staged<false>": 52.0,      # the phase-grouped form of the range kernel: same traffic
}
LAZY = os.environ.get("LCX_LAZY_GATHER", "1") != "0"
# --real f32 (the single-precision engine, a SECOND mode - never the headline): the same tables with 4-byte reals.
# (fixed bytes, number of real words) per SD of the kernels that dominate; the multiplicity stays 8 bytes, indices and keys 4
REAL_BYTES = 8
A_FULL_BYTES_F32 = 136.0      # 36 read + 36 write state (n 8 + 7 x 4), 52 sort, 12 random inputs
KERNEL_WORDS = {"(k_cond_range<M, true>)": (24, 8), "k_cond_range": (20, 4), "(k_cond_classed<M, true>)": (24, 9), "k_cond_classed": (20, 5), "k_cond_cells": (16, 4), "k_cond": (20, 4),
                "k_transport<true>": (20, 7), "k_transport": (16, 7), "k_coal_small": (12, 8), "k_coal_big": (12, 8),
                "k_vterm": (0, 3), "k_mv_count": (4, 0), "k_mv_list": (4, 0), "k_mv_place_stayers": (12, 0)}


def base_name(name):
    """profile names carry template arguments and parentheses (e.g. "(k_cond_range<M>)"): the bare kernel name"""
    return name.strip("()").split("<")[0]


def kernel_bytes(name):
    if REAL_BYTES == 4:
        key = name if name in KERNEL_WORDS else ("k_vterm" if base_name(name).startswith("k_vterm") else base_name(name))
        fixed, words = KERNEL_WORDS.get(key, (0, 0))
        return float(fixed + 4 * words)
    if name in KERNEL_BYTES_EXACT:
        return KERNEL_BYTES_EXACT[name]
    if LAZY and base_name(name) == "k_gather":
        return 12.0
    for k, v in KERNEL_BYTES.items():
        if base_name(k) == base_name(name):
            return v
    if base_name(name).startswith("k_vterm"):
        return KERNEL_BYTES["k_vterm"]
    return 0.0


def traffic_of(name, n_sd):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed ncu --set full capture (profiles/):
    recorded per SD there because the capture ran on a smaller box; scaled to this launch's SD count"""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p) or REAL_BYTES == 4:      # the capture is of the double-precision engine
        return None
    table = json.load(open(p))["dram_bytes_per_sd"]
    hits = [k for k in table if k in name]
    return table[max(hits, key=len)] * n_sd if hits else None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons of one GPU, sampled WHILE the timed region runs.

    The timed region lasts a few hundred milliseconds, one `nvidia-smi` process start takes about as long, so the samples
    come from NVML in this process (the same counters nvidia-smi prints: clocks.sm, clocks.max.sm,
    clocks_event_reasons.*), every 2 ms; `nvidia-smi` is only the fallback when NVML cannot be loaded.
    """
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index=0, pci_bus_id=None):
        super().__init__(daemon=True)
        self.index, self.pci, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, pci_bus_id, [], set(), False, None
        self.source = None
        self.nvml = self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.handle = None
            if pci_bus_id:
                try:
                    self.handle = pynvml.nvmlDeviceGetHandleByPciBusId(pci_bus_id.encode() if isinstance(pci_bus_id, str) else pci_bus_id)
                except Exception:
                    self.handle = None
            if self.handle is None:
                vis = [v for v in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if v.strip().isdigit()]
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(int(vis[index]) if index < len(vis) else index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml, self.source = pynvml, "nvml"
        except Exception:
            self.nvml = self.handle = None

    def _sample_nvml(self):
        n = self.nvml
        self.samples.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
        try:
            mask = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        except Exception:
            mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        for bit, nm in self.REASONS.items():
            if mask & bit:
                self.reasons.add(nm)

    def _sample_smi(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip().split(",")
        self.samples.append(float(out[0]))
        self.max_mhz = float(out[1])
        self.source = "nvidia-smi"
        for nm, v in zip(names, out[2:]):
            if v.strip().lower().startswith("active"):
                self.reasons.add(nm)

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            time.sleep(0.002 if self.nvml else 0.05)

    def result(self):
        self.stop_flag = True
        self.join(timeout=6)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples), "source": self.source}


def pinned(shape, dtype=np.float64):
    import torch
    return torch.empty(shape, dtype=torch.float64 if dtype == np.float64 else torch.float32, pin_memory=True).numpy()


def make_case(lib, nx, ny, nz, sd_conc, n_sd_max_factor=1.25, pin=False, rank=0, size=1, config="cfg4"):
    """BASELINE.json configs[3] (cfg4: Cx = 0.1, one aerosol spectrum) or configs[4] (cfg5: Cx = 0.5, a second large mode that
    rains out - tests/mpi/mpi_adve_test.cpp:23-31 - so super-droplets drain through z0 and ~0.8 % of a 64-column slab cross
    a face every step); SURVEY.md section 8d"""
    from libcloudphxx_b200 import lgrngn as L
    from tests import support as S
    dtype = getattr(lib, "dtype", np.float64)
    alloc = (lambda s: pinned(s, dtype)) if pin else (lambda s: np.empty(s, dtype=dtype))
    oi = lib.opts_init_t()
    oi.nx, oi.ny, oi.nz = nx, ny, nz
    oi.dx = oi.dy = oi.dz = 20.0
    oi.x1, oi.y1, oi.z1 = nx * 20.0, ny * 20.0, nz * 20.0
    oi.dt = 1.0
    oi.sd_conc = sd_conc
    oi.n_sd_max = int(nx * ny * nz * sd_conc * n_sd_max_factor)
    oi.kernel = L.kernel_t.hall_davis_no_waals
    oi.terminal_velocity = L.vt_t.beard77fast
    oi.adve_scheme = L.as_t.implicit
    oi.rng_seed = 44 + rank
    distros = [L.lognormal(0.61, S.AEROSOL_ICICLE)]
    if config == "cfg5":
        distros.append(L.lognormal(1.28, [(30e-6, 1.2, 1e5)]))
    oi.dry_distros = distros
    th_dry, rhod_col, _ = S.hydrostatic_column(nz, 20.0)
    f = {"th": alloc((nx, ny, nz)), "rv": alloc((nx, ny, nz)), "rhod": alloc((nx, ny, nz)),
         "Cx": alloc((nx + 1, ny, nz)), "Cy": alloc((nx, ny + 1, nz)), "Cz": alloc((nx, ny, nz + 1))}
    f["th"][:] = th_dry
    f["rv"][:] = 6e-3
    f["rv"][:, :, nz // 2:] = 8.2e-3
    f["rhod"][:] = rhod_col
    f["Cx"][:] = 0.5 if config == "cfg5" else 0.1
    f["Cy"][:] = 0.05
    f["Cz"][:] = 0.0
    return oi, lib.opts_t(), f


def api_step(p, o, f):
    p.step_sync(o, f["th"], f["rv"], f["rhod"], f["Cx"], f["Cy"], f["Cz"])
    p.step_async(o)


def run_reference_cuda(args):
    """informative extra arm: the reference's OWN Thrust/CUDA back-end (src/lib_cuda.cu, unmodified, built for sm_100 by
    oracle/build_ref_cuda.py with the reference's release flags) on the same GPU and the same cfg4 slab as the product -
    the only like-for-like number; through the same API with host arrays, timed by wall clock"""
    import torch
    from libcloudphxx_b200 import lgrngn as L
    if int(os.environ.get("RANK", "0")) != 0:
        return
    path = os.path.join(ROOT, "oracle", "_ref", "liblgrngn_ref_cuda.so")
    if not os.path.exists(path):
        print(json.dumps({"impl": "reference-cuda", "unavailable": "oracle/_ref/liblgrngn_ref_cuda.so not built (oracle/build_ref_cuda.py, ~25 min)"}), flush=True)
        return
    ref = L.Library(path)
    nx, ny, nz = args.nx, args.ny, args.nz
    oi, o, f = make_case(ref, nx, ny, nz, args.sd_conc, pin=True, config=args.config)
    t0 = time.time()
    p = ref.factory(L.backend_t.CUDA, oi)
    p.init(f["th"], f["rv"], f["rhod"], None, f["Cx"], f["Cy"], f["Cz"])
    t_init = time.time() - t0
    n_sd = nx * ny * nz * args.sd_conc
    for _ in range(args.warmup):
        api_step(p, o, f)
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(args.steps):
        api_step(p, o, f)
    torch.cuda.synchronize()
    dt = time.time() - t0
    v = n_sd * args.steps / dt
    print(json.dumps({
        "impl": "reference-cuda", "metric": "super-droplet updates/s (cond+coal+sedi+adve step)", "value": v, "unit": "SD-updates/s",
        "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s x-slab: %dx%dx%d cells x %d SD/cell, hall_davis_no_waals, beard77fast, implicit adve; reference Thrust/CUDA back-end "
                               "(nvcc -O3 -use_fast_math, sm_100), cuRAND MTGP32, host arrays through step_sync/step_async" % (args.config, nx, ny, nz, args.sd_conc),
                   "init_s": round(t_init, 2)},
        "e2e": {"value": v, "unit": "SD-updates/s", "h2d_bytes_per_step": sum(f[k].nbytes for k in ("th", "rv", "rhod", "Cx", "Cy", "Cz")),
                "d2h_bytes_per_step": f["th"].nbytes + f["rv"].nbytes}}), flush=True)


def run_reference(args):
    """CPU reference arm: the reference's OpenMP back-end (unmodified sources, oracle/_ref) on the host cores"""
    from libcloudphxx_b200 import lgrngn as L
    from tests import support as S
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        os.environ["OMP_NUM_THREADS"] = str(cores)      # torchrun exports OMP_NUM_THREADS=1; only this rank works, it may use every core
    else:
        os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    cores = int(os.environ["OMP_NUM_THREADS"])           # the threads the OpenMP back-end will really use
    # timing uses the build with the reference's own release optimisation (-Ofast, oracle/build_ref.py) when it is there and loads;
    # the IEEE-strict -O2 build that the parity tests use otherwise
    fast = os.path.join(ROOT, "oracle", "_ref", "liblgrngn_ref_fast.so")
    ref, flags = None, "-O2"
    if os.path.exists(fast) and not args.ref_strict:
        probe = subprocess.run([sys.executable, "-c", "import ctypes; ctypes.CDLL(%r)" % fast], capture_output=True)
        if probe.returncode == 0:
            ref, flags = L.Library(fast), "-Ofast"
    if ref is None:
        ref = S.oracle_library()
    nx, ny, nz = args.ref_nx, args.ref_ny, args.ref_nz
    oi, o, f = make_case(ref, nx, ny, nz, args.sd_conc)
    p = ref.factory(L.backend_t.OpenMP, oi)
    p.init(f["th"], f["rv"], f["rhod"], None, f["Cx"], f["Cy"], f["Cz"])
    n_sd = nx * ny * nz * args.sd_conc
    for _ in range(args.warmup):
        api_step(p, o, f)
    t0 = time.time()
    for _ in range(args.steps):
        api_step(p, o, f)
    dt = time.time() - t0
    v = n_sd * args.steps / dt
    line = {
        "impl": "reference", "metric": "super-droplet updates/s (cond+coal+sedi+adve step)", "value": v, "unit": "SD-updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "cfg4-shaped 3-D box %dx%dx%d cells x %d SD/cell (bounded sample of the 64x256x128 slab), hall_davis_no_waals, beard77fast, implicit adve" % (nx, ny, nz, args.sd_conc)},
        "cpu_baseline": {"value": v, "unit": "SD-updates/s", "cores": cores, "kind": "reference",
                         "sample": "%dx%dx%d cells x %d SD/cell = %.3g SDs, %d steps, reference OpenMP back-end (%s)" % (nx, ny, nz, args.sd_conc, n_sd, args.steps, flags)},
        "e2e": {"value": v, "unit": "SD-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def pin_to_gpu_numa_node(local):
    """host threads and the pinned staging memory of this rank go to the CPU socket its GPU hangs off (first touch): with eight
    ranks streaming 135 MB per step each, crossing the socket interconnect is what limits the end-to-end figure"""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = [v for v in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if v.strip().isdigit()]
        h = pynvml.nvmlDeviceGetHandleByIndex(int(vis[local]) if local < len(vis) else local)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return len(os.sched_getaffinity(0))
    except Exception:
        return None


def run_b200(args):
    import torch
    from libcloudphxx_b200 import lgrngn as L, distributed as D, engine as E
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # three ways to use N GPUs: one process per GPU (torchrun; what the driver launches), or - started as a plain process with
    # --gpus N - the reference's multi_CUDA back-end: one process, one host thread per GPU inside the library
    in_process = world == 1 and args.gpus > 1
    n_slabs = args.gpus if in_process else world
    torch.cuda.set_device(local)
    cores = pin_to_gpu_numa_node(local) if world > 1 else None
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    global REAL_BYTES
    f32 = args.real == "f32"
    if f32:
        assert world == 1 and args.gpus == 1, "--real f32 is a single-GPU informative mode"
        REAL_BYTES = 4
    lib = L.b200(args.real)
    if world > 1:       # torchrun exports OMP_NUM_THREADS=1; the host-side initialisation wants its share of the cores
        try:
            C.CDLL("libgomp.so.1").omp_set_num_threads(max(1, (cores or os.cpu_count() or 1) // (1 if cores else world)))
        except OSError:
            pass
    lib.lib.lgrngn_b200_set_rng_mode.argtypes = [C.c_int]
    lib.lib.lgrngn_b200_set_rng_mode(0)            # Philox in the kernels (tests/test_gpu_philox.py ties it to the oracle-verified path)
    ny, nz = args.ny, args.nz
    # weak scaling: args.nx columns per GPU; strong scaling: args.nx columns in total, split like distmem_opts.hpp:10-18
    if args.scaling == "strong":
        share = args.nx // n_slabs             # int / int in the reference: the .5 never rounds up
        nx_of = lambda r: share if r < n_slabs - 1 else args.nx - r * share
    else:
        nx_of = lambda r: args.nx
    nx_glob = sum(nx_of(r) for r in range(n_slabs))
    nx = nx_glob if in_process else nx_of(rank)
    if world > 1:
        D.configure(lib, rank, world, lft_x1=nx_of((rank - 1) % world) * 20.0, rgt_x0=0.0, n_x_tot=nx_glob)
    oi, o, f = make_case(lib, nx, ny, nz, args.sd_conc, pin=True, rank=rank, size=world, config=args.config)
    if in_process:
        oi.dev_count = n_slabs
    else:
        oi.dev_id = local
    t0 = time.time()
    p = lib.factory(L.backend_t.multi_CUDA if in_process else L.backend_t.CUDA, oi)
    p.init(f["th"], f["rv"], f["rhod"], None, f["Cx"], f["Cy"], f["Cz"])
    if world > 1:
        D.connect(lib, p, rank, world)             # one-time exchange of the inbox handles; step_async migrates from then on
    t_init = time.time() - t0
    engines = [D.engine_of(lib, p, d) for d in range(D.n_slabs(lib, p))]
    eng = engines[0]
    step_resident_fn = lib.lib.lgrngn_b200_step_resident_f32 if f32 else lib.lib.lgrngn_b200_step_resident
    step_resident_fn.argtypes = [C.c_void_p, C.c_int]
    proto = D.proto_of(lib, p)

    def resident_step():
        if step_resident_fn(proto, 0b1111) != 0:
            raise RuntimeError("resident step failed on rank %d (live SDs %s)" % (rank, [e_.n_part() for e_ in engines]))

    def host_step():
        api_step(p, o, f)

    def barrier():
        for e_ in engines:
            e_.sync()
        torch.cuda.synchronize()
        if dist:
            dist.barrier()

    def reduce(x, op):
        if not dist:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=getattr(dist.ReduceOp, op))
        return float(t.item())

    n_live = lambda: sum(e_.n_part() for e_ in engines)

    def timers_start():
        for e_ in engines:
            e_.timer_start()

    def timers_stop():
        return max(e_.timer_stop() for e_ in engines)

    def dry_volume():
        """4/3 pi sum(n rd^3) of the live super-droplets of this process + what left through the bottom (puddle) and the lid; the
        invariant of coalescence, transport and migration together"""
        p.diag_all()
        p.diag_dry_mom(3)
        cells = p.outbuf().reshape(f["rhod"].shape)
        live = float((cells * f["rhod"]).sum() * 20.0 ** 3 * 4.0 / 3.0 * np.pi)
        # freshly collided super-droplets carry the "invalid" fall speed -1 through the sedimentation step (reference quirk, kept):
        # the ones next to the lid leave through it, unaccounted for by the reference's puddle; the engine tallies them
        return live, float(p.diag_puddle()["dry_volume"]) + sum(e_.top_loss()[0] for e_ in engines)

    # the first API step uploads the fields; afterwards they are resident
    live0, pud0 = dry_volume()
    vol0 = reduce(live0 + pud0, "SUM")
    host_step()
    for _ in range(max(args.warmup - 1, 0)):
        resident_step()

    # ---- device-resident throughput, with the per-kernel profile of rank 0 taken INSIDE the timed window -----------------------
    props = torch.cuda.get_device_properties(local)
    try:
        pci = "%08x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
    except AttributeError:
        pci = None
    sampler = ClockSampler(local, pci)
    sampler.start()
    barrier()
    l0 = sum(e_.launches() for e_ in engines)
    updates = 0
    migrants = []
    timers_start()
    tw = time.time()
    for _ in range(args.steps):
        updates += n_live()
        resident_step()
        if n_slabs > 1:
            migrants.append(sum(D.migr_stats(lib, p)[:2]))
    ms = timers_stop()
    barrier()
    wall_ms = 1e3 * (time.time() - tw)
    launches = sum(e_.launches() for e_ in engines) - l0
    ms = reduce(ms, "MAX")
    total_updates = reduce(float(updates), "SUM")
    value = total_updates / (ms * 1e-3)
    live_end = reduce(float(n_live()), "SUM")

    # per-kernel profile: a second pass over the SAME number of steps right after the timed one (events around every launch
    # perturb the overlap of copies and kernels, so it is not folded into the timed loop; its state - steps warmup+K .. warmup+2K -
    # is the closest like-for-like one)
    roofline = None
    prof_table = None
    n_prof = eng.n_part()
    eng.profile(True)
    for _ in range(args.profile_steps):
        resident_step()
    rep = eng.profile_report() if rank == 0 else {}
    eng.profile(False)
    if rep:
        tot = sum(ms_ for _, ms_ in rep.values())
        prof_table = {k: {"launches": n, "ms": round(ms_, 3), "share": round(ms_ / tot, 4)} for k, (n, ms_) in sorted(rep.items(), key=lambda kv: -kv[1][1])}
        top = max(rep.items(), key=lambda kv: kv[1][1])
        name, (n_l, t_ms) = top
        per_sd = kernel_bytes(name)
        peak, src = peaks()
        achieved = per_sd * n_prof / (t_ms / n_l * 1e-3) / 1e9 if per_sd else None
        traffic_step = sum((traffic_of(k, n_prof) or 0.0) * n / args.profile_steps for k, (n, _) in rep.items())
        roofline = {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": peak, "peak_source": src, "unit": "GB/s",
                    "frac": (achieved / peak) if achieved else None, "traffic": traffic_of(name, n_prof),
                    "algorithmic_bytes_per_sd": per_sd, "sd_per_launch": n_prof, "mean_launch_ms": t_ms / n_l,
                    "step_frac_of_hbm_roofline": value * (A_FULL_BYTES_F32 if f32 else A_FULL_BYTES) / (n_slabs * peak * 1e9),
                    "step_dram_bytes_per_sd": round(traffic_step / n_prof, 1) if traffic_step else None,
                    "note": "the condensation kernel is %s-pipe / issue bound, not HBM bound (ncu: profiles/)" % ("FP32" if f32 else "FP64"),
                    "per_kernel": {k: {"GB/s": round(kernel_bytes(k) * n_prof / (ms_ / n * 1e-3) / 1e9, 1),
                                       "frac": round(kernel_bytes(k) * n_prof / (ms_ / n * 1e-3) / 1e9 / peak, 4)}
                                   for k, (n, ms_) in rep.items() if kernel_bytes(k) and ms_ > 0}}

    # ---- end to end through the API with host arrays ----------------------------------------------------------------
    barrier()
    upd2 = 0
    t0 = time.time()
    for _ in range(args.steps):
        upd2 += n_live()
        host_step()
    barrier()
    e2e_s = reduce(time.time() - t0, "MAX")
    clocks = sampler.result()          # sampled over both timed regions (device-resident and end-to-end)
    e2e_value = reduce(float(upd2), "SUM") / e2e_s
    h2d = sum(f[k].nbytes for k in ("th", "rv", "rhod", "Cx", "Cy", "Cz"))
    d2h = f["th"].nbytes + f["rv"].nbytes

    # ---- conservation over everything that ran so far (every rank, every N): dry volume live + rained out ------------------------
    live1, pud1 = dry_volume()
    vol1 = reduce(live1 + pud1, "SUM")
    conservation = {"dry_volume_rel_err": abs(vol1 - vol0) / vol0, "dry_volume_left_domain_frac": reduce(pud1, "SUM") / vol0,
                    "sd_live_start": nx_glob * ny * nz * args.sd_conc, "sd_live_end": int(reduce(float(n_live()), "SUM")),
                    "migrants_per_step": (reduce(float(np.mean(migrants)), "SUM") if migrants else 0.0),
                    "migrant_frac_per_slab_step": (reduce(float(np.mean(migrants)), "SUM") / max(live_end, 1.0) if migrants else 0.0)}
    assert conservation["dry_volume_rel_err"] < (1e-4 if f32 else 1e-10), "dry volume not conserved: %r" % (conservation,)

    # ---- the opt-in fast root search, for information (not the headline: see include/lcx_b200.h lcx_set_cond_solver) ----
    alt = None
    if n_slabs == 1 and not args.no_alt and not f32:
        E.set_cond_solver("secant")
        for _ in range(2):
            resident_step()
        upd3 = 0
        eng.timer_start()
        for _ in range(args.steps):
            upd3 += eng.n_part()
            resident_step()
        ms3 = eng.timer_stop()
        E.set_cond_solver("toms748")
        alt = {"cond_solver": "secant", "value": upd3 / (ms3 * 1e-3), "ms_per_step": ms3 / args.steps,
               "note": "safeguarded secant instead of the reference's TOMS 748 trial points: within 2^-15 per step of the reference, "
                       "different trajectory; informative only"}

    # ---- CPU baseline beside it (rank 0, N = 1) -----------------------------------------------------------------------
    cpu = None
    if rank == 0 and n_slabs == 1 and not args.no_cpu_baseline and not f32:
        try:
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "3", "--warmup", "1"],
                                 capture_output=True, text=True, timeout=900)
            cpu = json.loads(out.stdout.strip().splitlines()[-1])["cpu_baseline"]
        except Exception as ex:      # the baseline is informative; its absence must not hide the GPU number
            cpu = {"value": None, "unit": "SD-updates/s", "cores": os.cpu_count(), "kind": "reference", "sample": "failed: %r" % (ex,)}

    if rank == 0:
        n_cell, max_count = eng.cell_stats()
        sd_total = nx_glob * ny * nz * args.sd_conc
        line = {
            "metric": "super-droplet updates/s (cond+coal+sedi+adve step)", "value": value, "unit": "SD-updates/s",
            "n_gpus": n_slabs, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": args.real, "data": "synthetic",
            "config": {"workload": "%s%s x-slab%s: %dx%dx%d cells x %d SD/cell, hall_davis_no_waals, beard77fast, implicit adve, sstp 1/1, Cx=%s%s"
                                   % ("SINGLE-PRECISION ENGINE (factory<float>, second mode, not the headline) " if f32 else "", args.config, " per GPU" if args.scaling == "weak" else "s of a fixed domain", nx_of(0), ny, nz, args.sd_conc,
                                      "0.5" if args.config == "cfg5" else "0.1", ", rain mode (live-SD-weighted throughput)" if args.config == "cfg5" else ""),
                       "sd_per_gpu": sd_total // n_slabs, "global_cells": [nx_glob, ny, nz], "rng": "philox4x32-10",
                       "multi_gpu": ("multi_CUDA: one process, one host thread per GPU" if in_process else
                                     "one process per GPU, migrants packed into the neighbours' inboxes over CUDA-IPC peer memory" if world > 1 else "single GPU"),
                       "cond_solver": "toms748 (the reference's trial points)", "cond_layout": E.get_cond_layout(), "lazy_gather": os.environ.get("LCX_LAZY_GATHER", "1") != "0",
                       "l2": "inputs_exceed_l2 (%.1f GB of SD state per GPU)" % (sd_total / n_slabs * (44 if f32 else 76) / 1e9),
                       "init_s": round(t_init, 2), "max_sd_per_cell": max_count, "wall_ms_per_step": wall_ms / args.steps},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "SD-updates/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * e2e_s / args.steps},
            "gpu_launches": int(launches), "conservation": conservation,
            "roofline": roofline, "cpu_baseline": cpu, "opt_in_fast_solver": alt, "kernels": prof_table,
        }
        print(json.dumps(line), flush=True)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "reference-cuda"])
    ap.add_argument("--config", default="cfg4", choices=["cfg4", "cfg5"], help="BASELINE.json configs[3] / configs[4]")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="weak: --nx columns per GPU; strong: --nx columns in total")
    ap.add_argument("--no-alt", action="store_true", help="skip the informative opt-in secant pass")
    ap.add_argument("--real", default="f64", choices=["f64", "f32"], help="f32: the single-precision engine (factory<float>), single GPU, informative")
    ap.add_argument("--nx", type=int, default=64)
    ap.add_argument("--ny", type=int, default=256)
    ap.add_argument("--nz", type=int, default=128)
    ap.add_argument("--sd-conc", type=int, default=40)
    ap.add_argument("--ref-nx", type=int, default=64)      # CPU arm: 64^3 cells x 40 SD = 1.05e7 SDs, about 1.5 s per step on 16 cores
    ap.add_argument("--ref-ny", type=int, default=64)
    ap.add_argument("--ref-nz", type=int, default=64)
    ap.add_argument("--profile-steps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-strict", action="store_true", help="time the IEEE-strict -O2 build of the reference instead of its -Ofast build")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "reference-cuda":
        run_reference_cuda(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
