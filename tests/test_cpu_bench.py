"""CPU-only checks of bench.py's contract: the reference arm (the reference's own OpenMP back-end on the host cores) prints one JSON
line with the keys the driver reads, also under a 2-rank torchrun launch (rank 0 alone works and prints)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMALL = ["--steps", "1", "--warmup", "1", "--ref-nx", "8", "--ref-ny", "8", "--ref-nz", "8"]


def check_line(out, n_gpus):
    lines = [l for l in out.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "SD-updates/s" and d["higher_is_better"] is True and d["n_gpus"] == n_gpus
    assert d["metric"].startswith("super-droplet updates/s") and d["value"] > 0 and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["value"] == d["value"] and cb["cores"] >= 1 and "OpenMP" in cb["sample"]
    assert "workload" in d["config"]
    return d


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference"] + SMALL, capture_output=True, text=True,
                       timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    check_line(r.stdout, 1)


def test_reference_arm_under_torchrun_only_rank0_works():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"] + SMALL
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    d = check_line(r.stdout, 2)
    assert d["cpu_baseline"]["cores"] == (os.cpu_count() or 1)      # torchrun's OMP_NUM_THREADS=1 is overridden for the one working rank


def test_roofline_tables_cover_the_kernels_the_profile_names():
    """the per-kernel roofline of bench.py maps profile names (with template arguments) to algorithmic bytes and to the DRAM traffic of
    the committed ncu captures"""
    sys.path.insert(0, ROOT)
    import bench
    bench.LAZY = False
    for name, per_sd in (("(k_cond_range<M, false>)", 52.0), ("(k_cond_range<M, true>)", 88.0), ("(k_cond_cells<M>)", 48.0), ("k_gather", 136.0),
                         ("k_coal_small", 76.0), ("k_transport<false>", 72.0), ("k_transport<true>", 76.0), ("(k_vterm_beard77<true>)", 24.0),
                         ("k_mv_count", 4.0), ("k_cond_staged<true>", 88.0), ("k_cond_staged<false>", 52.0),
                         ("(k_cond_classed<M, true>)", 96.0), ("(k_cond_classed<M, false>)", 60.0)):
        assert bench.kernel_bytes(name) == per_sd, name
    bench.LAZY = True
    assert bench.kernel_bytes("k_gather") == 12.0        # gather-on-read: only the storage index is left to the gather kernel
    for name in ("k_mv_counts", "k_scan_tiles", "k_radix_scatter", "k_cells_Tpr"):
        assert bench.kernel_bytes(name) == 0.0, name          # per-cell / per-mover kernels carry no per-SD figure
    for name in ("(k_cond_range<M, false>)", "k_gather", "k_transport<false>", "k_coal_small", "(k_vterm_beard77<true>)"):
        t = bench.traffic_of(name, 1)
        assert t is not None and 10. < t < 250., (name, t)
    assert bench.peaks()[0] > 1000.
