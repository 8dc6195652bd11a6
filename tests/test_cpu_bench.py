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
