"""CPU-side check of bench.py's output contract: the reference arm runs without a GPU, prints exactly one JSON line on
stdout with the keys the driver reads, and the product arm refuses to run without CUDA."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=300, env=env)


def test_reference_arm_prints_one_json_line():
    out = _run("--impl", "reference", "--ref-points", "60000", "--queries", "20000", "--steps", "2", "--warmup", "1")
    assert out.returncode == 0, out.stderr
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "queries/s" and d["value"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]


def test_product_arm_needs_cuda():
    import torch
    if torch.cuda.is_available():
        return
    out = _run("--ref-points", "1000", "--queries", "1000", "--steps", "1")
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)
    assert out.stdout.strip() == ""


def test_reference_arm_ignores_omp_num_threads_1():
    """torch.distributed.run exports OMP_NUM_THREADS=1; the reference arm must still use every core it may run on, so that the
    CPU baseline is the same at every --gpus N (round-1 VERDICT: the N >= 2 ratios were void because `cores` fell to 1)."""
    avail = len(os.sched_getaffinity(0))
    out = _run("--impl", "reference", "--gpus", "2", "--ref-points", "60000", "--queries", "20000", "--steps", "1", "--warmup", "1", env={**os.environ, "OMP_NUM_THREADS": "1"})
    assert out.returncode == 0, out.stderr
    d = json.loads([ln for ln in out.stdout.splitlines() if ln.strip()][0])
    assert d["cpu_baseline"]["cores"] == avail and d["n_gpus"] == 2 and d["scaling"] == "strong"
