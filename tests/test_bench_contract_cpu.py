"""The reference arm of bench.py runs without a GPU (it times the oracle's C restatement of the reference CPU path): its JSON line must
carry the keys the driver reads, for the same metric / unit the GPU arm reports."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None):
    env = dict(os.environ)
    env.update(extra_env or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run()
    assert d["impl"] == "reference"
    assert d["metric"] == "MiniLM-L6 seq128 embeddings/sec" and d["unit"] == "embeddings/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1 and d["gpu_launches"] == 0
    assert d["value"] > 0 and abs(d["ms_per_step"] * 1e-3 * d["value"] - 32) < 1e-3 * 32  # one step = 32 sequences (BASELINE configs[0])
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_ignores_torchrun_thread_pin():
    """torchrun exports OMP_NUM_THREADS=1; the reference arm still uses the host's cores (per-N ratios were inflated)."""
    d = _run({"OMP_NUM_THREADS": "1"})
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    if ncpu >= 2:
        assert d["cpu_baseline"]["cores"] >= 2
