"""bench.py's contract as far as a machine without a GPU can check it: the reference arm (the oracle port of the reference's CPU
path, the one place besides tests/ and smoke() that may execute oracle/) prints ONE JSON line with the keys the driver reads, and
the native arm refuses to run without a CUDA device instead of falling back to anything."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=600):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout,
                          cwd=ROOT)


def test_reference_arm_prints_one_json_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines                      # only the JSON line reaches stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "frames_per_sec_720p" and d["unit"] == "frames/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["value"] > 0 and d["gpu_launches"] == 0 and d["vs_baseline"] is None
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert "workload" in d["config"] and "model" not in d["config"]


def test_native_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present: the native arm runs (covered by the -m gpu tests and the driver)")
    r = _run("--steps", "1", "--warmup", "0", "--no-e2e", "--no-cpu-baseline", timeout=300)
    assert r.returncode != 0
    assert "no CPU fallback" in r.stderr or "CUDA" in r.stderr
    assert not [l for l in r.stdout.splitlines() if l.strip().startswith("{")]        # no number without the CUDA path


def test_defaults_are_the_documented_workload():
    sys.path.insert(0, ROOT)
    import importlib
    bench = importlib.import_module("bench")
    argv = sys.argv
    try:
        sys.argv = ["bench.py"]
        a = bench.parse_args()
    finally:
        sys.argv = argv
    assert (a.gpus, a.steps, a.warmup, a.chunk, a.lanes, a.impl, a.workload) == (1, 80, 3, 4050, 2, "native", "game")
    assert a.steps * a.chunk == 324_000                  # configs[1]: the full game
