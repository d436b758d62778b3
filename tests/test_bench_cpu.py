"""bench.py contract checks that need no GPU: the reference arm (CPU oracle) prints one JSON line with the agreed keys, and the
CUDA arm refuses to run -- loudly, non-zero -- when there is no device (no CPU fallback for the product path)."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, capture_output=True, text=True, timeout=900)


def test_reference_arm_json_contract():
    r = _run("--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "1", "--cpu-batch", "2")
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["impl"] == "reference" and d["unit"] == "crops/s" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["value"] > 0 and d["vs_baseline"] is None and d["data"] == "synthetic"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    # same metric / unit as the CUDA arm (BASELINE.json's metric)
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert isinstance(base, dict)
    assert "crops" in d["metric"]


def test_cuda_arm_fails_loudly_without_a_gpu():
    if torch.cuda.is_available():
        import pytest

        pytest.skip("a GPU is present")
    r = _run("--steps", "1", "--warmup", "1")
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stdout + r.stderr)
