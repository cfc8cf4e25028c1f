"""bench.py command line on a machine without a GPU: the reference arm (CPU oracle port) prints the contract's
JSON line, and the product arm refuses to run instead of falling back to anything."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMALL = ["--rows", "4000", "--dim", "32", "--lists", "8", "--queries", "64", "--batch", "32", "--steps", "1", "--warmup", "1"]


def _run(args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT, capture_output=True, text=True,
                          timeout=600)


def test_reference_arm_prints_one_json_line():
    p = _run(["--impl", "reference"] + SMALL)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, p.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["metric"] == json.load(open(os.path.join(ROOT, "BASELINE.json")))["metric"]
    assert d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]


def test_product_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    p = _run(SMALL)
    assert p.returncode != 0
    assert not p.stdout.strip()          # no number is ever printed from a CPU path


def test_reference_arm_of_the_other_configs():
    """--config A (exact scan) and --config D (k-means build) also have a CPU arm that prints the contract's line"""
    for extra, unit in ((["--config", "A", "--steps", "1", "--warmup", "0"], "queries/s"),
                        (["--config", "D", "--dim", "16"], "rows/s")):
        p = _run(["--impl", "reference"] + extra)
        assert p.returncode == 0, p.stderr[-2000:]
        lines = [l for l in p.stdout.splitlines() if l.strip()]
        assert len(lines) == 1, p.stdout
        d = json.loads(lines[0])
        assert d["impl"] == "reference" and d["unit"] == unit and d["value"] > 0
        assert d["cpu_baseline"]["value"] == d["value"] and "workload" in d["config"]
