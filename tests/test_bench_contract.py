"""The parts of bench.py's contract that can be checked without a GPU: the reference arm (the oracle port timed on the
host cores) prints ONE JSON line with the keys the driver reads, and the product arm refuses to run without a B200."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _run(*args):
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, timeout=600)


def test_reference_arm_prints_the_contract_line():
    out = _run("--impl", "reference", "--items", "20000", "--features", "64", "--queries", "64", "--steps", "1",
               "--warmup", "1")
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1                                     # library banners go to stderr, the line to stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "lambda_tau_build_items_per_s" and d["unit"] == "items/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
    assert d["value"] > 0 and d["search_qps"] > 0 and d["vs_baseline"] is None and d["dtype"] == "f64"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "items/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("GPU present")
    out = _run("--items", "2000", "--features", "32", "--queries", "8", "--steps", "1", "--warmup", "1",
               "--no-cpu-baseline", "--no-e2e")
    assert out.returncode != 0                                 # no CPU fallback: the CUDA context cannot be created
    assert not [ln for ln in out.stdout.splitlines() if ln.strip().startswith("{")]
