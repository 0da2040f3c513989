"""bench.py's contract, as far as it can be checked without a GPU: the reference arm (`--impl reference`, the reference's own
CPU plQuery -- oracle/_ref when it is built here, else the oracle port) prints one JSON line with the agreed keys, and the
product arm refuses to run without a CUDA device instead of falling back to anything."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, **kw):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                          cwd=ROOT, timeout=600, **kw)


def test_reference_arm_prints_the_contract_line(oracle_built):
    r = _run(["--impl", "reference", "--workload", "small", "--steps", "2", "--warmup", "3"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "k=21 SA queries/sec" and d["unit"] == "queries/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["dtype"] == "int64" and d["data"] == "synthetic" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert d["gpu_launches"] == 0


def test_reference_arm_runs_on_rank_0_only(oracle_built):
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = _run(["--impl", "reference", "--workload", "small", "--gpus", "2", "--steps", "1"], env=env)
    assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_product_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present")
    r = _run(["--workload", "small", "--steps", "1"])
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]
