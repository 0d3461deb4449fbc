"""CPU: the reference arm of bench.py (`--impl reference`) prints ONE JSON line that carries the contract's keys; the GPU arm
refuses to run without a CUDA device instead of falling back to anything."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT, env=e, capture_output=True, text=True,
                          timeout=600)


def test_reference_arm_line():
    r = run(["--impl", "reference", "--steps", "1", "--warmup", "1", "--scale", "0.004"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    assert d["metric"] == "effective_GBps_reduce_scan_where" and d["unit"] == "GB/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] == 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and d["config"]["scale"] == 0.004
    assert d["vs_baseline"] is None and d["data"] == "synthetic"


def test_reference_arm_other_ranks_stay_silent():
    r = run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1", "--scale", "0.004"],
            env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.strip() == ""


def test_gpu_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present")
    r = run(["--steps", "1", "--warmup", "1", "--scale", "0.004", "--no-e2e", "--no-cpu-baseline"])
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout)
