"""bench.py prints exactly ONE JSON line on stdout with the contract's keys.  Only the reference
arm can run without a GPU (it times the CPU oracle port); the GPU arm prints an error line here."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=e, timeout=600)
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    return out.returncode, lines


def test_reference_arm_json_contract():
    rc, lines = run(["--impl", "reference", "--steps", "2", "--warmup", "1"], {"DTE_BENCH_REF_SECONDS": "0.3"})
    assert rc == 0 and len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "tuples/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and d["vs_baseline"] is None


def test_reference_arm_other_ranks_exit_quietly():
    rc, lines = run(["--impl", "reference", "--gpus", "2"], {"RANK": "1", "WORLD_SIZE": "2", "DTE_BENCH_REF_SECONDS": "0.3"})
    assert rc == 0 and lines == []


def test_gpu_arm_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("GPU present")
    rc, lines = run(["--steps", "1"])
    assert rc != 0 and len(lines) == 1 and "error" in json.loads(lines[0])
