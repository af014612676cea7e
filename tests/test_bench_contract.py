"""CPU: the reference arm of bench.py prints one JSON line with the contract's keys; a GPU arm without
a GPU refuses loudly instead of falling back."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True,
                          text=True, env=e, timeout=600)


@pytest.mark.timeout(600)
def test_reference_arm_line():
    p = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--workload", "volume2a",
                  "--cpu-seconds", "0.5")
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "impl"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "pair-evals/s" and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["value"] > 1e6 and d["cpu_baseline"]["value"] == d["value"]


def test_reference_arm_other_ranks_are_silent():
    p = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--workload", "volume2a",
                  env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_gpu_arm_without_gpu_fails_loudly():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    p = run_bench("--steps", "1", "--warmup", "3")
    assert p.returncode != 0
    assert "no CPU fallback" in (p.stderr + p.stdout)
