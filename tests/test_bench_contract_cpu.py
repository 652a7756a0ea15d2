"""CPU: bench.py's reference arm prints exactly one JSON line with the contract's keys, and the CUDA arm refuses to
run without a GPU (no silent CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, cwd=ROOT,
                          env={**os.environ, **(env or {})}, timeout=600)


def test_reference_arm_line():
    run = _run("--impl", "reference", "--steps", "2", "--warmup", "1", "--ref-rays", "3000", "--pyref-rays", "16")
    assert run.returncode == 0, run.stderr[-2000:]
    lines = [l for l in run.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "interactions/s" and d["dtype"] == "f64" and d["steps"] == 2
    assert d["value"] > 1e4 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["vs_baseline"] is None
    assert d["config"]["workload"] == "c2_4f_telescope"
    # beside the C port: the unmodified Python reference (pip --target copy or /root/reference), one process per core
    py = d["cpu_baseline_reference"]
    from oracle import ref_harness as RH

    if RH.reference_available():
        assert py["kind"] == "reference" and py["value"] > 1 and py["cores"] >= 1
        assert d["value"] > 20 * py["value"]  # the port arm is the conservative (much faster) CPU denominator


def test_reference_arm_other_ranks_do_nothing():
    run = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-rays", "1000", "--pyref-rays", "8", env={"RANK": "1", "WORLD_SIZE": "2"})
    assert run.returncode == 0 and run.stdout.strip() == ""


def test_cuda_arm_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    run = _run("--steps", "1", "--warmup", "3")
    assert run.returncode != 0 and run.stdout.strip() == ""
