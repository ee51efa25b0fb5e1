"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`) prints the line the driver reads,
alone and under a two-process torchrun launch (rank 0 prints, the other rank exits 0 without work)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ARGS = ["--impl", "reference", "--steps", "2", "--warmup", "1", "--batch", "6", "--size", "32"]
KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"}


def _check(line, n_gpus):
    assert KEYS <= line.keys()
    assert line["impl"] == "reference" and line["n_gpus"] == n_gpus and line["steps"] == 2
    assert line["unit"] == "Mpix/s" and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["value"] > 0 and line["gpu_launches"] == 0
    assert "workload" in line["config"] and "model" not in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def _json_lines(text):
    return [json.loads(l) for l in text.splitlines() if l.startswith("{")]


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--gpus", "1"] + ARGS, capture_output=True,
                         text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = _json_lines(out.stdout)
    assert len(lines) == 1
    _check(lines[0], 1)


@pytest.mark.timeout(600)
def test_reference_arm_under_torchrun_prints_once():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29577", os.path.join(ROOT, "bench.py"), "--gpus", "2"] + ARGS
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = _json_lines(out.stdout)
    assert len(lines) == 1                       # rank 0 alone runs and prints
    _check(lines[0], 2)
