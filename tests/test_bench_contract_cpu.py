"""CPU: the reference arm of bench.py (the one leg that needs no GPU) keeps the JSON contract of the task: one line,
metric / unit / n_gpus / steps / warmup / higher_is_better / impl / cpu_baseline / e2e with zero transfer bytes; under a
world-size-2 launch only rank 0 prints."""

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lines(out):
    return [json.loads(l) for l in out.splitlines() if l.startswith("{")]


def test_reference_arm_contract():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                          "--cpu-sample", "200000"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = _lines(res.stdout)
    assert len(lines) == 1
    d = lines[0]
    assert d["impl"] == "reference" and d["unit"] == "QP/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1 and d["vs_baseline"] is None and d["dtype"] == "f64"
    assert d["metric"].startswith("quadrature points per second") and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] > 0 and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "QP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_world2_only_rank0_prints():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(29300 + os.getpid() % 90), os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
           "--steps", "2", "--warmup", "1", "--cpu-sample", "200000"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = _lines(res.stdout)
    assert len(lines) == 1 and lines[0]["impl"] == "reference" and lines[0]["n_gpus"] == 2
