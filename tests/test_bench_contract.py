"""bench.py contract checks that need no GPU: the reference arm prints exactly one JSON line with the agreed keys, ranks
other than 0 stay silent, and the engine arm refuses to run (no CPU fallback) without a CUDA device."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env_extra=None):
    env = dict(os.environ, AHR_BENCH_CPU_ROWS="20000", AHR_BENCH_CPU_QUERIES="8")
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=env,
                          cwd=ROOT, timeout=300)


@pytest.mark.timeout(300)
def test_reference_arm_prints_one_json_line():
    p = _run(["--impl", "reference", "--steps", "2", "--warmup", "1"])
    assert p.returncode == 0, p.stderr[-500:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "queries/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("queries/sec (top-k=50, 10M")
    assert d["value"] > 0 and d["steps"] == 2 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] and "sample" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    # the line is self-consistent: ms_per_step is the time of one full 4096-query step at the measured rate
    assert abs(d["ms_per_step"] * d["value"] / 1000.0 - 4096) < 1e-6 * 4096


@pytest.mark.timeout(120)
def test_reference_arm_uses_all_host_threads_under_torchrun_env():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm must still use every host core it may run on."""
    p = _run(["--impl", "reference", "--steps", "1", "--warmup", "0"], {"OMP_NUM_THREADS": "1", "RANK": "0", "WORLD_SIZE": "2"})
    assert p.returncode == 0, p.stderr[-500:]
    d = json.loads(p.stdout.strip())
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))


@pytest.mark.timeout(120)
def test_reference_arm_other_ranks_are_silent():
    p = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], {"RANK": "1", "WORLD_SIZE": "2"})
    assert p.returncode == 0 and p.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a machine without a GPU")
def test_engine_arm_fails_loudly_without_gpu():
    p = _run(["--steps", "1", "--warmup", "0"])
    assert p.returncode != 0 and p.stdout.strip() == "" and "no CUDA device" in p.stderr
