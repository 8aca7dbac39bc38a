"""bench.py contract checks that need no GPU: the reference arm (the reference's CPU decode through
oracle/_ref) prints one JSON line with the keys the driver reads, and exits 0."""
import json
import os
import subprocess
import sys

import gst_fixtures as fx

BENCH = os.path.join(fx.ROOT, "bench.py")


def _run(*extra, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--config", "1", "--steps", "1", "--warmup", "0",
                        "--distinct", "2", "--cpu-sample", "4", *extra], capture_output=True, text=True, env=e, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout.strip().splitlines()


def test_reference_arm_prints_the_contract_line(ref_lib):
    lines = _run()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "GTexel/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("decoded GTexel/s") and d["value"] > 0
    assert d["config"]["workload"].startswith("configs[1]")
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "GTexel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["steps"] == 1 and d["n_gpus"] == 1 and d["gpu_launches"] == 0


def test_reference_arm_only_rank0_works_under_torchrun(ref_lib):
    """Launched as N ranks, rank 0 alone runs and prints; the others exit 0 without output."""
    assert _run(env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}) == []
    lines = _run(env={"RANK": "0", "LOCAL_RANK": "0", "WORLD_SIZE": "2"})
    assert len(lines) == 1 and json.loads(lines[0])["impl"] == "reference"
