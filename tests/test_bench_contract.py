"""bench.py contract checks that need no GPU: the reference arm (CPU oracle port) prints ONE JSON line with the
keys the driver reads, and the B200 arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--points", "4000", "--voxel", "0.02"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "points/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == os.cpu_count() and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    for k in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config"):
        assert k in d
    assert "workload" in d["config"] and "model" not in d["config"]
    # the line declares the size it actually ran (round-1 finding: 100 k points timed under a 1 M-point label)
    assert d["config"]["points_per_tree"] == 4000 and d["config"]["total_points"] == 4000 and "4000 points" in d["cpu_baseline"]["sample"]
    assert d["steps_timed"] == 1


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a box without a GPU")
def test_b200_arm_fails_loudly_without_gpu():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True, text=True,
                       timeout=600, cwd=ROOT)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_workload_config_names_replicas_and_distinct_trees():
    """c2 / c4 name one tree: at N ranks every rank runs a replica (fixed per-GPU work); c3 is a batch of distinct trees."""
    sys.path.insert(0, ROOT)
    import importlib
    bench = importlib.import_module("bench")
    old = sys.argv
    try:
        sys.argv = ["bench.py", "--gpus", "8"]
        a = bench.parse()
        c = bench.workload_config(a, 8)
        assert c["config"] == "c2" and c["tree_seeds"] == "0 on every rank" and c["total_points"] == 8_000_000 and "replica" in c["parallelism"]
        sys.argv = ["bench.py", "--gpus", "8", "--config", "c3"]
        a = bench.parse()
        c = bench.workload_config(a, 8)
        assert c["tree_seeds"] == "0..7" and c["points_per_tree"] == 500_000 and c["total_points"] == 4_000_000
        sys.argv = ["bench.py", "--gpus", "2", "--distinct-trees", "--seed0", "3"]
        a = bench.parse()
        assert bench.workload_config(a, 2)["tree_seeds"] == "3..4"
        sys.argv = ["bench.py", "--config", "c5"]
        a = bench.parse()
        c = bench.workload_config(a, 8)
        assert c["total_points"] == 20_000_000 and c["block_size"] == 0.64
    finally:
        sys.argv = old
