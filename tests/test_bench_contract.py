"""CPU: the driver-facing contract of bench.py that can be checked without a GPU — the reference arm
prints ONE JSON line with the required keys, and under a multi-rank launch only rank 0 speaks."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REQUIRED = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"}


def test_reference_arm_json_line():
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                                   "--warmup", "1", "--scale", "0.05"], text=True, timeout=600)
    lines = [ln for ln in out.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert REQUIRED <= set(d)
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["value"] > 0 and "workload" in d["config"]
    # both arms print the SAME config object (the driver compares them): built by one function
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.workload_config(0.05)
    assert d["cpu_baseline_1thread"]["cores"] == 1 and d["cpu_baseline"]["cores"] == bench.usable_threads()
    assert d["sample_pods_per_step"] == d["config"]["pods_per_gpu"]   # the whole snapshot per step


def test_reference_arm_other_ranks_are_silent():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                                   "--steps", "1", "--warmup", "0"], text=True, timeout=120, env=env)
    assert out.strip() == ""
