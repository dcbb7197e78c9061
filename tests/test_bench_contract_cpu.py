"""bench.py contract checks that need no GPU: the reference arm runs the UNMODIFIED reference (baseline/_ref or
/root/reference), prints ONE JSON line with the contract keys, and names the same `config` as the B200 arm."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_reference_arm_line():
    from baseline import refshim
    if not refshim.available():
        pytest.skip("reference tree not present")
    env = dict(os.environ, OMP_NUM_THREADS=str(min(8, os.cpu_count() or 1)))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                          "--warmup", "1"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, env=env,
                         timeout=600, cwd=ROOT)
    assert out.returncode == 0
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, "exactly one JSON line"
    line = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["metric"] == "clips_per_sec" and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    import bench
    assert line["config"] == bench.base_config("fp32")          # identical to the B200 arm's `config`
    assert line["vs_baseline"] is None and line["value"] > 0


def test_other_ranks_of_the_reference_arm_exit_without_work():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "2"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, env=env,
                         timeout=120, cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == ""
