"""bench.py contract on CPU: the reference arm (`--impl reference`) runs here without a GPU and prints ONE JSON line with the
keys the driver reads; the b200 arm needs a GPU and is covered by the round-end run."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("force_port", [False, True])
def test_reference_arm_prints_one_json_line(force_port):
    env = dict(os.environ)
    if force_port:
        env["CEED_B200_BENCH_FORCE_PORT"] = "1"   # as if oracle/_ref had not travelled: oracle restatement, kind = "port"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                "config", "impl", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "GDoF/s" and d["value"] > 0 and d["dtype"] == "f64"
    assert d["config"]["workload"].startswith("BP1 mass p=3")
    assert d["cpu_baseline"]["kind"] == ("port" if force_port or not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "lib", "libceed.so")) else "reference")
    assert d["cpu_baseline"]["cores"] >= 1 and d["e2e"]["h2d_bytes_per_step"] == 0


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
