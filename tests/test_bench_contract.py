"""bench.py contract (CPU part): the reference arm prints one JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    # PB_BENCH_REFERENCE=port: skip the minute of numba compilation the unmodified reference needs where it is present
    env = dict(os.environ, OMP_NUM_THREADS="4", PB_BENCH_REFERENCE="port")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1",
                        "--steps", "1", "--warmup", "1"], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "wave-points/s"
    assert d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1 and d["dtype"] == "f64"
    assert d["config"]["workload"].startswith("reflected_toon_1d L=60 W=10000")
    # the reference is single-threaded numba: `value` is a 1-core figure, the OpenMP all-threads figure sits beside it
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 1
    assert d["cpu_baseline"]["all_threads_value"] > 0 and d["cpu_baseline"]["all_threads_cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["e2e_spectrum"]["value"] > 0


def test_reference_arm_other_ranks_are_silent():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "1"], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_reference_arm_of_another_config():
    """bench.py --config cfg2 --impl reference: same line shape, the configuration's own metric and workload"""
    env = dict(os.environ, OMP_NUM_THREADS="4")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--config", "cfg2", "--impl", "reference",
                        "--steps", "1", "--warmup", "1"], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][0])
    assert d["impl"] == "reference" and d["config"]["bench_config"] == "cfg2" and d["value"] > 0
    assert d["config"]["workload"].startswith("thermal_toon_1d L=90 W=10000")
