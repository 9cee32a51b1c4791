"""CPU check of the bench.py contract: the reference arm (the only arm that runs without a GPU) prints exactly one JSON line on
stdout with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--cpu-sample-nodes", "400", "--no-also"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "GCL nodes/sec" and d["unit"] == "nodes/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "nodes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["steps"] == 1 and d["warmup"] == 1                      # the driver's --steps / --warmup are honoured
    assert d["config"]["workload"].startswith("cfg4") and d["scaling"] == "strong"


def test_config_is_shared_by_both_arms():
    """Both arms print the SAME config object for a given (--config, N, --mode): the driver compares them key by key."""
    sys.path.insert(0, ROOT)
    import bench

    c1 = bench.config_dict("cfg4", 8, "rowshard")
    assert c1 == bench.config_dict("cfg4", 8, "rowshard") and c1["nodes"] == 130_000 and "rowshard8" in c1["parallelism"]
    assert bench.config_dict("cfg2", 4, "dp")["nodes"] == 4 * 28_000


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True,
                       text=True, timeout=300, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
