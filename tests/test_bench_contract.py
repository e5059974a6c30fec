"""bench.py's output contract on the arm that runs without a GPU (`--impl reference`: the oracle timed on the host cores): exactly
ONE line on stdout, a JSON object with the driver's keys on the engine arm's metric / unit / config; under torchrun only rank 0 works."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(extra_env, *args):
    env = dict(os.environ)
    env.update(extra_env)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"] + list(args),
                          capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)


def test_reference_arm_prints_one_json_line():
    r = run_bench({})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = r.stdout.splitlines()
    assert len(lines) == 1, r.stdout[:500]          # nothing but the JSON line reaches stdout (library chatter goes to stderr)
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "frames/sec at 656x368 COCO-18" and d["unit"] == "frames/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0
    assert d["config"]["workload"].startswith("C2: COCO 656x368")
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "full" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_do_no_work():
    r = run_bench({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout == ""
