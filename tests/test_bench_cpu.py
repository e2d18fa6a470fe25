"""bench.py's reference arm (the oracle port on host cores) honours the driver's JSON contract without a GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                          capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)


def test_reference_arm_json_line():
    p = _run()
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "pixels/sec full-scene inference"
    assert line["unit"] == "pixels/s" and line["higher_is_better"] is True and line["value"] > 0
    assert line["steps"] == 1 and line["warmup"] == 0 and line["n_gpus"] == 1 and line["vs_baseline"] is None
    assert "610x340x103" in line["config"]["workload"] and "model" not in line["config"]
    cb = line["cpu_baseline"]
    have_ref = os.path.exists(os.path.join(ROOT, "baseline", "_ref", "tools", "hyper_tools.py"))
    # the stock reference (baseline/_ref, staged by __graft_entry__.build()) when present, else the oracle port
    assert cb["kind"] == ("reference" if have_ref else "port")
    assert cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    e = line["e2e"]
    assert e["value"] == line["value"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    # both arms print the same `config` (the driver's same_config check)
    sys.path.insert(0, ROOT)
    import bench
    assert line["config"] == bench.workload_config(bench.CONFIGS["c1"], 1)
    assert line["scaling"] == "strong"


def test_reference_arm_port_fallback():
    p = _run({"CMLPL_BENCH_NO_REF": "1"})
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["cpu_baseline"]["kind"] == "port" and line["value"] > 0


def test_reference_arm_other_ranks_do_no_work():
    p = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert p.returncode == 0 and p.stdout.strip() == ""
