"""The reference arm of bench.py (the reference's own CPU code, or the oracle port where oracle/_ref did not travel)
on a tiny cloud: it must print ONE JSON line with the keys the driver reads, run on rank 0 only, and name its sample."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--particles", "20000",
                        "--ref-stride", "4", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600,
                       env=env, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    return p.stdout.strip()


def test_reference_arm_line():
    out = _run()
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "steps/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["config"]["n_particles"] == 20000 and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "FULL un-sampled" in cb["sample"]
    # the timed steps are full steps: the claimed work fits the run, and the sampled estimate is reported beside it
    assert abs(d["ms_per_step"] - 1e3 * d["full_step_s"]) < 1e-9 and d["full_steps_timed"] == d["steps"]
    assert d["sampled_estimate_s"] > 0 and 0.2 < d["sampled_over_full"] < 5
    # both arms share the config dictionary
    sys.path.insert(0, ROOT)
    import argparse
    import bench
    w = bench.make_workload("lamb", 20000)
    assert d["config"] == bench.make_config(argparse.Namespace(), w)
    assert d["e2e"] == {"value": d["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    assert _run({"RANK": "1", "WORLD_SIZE": "2"}) == ""
