"""bench.py contract pieces that can be checked without a GPU: the reference arm (the oracle's reference-style CPU drivers on a bounded
sample) prints one JSON line with the agreed keys; under torchrun only rank 0 works."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(env_extra):
    env = dict(os.environ, **env_extra)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--res", "5", "--scene", "pile",
                          "--tets", "30000"], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout.strip()


def test_reference_arm_line():
    line = json.loads(run({}).splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "newton_iteration_ms_assembly_ccd" and line["unit"] == "ms"
    assert line["higher_is_better"] is False and line["value"] > 0 and line["dtype"] == "f64"
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "tets" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_do_nothing():
    assert run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == ""
