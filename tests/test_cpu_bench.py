"""bench.py's reference arm (the one leg that runs without a GPU): the JSON line keeps the contract's keys, names the same
config as the repo arm, and under a multi-rank launch only rank 0 works."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env, *flags):
    env = dict(os.environ, **extra_env)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                           "--n-src", "200000", "--grid", "8"] + list(flags), capture_output=True, text=True, env=env, timeout=300)


def test_reference_arm_line_and_rank_gating():
    res = _run({"RANK": "0", "WORLD_SIZE": "1"})
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in line, key
    assert line["impl"] == "reference" and line["metric"] == "pairwise_grav_interactions_per_sec" and line["unit"] == "G/s"
    assert line["higher_is_better"] is True and line["vs_baseline"] is None and line["gpu_launches"] == 0
    assert line["config"]["grid"] == 8 and line["config"]["n_sources_total"] == 200000 and "workload" in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] > 0 and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "G/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the measured wall time of the bounded sample is what ms_per_step reports; the extrapolation is labelled as such
    assert line["ms_per_step"] > 0 and line["ms_per_full_step_extrapolated"] > 0
    # under torchrun the other ranks exit 0 without work and without output
    other = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--gpus", "2")
    assert other.returncode == 0 and other.stdout.strip() == ""
