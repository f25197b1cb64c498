"""The bench.py contract (task section 4) on the CPU: the reference arm runs here on a bounded sample and prints one JSON line with the
required keys; the committed B200 bench lines under profiles/ (fresh-box runs of round 1) carry the keys the driver reads."""
import glob
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
             "e2e", "gpu_launches"}


def _run(extra, env=None):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--ref-sample", "water3x3x3"] + extra,
                         capture_output=True, text=True, timeout=300, env=env, check=True)
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    return json.loads(lines[0])


def test_reference_arm_prints_one_contract_line():
    d = _run([])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["value"] > 0 and d["unit"] == "list-pairs/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_exit_quietly_and_rank0_keeps_all_threads():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
    # torchrun exports OMP_NUM_THREADS=1 when the variable is unset: rank 0 restores the team size of the OpenMP reference build
    env = dict(os.environ, RANK="0", WORLD_SIZE="2", LOCAL_RANK="0", OMP_NUM_THREADS="1", TORCHELASTIC_RUN_ID="t")
    d = _run(["--gpus", "2"], env=env)
    if d["cpu_baseline"]["kind"] == "reference" and len(os.sched_getaffinity(0)) > 1:
        assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert d["n_gpus"] == 2


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_bench_head_n*.json"))))
def test_committed_b200_lines_carry_the_contract_keys(path):
    d = json.loads(open(path).read())
    assert BASE_KEYS | {"roofline", "clocks"} <= set(d), sorted((BASE_KEYS | {"roofline", "clocks"}) - set(d))
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert d["gpu_launches"] > 0 and d["value"] > 0 and "workload" in d["config"]
    assert abs(d["value"] - d["config"]["list_pairs"] / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if d["n_gpus"] == 1:
        assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
        assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["value"] < d["value"]
