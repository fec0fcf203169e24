"""bench.py's reference arm runs on host cores only, so its JSON contract can be checked without a GPU."""
import json
import os
import subprocess
import sys

import pytest

import oracle_util as ou


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([sys.executable, str(ou.ROOT / "bench.py"), *args], capture_output=True, text=True, env=e, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def test_reference_arm_prints_one_contract_line(libs, ref_bin, tmp_path):
    if ref_bin is None:
        pytest.skip("reference binary not built (oracle/_ref)")
    small = ["--samples", "40", "--clusters", "4", "--genome-kmers", "20000", "--steps", "2", "--warmup", "1", "--cache-dir", str(tmp_path)]
    out = _run(["--impl", "reference", "--gpus", "1", *small])
    lines = [ln for ln in out.splitlines() if ln.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"].startswith("k-mer-pair updates/sec") and j["unit"] == "updates/s"
    assert j["higher_is_better"] is True and j["steps"] == 2 and j["warmup"] == 1 and j["value"] > 0
    assert j["cpu_baseline"]["kind"] == "reference" and j["cpu_baseline"]["cores"] >= 1 and j["cpu_baseline"]["value"] == j["value"]
    assert j["e2e"] == {"value": j["value"], "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert j["vs_baseline"] is None and "workload" in j["config"]
    # the arm times the WHOLE workload (same .db file as the GPU arm), keeps its CSV for the GPU arm's cmp, and loads
    # none of this repository's libraries: the database comes from the stand-alone generator
    assert j["config"]["num_samples"] == 40 and j["config"]["updates_per_step"] > 0
    assert "whole workload" in j["cpu_baseline"]["sample"] and j["same_config"] is True
    assert j["full_config_run"]["updates"] == j["config"]["updates_per_step"] == j["arm"]["updates_per_step"]
    db = tmp_path / j["config"]["database"]
    assert db.exists() and (tmp_path / (j["config"]["database"] + ".ref.csv")).exists()
    src = (ou.ROOT / "bench.py").read_text()
    ref_fn = src[src.index("def reference_arm("):src.index("def _own_cells(")]
    assert "kdbx." not in ref_fn and "import kdbx" not in ref_fn
    # N > 1, weak scaling: the bounded sample is the one-GPU database, and the line says so
    out = _run(["--impl", "reference", "--gpus", "2", *small], env={"RANK": "0", "LOCAL_RANK": "0", "WORLD_SIZE": "2"})
    j2 = json.loads(out.strip().splitlines()[-1])
    assert j2["config"]["num_samples"] == 80 and j2["config"]["sample_of_workload"]["num_samples"] == 40
    assert "1/2 of the 2-GPU workload" in j2["cpu_baseline"]["sample"]
    # runs of the whole database that do not fit the arm's budget: the whole database still runs once (its CSV is kept for
    # the GPU arm's cmp), the steps are the same shape at a fraction of the genome length, and the line carries both rates
    out = _run(["--impl", "reference", "--gpus", "1", "--ref-budget-s", "0", *small])
    j3 = json.loads(out.strip().splitlines()[-1])
    assert j3["same_config"] is False and j3["steps"] == 2 and j3["warmup"] == 1 and len(j3["arm"]["seconds_per_step"]) == 2
    assert j3["config"]["num_samples"] == 40 and j3["config"]["database"] == j["config"]["database"]
    assert j3["config"]["step_sample"]["num_samples"] == 40 and j3["config"]["step_sample"]["database"] == j3["arm"]["step_database"] != j["config"]["database"]
    assert 0 < j3["arm"]["updates_per_step"] == j3["config"]["step_sample"]["updates_per_step"] < j["config"]["updates_per_step"]
    assert j3["full_config_run"]["updates"] == j["config"]["updates_per_step"] and j3["full_config_run"]["value"] > 0
    assert j3["cpu_baseline"]["value"] == j3["value"] == j3["e2e"]["value"] and "full_config_run" in j3["cpu_baseline"]["sample"]
    assert abs(j3["steps_rate_over_full_config_rate"] - j3["value"] / j3["full_config_run"]["value"]) < 1e-9
    # under torchrun only rank 0 works and prints; the other ranks exit 0 without output
    out = _run(["--impl", "reference", "--gpus", "2", *small], env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert out.strip() == ""
