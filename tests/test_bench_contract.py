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
    assert "whole workload" in j["cpu_baseline"]["sample"]
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
    # under torchrun only rank 0 works and prints; the other ranks exit 0 without output
    out = _run(["--impl", "reference", "--gpus", "2", *small], env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert out.strip() == ""
