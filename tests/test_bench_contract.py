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
    # under torchrun only rank 0 works and prints; the other ranks exit 0 without output
    out = _run(["--impl", "reference", "--gpus", "2", *small], env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert out.strip() == ""
