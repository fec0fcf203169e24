import gzip
import os
import shutil
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "kmer-db_b200"
sys.path.insert(0, str(PKG))
sys.path.insert(0, str(ROOT / "tests"))
GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


def _make(target_dir, *args):
    subprocess.run(["make", "-C", str(target_dir), *args], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)


@pytest.fixture(scope="session")
def libs():
    """Built product libraries (built on demand here; prebuilt .so files travel to the GPU box)."""
    if not (PKG / "lib" / "libkdbx.so").exists() or not (PKG / "lib" / "libkdbx_host.so").exists():
        _make(PKG)
    import kdbx
    kdbx.load()
    return kdbx


@pytest.fixture(scope="session")
def oracle():
    """ctypes handle of the C restatement (oracle/kdb_oracle.c) — the checker."""
    import oracle_util
    return oracle_util.load_oracle()


@pytest.fixture(scope="session")
def ref_bin():
    """The unmodified reference binary, if it was built (oracle/_ref/kmer-db)."""
    p = ROOT / "oracle" / "_ref" / "kmer-db"
    if not p.exists() and Path("/root/reference/src").exists():
        subprocess.run([str(ROOT / "oracle" / "build_ref.sh")], check=True, stdout=subprocess.DEVNULL)
    return p if p.exists() else None


@pytest.fixture(scope="session")
def golden_dbs(tmp_path_factory):
    """name -> (db path, dense csv path, sparse csv path or None)"""
    d = tmp_path_factory.mktemp("golden")
    k24 = d / "virus.k24.db"
    with gzip.open(GOLDEN / "virus.k24.db.gz", "rb") as src, open(k24, "wb") as dst:
        shutil.copyfileobj(src, dst)
    return {
        "virus.k18": (GOLDEN / "virus.k18.db", GOLDEN / "virus.k18.csv", GOLDEN / "virus.k18.sparse.csv"),
        "virus.k18.f01": (GOLDEN / "virus.k18.f01.db", GOLDEN / "virus.k18.f01.csv", None),
        "virus.k24": (k24, GOLDEN / "virus.k24.csv", None),
        "synth.k21": (GOLDEN / "synth.k21.db", GOLDEN / "synth.k21.csv", GOLDEN / "synth.k21.sparse.csv"),
    }


@pytest.fixture(scope="session")
def ref_fixtures(tmp_path_factory):
    """The reference's own test inputs and golden outputs (tests/golden/reference_fixtures.tar.xz,
    made by tests/golden/make_golden.sh), unpacked so that the list files' relative paths
    (./test/virus/data/...) resolve from the returned directory."""
    import tarfile
    d = tmp_path_factory.mktemp("ref")
    with tarfile.open(GOLDEN / "reference_fixtures.tar.xz", "r:xz") as tf:
        tf.extractall(d, filter="data")
    return d


@pytest.fixture(scope="session")
def cli(libs):
    """Runs kmer-db-b200 from a working directory; returns the CompletedProcess."""
    exe = PKG / "bin" / "kmer-db-b200"
    if not exe.exists():
        _make(PKG)

    def run(cwd, *args, check=True):
        r = subprocess.run([str(exe), *map(str, args)], cwd=str(cwd), capture_output=True, text=True)
        if check and r.returncode != 0:
            raise AssertionError(f"kmer-db-b200 {' '.join(map(str, args))} failed ({r.returncode}): {r.stderr[-2000:]}")
        return r
    return run
