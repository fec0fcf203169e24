"""The oracle (oracle/kdb_oracle.c) against the reference's golden vectors, a brute-force second
opinion, and the unmodified reference binary.  CPU only."""
import subprocess

import numpy as np
import pytest

import oracle_util as ou


@pytest.mark.parametrize("name", ["virus.k18", "virus.k18.f01", "virus.k24", "synth.k21"])
def test_oracle_reproduces_golden_csv(oracle, golden_dbs, tmp_path, name):
    db, dense, sparse = golden_dbs[name]
    out = tmp_path / "o.csv"
    U = oracle.oracle_all2all_file(str(db).encode(), str(out).encode(), 0)
    assert U != 2**64 - 1
    assert ou.read_bytes(out) == ou.read_bytes(dense)
    if sparse is not None:
        oracle.oracle_all2all_file(str(db).encode(), str(out).encode(), 1)
        assert ou.read_bytes(out) == ou.read_bytes(sparse)


def test_oracle_update_counts_match_survey(oracle, golden_dbs):
    # U of the reference's own addition counter on these inputs
    expect = {"virus.k18": 1945823, "virus.k24": 2064457, "virus.k18.f01": 1284890, "synth.k21": 10}
    for name, u in expect.items():
        assert oracle.oracle_all2all_file(str(golden_dbs[name][0]).encode(), None, 0) == u


@pytest.mark.parametrize("seed", range(6))
def test_oracle_equals_bruteforce_on_random_tries(oracle, seed):
    rng = np.random.default_rng(seed)
    N = int(rng.integers(2, 60))
    a, _ = ou.random_trie(rng, N, int(rng.integers(2, 200)), big_weights=(seed % 2 == 1))
    tri, U = ou.oracle_all2all(oracle, N, a)
    brute = ou.oracle_bruteforce(oracle, N, a)
    assert np.array_equal(tri, brute)
    nn, ll = a["n"].astype(np.int64), a["l"].astype(np.int64)
    assert U == int((ll * (2 * nn - ll - 1) // 2).sum())


def test_oracle_decode_matches_python_lists(oracle):
    rng = np.random.default_rng(5)
    a, lists = ou.random_trie(rng, 500, 300, max_local=40)
    for p in range(1, len(lists)):
        out = np.zeros(len(lists[p]), np.uint32)
        w = a["payload"][int(a["payload_off"][p]):]
        oracle.oracle_decode_local(w.ctypes.data, len(lists[p]), int(a["last"][p]), out.ctypes.data)
        assert out.tolist() == lists[p]


def test_oracle_equals_reference_binary_on_generated_db(oracle, libs, ref_bin, tmp_path):
    if ref_bin is None:
        pytest.skip("reference binary not built (oracle/_ref)")
    t = libs.Trie.synth(num_samples=120, num_clusters=3, genome_kmers=60000, seed=11)
    db = tmp_path / "g.db"
    t.write_db(db)
    subprocess.run([str(ref_bin), "all2all", "-t", "4", str(db), str(tmp_path / "ref.csv")], check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    oracle.oracle_all2all_file(str(db).encode(), str(tmp_path / "o.csv").encode(), 0)
    assert ou.read_bytes(tmp_path / "o.csv") == ou.read_bytes(tmp_path / "ref.csv")


@pytest.mark.parametrize("seed", range(6))
def test_regrouped_formulation_gives_the_same_bits(oracle, libs, golden_dbs, seed):
    """summing the weights of a node's descendants per row before touching the matrix is exact
    (uint32 wrap-around included) and, on tries with deep sharing, needs fewer operations than U."""
    rng = np.random.default_rng(200 + seed)
    N = int(rng.integers(2, 80))
    a, _ = ou.random_trie(rng, N, int(rng.integers(2, 400)), max_local=int(rng.integers(1, 12)), big_weights=(seed % 2 == 0))
    tri, U = ou.oracle_all2all(oracle, N, a)
    re, ops = ou.oracle_regrouped(oracle, N, a)
    assert np.array_equal(tri, re)
    if seed == 0:
        for name in ("virus.k18", "synth.k21"):
            t = libs.Trie.read_db(golden_dbs[name][0])
            tri, U = ou.oracle_all2all(oracle, t.num_samples, t.arrays())
            re, ops = ou.oracle_regrouped(oracle, t.num_samples, t.arrays())
            assert np.array_equal(tri, re)
        t = libs.Trie.synth(num_samples=120, num_clusters=2, genome_kmers=30000, seed=3)
        tri, U = ou.oracle_all2all(oracle, 120, t.arrays())
        re, ops = ou.oracle_regrouped(oracle, 120, t.arrays())
        assert np.array_equal(tri, re)
        assert ops < U   # the generated cluster tries share deeply: the regrouped form does less work


@pytest.mark.parametrize("seed", range(6))
def test_boundary_formulation_gives_the_same_bits(oracle, libs, golden_dbs, seed):
    """kmer-db_b200/csrc/diff.cuh: lists as run boundaries, +w / -w into a difference matrix, prefix sums along the
    rows — exact modulo 2^32, and fewer updates than U wherever the lists hold runs of consecutive ids."""
    rng = np.random.default_rng(300 + seed)
    N = int(rng.integers(2, 90))
    a, _ = ou.random_trie(rng, N, int(rng.integers(2, 400)), max_local=int(rng.integers(1, 14)), big_weights=(seed % 2 == 0),
                          dense_lists=(seed % 3 == 0))
    tri, U = ou.oracle_all2all(oracle, N, a)
    bd, U2, phys = ou.oracle_boundary(oracle, N, a)
    assert np.array_equal(tri, bd) and U2 == U
    if seed == 0:
        for name in ("virus.k18", "virus.k24", "synth.k21"):
            t = libs.Trie.read_db(golden_dbs[name][0])
            tri, U = ou.oracle_all2all(oracle, t.num_samples, t.arrays())
            bd, U2, phys = ou.oracle_boundary(oracle, t.num_samples, t.arrays())
            assert np.array_equal(tri, bd) and U2 == U
        t = libs.Trie.synth(num_samples=120, num_clusters=2, genome_kmers=30000, seed=3)
        tri, U = ou.oracle_all2all(oracle, 120, t.arrays())
        bd, U2, phys = ou.oracle_boundary(oracle, 120, t.arrays())
        assert np.array_equal(tri, bd) and U2 == U
        assert phys < 0.7 * U   # cluster members sit side by side in sample order: their lists are mostly runs
