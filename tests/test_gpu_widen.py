"""GPU parity of the widened rows: sparse all2all, new2all, and the CLI end
to end on the reference's CI vectors.  Needs a B200: -m gpu."""
import ctypes as C

import numpy as np
import pytest

import oracle_util as ou

pytestmark = pytest.mark.gpu


def _dense_to_rows(tri, N, keep):
    rows = []
    for s in range(N):
        r = tri[ou.tri_cells(s):ou.tri_cells(s) + s]
        cols = [c for c in range(s) if r[c] != 0 and keep(int(r[c]), s, c)]
        rows.append((cols, [int(r[c]) for c in cols]))
    return rows


def _csr_rows(row_ptr, col, val):
    return [(col[int(row_ptr[s]):int(row_ptr[s + 1])].tolist(), val[int(row_ptr[s]):int(row_ptr[s + 1])].tolist())
            for s in range(len(row_ptr) - 1)]


@pytest.mark.parametrize("seed", range(4))
def test_sparse_equals_flat_clique_oracle(libs, oracle, seed):
    """all2all_sp's own algorithm (every node a clique with its raw num_kmers, no W accumulation,
    src/similarity_calculator.cpp:596-638) restated in the oracle must give the same rows."""
    rng = np.random.default_rng(40 + seed)
    N = int(rng.integers(2, 300))
    a, _ = ou.random_trie(rng, N, int(rng.integers(2, 1200)), max_local=int(rng.integers(1, 20)), big_weights=(seed % 2 == 0))
    want = ou.oracle_bruteforce(oracle, N, a)
    cfgs = [dict(), dict(sparse_block_cells=max(N, ou.tri_cells(N) // 7))]  # one block / many row blocks
    for cfg in cfgs:
        with libs.Context(device=0, **cfg) as c:
            v, keep = libs.view_from_arrays(N, a["num_kmers"], a["parent_id"], a["n"], a["l"], a["last"], a["bits"], a["payload_off"], a["payload"])
            c.load_patterns(v, keep)
            rp, col, val, st = c.all2all_sparse()
        assert _csr_rows(rp, col, val) == _dense_to_rows(want, N, lambda v_, s, c_: True)
        assert int(rp[-1]) == int(np.count_nonzero(want))
    # a block of rows (kdbx_all2all_sparse_rows, the unit of the multi-GPU sparse run): those rows, the others empty
    full_rows = _dense_to_rows(want, N, lambda v_, s, c_: True)
    with libs.Context(device=0) as c:
        v, keep = libs.view_from_arrays(N, a["num_kmers"], a["parent_id"], a["n"], a["l"], a["last"], a["bits"], a["payload_off"], a["payload"])
        c.load_patterns(v, keep)
        b, e = N // 3, max(N // 3, N - N // 4)
        rp, col, val, st = c.all2all_sparse(rows=(b, e))
        got = _csr_rows(rp, col, val)
        assert got[b:e] == full_rows[b:e] and all(r == ([], []) for r in got[:b] + got[e:])


def test_sparse_filters_on_device(libs, oracle):
    rng = np.random.default_rng(77)
    N = 200
    a, _ = ou.random_trie(rng, N, 900, max_local=15)
    dense, _ = ou.oracle_all2all(oracle, N, a)
    cnt = rng.integers(1, 5000, size=N).astype(np.uint32)
    cnt[5] = 0
    with libs.Context(device=0, sparse_block_cells=5000) as c:
        v, keep = libs.view_from_arrays(N, a["num_kmers"], a["parent_id"], a["n"], a["l"], a["last"], a["bits"], a["payload_off"], a["payload"])
        c.load_patterns(v, keep)
        lo, hi = int(np.percentile(dense[dense > 0], 30)), int(np.percentile(dense[dense > 0], 90))
        rp, col, val, _ = c.all2all_sparse(min_common=lo, max_common=hi)
        assert _csr_rows(rp, col, val) == _dense_to_rows(dense, N, lambda x, s, t: lo <= x <= hi)
        # metric bounds, evaluated with the reference's arithmetic (uint32 wrap, IEEE double)
        def jac(x, s, t):
            d = (int(cnt[s]) + int(cnt[t]) - int(x)) & 0xFFFFFFFF
            with np.errstate(divide="ignore", invalid="ignore"):
                return np.float64(x) / np.float64(d)
        def cosine(x, s, t):
            with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
                return np.float64(x) / np.sqrt(np.float64((int(cnt[s]) * int(cnt[t])) & 0xFFFFFFFF))
        for name, fn, b in (("jaccard", jac, (0.002, 0.5)), ("cosine", cosine, (0.001, 0.2)),
                            ("min", lambda x, s, t: np.float64(x) / np.float64(min(cnt[s], cnt[t])), (0.01, 1.0)),
                            ("max", lambda x, s, t: np.float64(x) / np.float64(max(cnt[s], cnt[t])), (0.001, 0.05))):
            with np.errstate(divide="ignore", invalid="ignore"):
                rp, col, val, _ = c.all2all_sparse(metric_bounds=[(name, b[0], b[1])], sample_kmers=cnt)
                want = _dense_to_rows(dense, N, lambda x, s, t: bool(fn(x, s, t) >= b[0]) and bool(fn(x, s, t) <= b[1]))
            assert _csr_rows(rp, col, val) == want, name


def _random_db(libs, rng, N, universe_size=4000, k=20):
    universe = np.unique(rng.integers(0, 1 << 40, size=universe_size, dtype=np.uint64))
    sets = []
    for s in range(N):
        base = sets[int(rng.integers(0, s))] if s and rng.random() < 0.7 else universe[rng.random(universe.size) < 0.2]
        sets.append(np.unique(np.concatenate([base[rng.random(base.size) < 0.9], universe[rng.random(universe.size) < 0.02]])))
    return libs.Trie.build([(f"s{i}", x) for i, x in enumerate(sets)], k=k, pinned=False), sets, universe


@pytest.mark.parametrize("seed", range(3))
def test_new2all_equals_set_intersections(libs, seed):
    rng = np.random.default_rng(200 + seed)
    N = int(rng.integers(2, 120))
    t, sets, universe = _random_db(libs, rng, N)
    queries = [np.unique(np.concatenate([universe[rng.random(universe.size) < 0.3], rng.integers(0, 1 << 40, size=50, dtype=np.uint64)]))
               for _ in range(int(rng.integers(1, 9)))]
    queries.append(np.zeros(0, np.uint64))          # empty query
    queries.append(sets[0].copy())                    # a database sample itself
    for cfg in (dict(), dict(query_batch_kmers=1500)):  # one device pass / several
        with libs.Context(device=0, **cfg) as c:
            c.load_patterns(t)
            c.load_hashtables(t)
            out, st = c.new2all_batch(queries)
        want = np.array([[np.intersect1d(q, s, assume_unique=True).size for s in sets] for q in queries], np.uint32)
        assert np.array_equal(out, want)
        assert st.probes == sum(len(q) for q in queries)
        assert st.hits == sum(int(np.isin(q, universe).sum()) for q in queries if len(q)) or st.hits <= st.probes


def test_new2all_matches_oracle_on_reference_database(libs, oracle, ref_fixtures, golden_dbs):
    """The reference-built database of 100 virus genomes and the CI's 65 queries: device result ==
    oracle's one2all row by row."""
    db = ou.ROOT / "tests" / "golden" / "virus.k18.part1.db"
    t = libs.Trie.read_db_full(db)
    import os
    old = os.getcwd()
    os.chdir(ref_fixtures)
    try:
        samples = libs.load_samples("test/virus/seqs.part2.list", k=18)
    finally:
        os.chdir(old)
    assert len(samples) == 65
    with libs.Context(device=0) as c:
        c.load_patterns(t)
        c.load_hashtables(t)
        out, st = c.new2all_batch([k for _, k in samples])
    oracle.oracle_db_read_full.restype = C.c_void_p
    oracle.oracle_one2all.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
    oracle.oracle_one2all.restype = C.c_uint64
    oracle.oracle_db_free.argtypes = [C.c_void_p]
    odb = oracle.oracle_db_read_full(str(db).encode())
    assert odb
    hits = 0
    for q, (_, kmers) in enumerate(samples):
        row = np.zeros(t.num_samples, np.uint32)
        hits += oracle.oracle_one2all(odb, kmers.ctypes.data, kmers.size, row.ctypes.data)
        assert np.array_equal(out[q], row), q
    oracle.oracle_db_free(odb)
    assert st.hits == hits


def test_misuse_is_an_error(libs):
    t = libs.Trie.synth(num_samples=20, num_clusters=1, genome_kmers=2000, seed=1)
    with libs.Context(device=0) as c:
        c.load_patterns(t)
        with pytest.raises(libs.KdbxError, match="no k-mer tables loaded"):
            c.new2all_batch([np.arange(5, dtype=np.uint64)])
        with pytest.raises(libs.KdbxError, match="without its k-mer tables"):
            c.load_hashtables(t)
        with pytest.raises(libs.KdbxError, match="need sample_kmers"):
            c.all2all_sparse(metric_bounds=[("jaccard", 0.0, 1.0)])


# ---- the CLI on the reference's CI sequences (main.yml:71-165, self-hosted.yml:110-260) ------------
def test_cli_virus_sequence(cli, ref_fixtures, tmp_path):
    g = lambda p: ou.read_bytes(ref_fixtures / p)
    cli(ref_fixtures, "build", "test/virus/seqs.part1.list", tmp_path / "parts.db")
    cli(ref_fixtures, "new2all", tmp_path / "parts.db", "test/virus/seqs.part2.list", tmp_path / "n2a.csv")
    assert ou.read_bytes(tmp_path / "n2a.csv") == g("test/virus/k18.n2a.csv")
    cli(ref_fixtures, "new2all", "-sparse", tmp_path / "parts.db", "test/virus/seqs.part2.list", tmp_path / "n2a.s.csv")
    assert ou.read_bytes(tmp_path / "n2a.s.csv") == g("test/virus/k18.n2a.sparse.csv")
    cli(ref_fixtures, "build", "-extend", "-k", "25", "test/virus/seqs.part2.list", tmp_path / "parts.db")
    cli(ref_fixtures, "all2all", tmp_path / "parts.db", tmp_path / "k18.csv")
    assert ou.read_bytes(tmp_path / "k18.csv") == g("test/virus/k18.csv")
    cli(ref_fixtures, "all2all", "-sparse", tmp_path / "parts.db", tmp_path / "k18.s.csv")
    assert ou.read_bytes(tmp_path / "k18.s.csv") == g("test/virus/k18.sparse.csv")
    cli(ref_fixtures, "all2all-sp", tmp_path / "parts.db", tmp_path / "k18.sp.csv")
    assert ou.read_bytes(tmp_path / "k18.sp.csv") == g("test/virus/k18.sparse.csv")
    for m in ("jaccard", "min", "max", "cosine", "mash"):
        cli(ref_fixtures, "distance", m, tmp_path / "k18.csv", tmp_path / f"k18.{m}")
        assert ou.read_bytes(tmp_path / f"k18.{m}") == g(f"test/virus/k18.csv.{m}")
    cli(ref_fixtures, "build", "-f", "0.1", "test/virus/seqs.list", tmp_path / "frac.db")
    cli(ref_fixtures, "all2all", tmp_path / "frac.db", tmp_path / "frac.csv")
    assert ou.read_bytes(tmp_path / "frac.csv") == g("test/virus/k18.frac.csv")
    cli(ref_fixtures, "build", "test/virus/seqs.list", tmp_path / "k18.db")
    cli(ref_fixtures, "new2all", tmp_path / "k18.db", "test/virus/seqs.list", tmp_path / "itself.csv")
    assert ou.read_bytes(tmp_path / "itself.csv") == g("test/virus/k18.n2a.itself.csv")
    cli(ref_fixtures, "new2all", "-multisample-fasta", tmp_path / "k18.db", "test/virus/multi.list", tmp_path / "itself.m.csv")
    assert ou.read_bytes(tmp_path / "itself.m.csv") == g("test/virus/k18.n2a.itself.csv")


def test_cli_synth_sequence(cli, ref_fixtures, tmp_path):
    g = lambda p: ou.read_bytes(ref_fixtures / p)
    db = tmp_path / "synth.db"
    cli(ref_fixtures, "build", "-multisample-fasta", "-k", "21", "test/synth/synth.list", db)
    cli(ref_fixtures, "all2all", db, tmp_path / "a2a")
    assert ou.read_bytes(tmp_path / "a2a") == g("test/synth/a2a")
    cli(ref_fixtures, "all2all", "-sparse", db, tmp_path / "a2a-sparse")
    assert ou.read_bytes(tmp_path / "a2a-sparse") == g("test/synth/a2a-sparse")
    cli(ref_fixtures, "all2all", "-sparse", "-max", "39", "-min", "num-kmers:31", db, tmp_path / "mm")
    assert ou.read_bytes(tmp_path / "mm") == g("test/synth/a2a.sparse.above-below")
    cli(ref_fixtures, "all2all-sp", db, tmp_path / "a2a-sp")
    assert ou.read_bytes(tmp_path / "a2a-sp") == g("test/synth/a2a-sparse")
    cli(ref_fixtures, "all2all-sp", "-max", "39", "-min", "num-kmers:31", db, tmp_path / "sp-mm")
    assert ou.read_bytes(tmp_path / "sp-mm") == g("test/synth/a2a.sparse.above-below")
    cli(ref_fixtures, "new2all", "-multisample-fasta", db, "test/synth/synth.list", tmp_path / "n2a")
    assert ou.read_bytes(tmp_path / "n2a") == g("test/synth/n2a")
    cli(ref_fixtures, "new2all", "-multisample-fasta", "-sparse", db, "test/synth/synth.list", tmp_path / "n2a-s")
    assert ou.read_bytes(tmp_path / "n2a-s") == g("test/synth/n2a-sparse")
    cli(ref_fixtures, "new2all", "-multisample-fasta", "-sparse", "-max", "69", "-min", "num-kmers:21", db, "test/synth/synth.list", tmp_path / "n2a-mm")
    assert ou.read_bytes(tmp_path / "n2a-mm") == g("test/synth/n2a.sparse.above-below")
    # a log-based metric bound is applied on the host side of all2all-sp: same rows as distance -sparse
    cli(ref_fixtures, "all2all-sp", "-min", "ani:0.95", db, tmp_path / "sp-ani")
    cli(ref_fixtures, "all2all", "-sparse", "-min", "ani:0.95", db, tmp_path / "dense-ani")
    assert ou.read_bytes(tmp_path / "sp-ani") == ou.read_bytes(tmp_path / "dense-ani")


def test_cli_protein_alphabet(cli, ref_fixtures, tmp_path):
    cli(ref_fixtures, "build", "-k", "8", "-multisample-fasta", "-alphabet", "aa12_mmseqs", "test/protein/aa_100x1000.fasta", tmp_path / "aa.db")
    cli(ref_fixtures, "all2all", tmp_path / "aa.db", tmp_path / "aa.a2a")
    assert ou.read_bytes(tmp_path / "aa.a2a") == ou.read_bytes(ref_fixtures / "test/protein/aa12_mmseqs.a2a")


def test_cli_multi_gpu_flag(cli, libs, ref_fixtures, golden_dbs, tmp_path):
    k, _ = libs.load()
    n = k.kdbx_device_count()
    if n < 2:
        r = cli(ref_fixtures, "all2all", "-gpus", "2", golden_dbs["virus.k18"][0], tmp_path / "x.csv", check=False)
        assert r.returncode != 0 and "only 1 B200" in r.stderr
        return
    cli(ref_fixtures, "all2all", "-gpus", str(min(n, 4)), golden_dbs["virus.k18"][0], tmp_path / "x.csv")
    assert ou.read_bytes(tmp_path / "x.csv") == ou.read_bytes(golden_dbs["virus.k18"][1])
    # all2all-sp over row blocks on several devices, rows concatenated: the reference's sparse golden file
    cli(ref_fixtures, "all2all-sp", "-gpus", str(min(n, 4)), golden_dbs["virus.k18"][0], tmp_path / "xs.csv")
    assert ou.read_bytes(tmp_path / "xs.csv") == ou.read_bytes(golden_dbs["virus.k18"][2])


def test_dense_csv_formatted_on_the_device(libs, oracle, cli, golden_dbs, tmp_path):
    """kdbx_csv_dense_rows: the cells' decimal text from the matrix in HBM equals the host formatter's, for values of
    every length (0 .. 2^32-1), for row sub-ranges, and through the CLI (default) against -host-csv and the golden file."""
    rng = np.random.default_rng(31)
    N = 333
    a, _ = ou.random_trie(rng, N, 3000, max_local=30, big_weights=True)
    want, _ = ou.oracle_all2all(oracle, N, a)
    with libs.Context(device=0) as c:
        v, keep = libs.view_from_arrays(N, a["num_kmers"], a["parent_id"], a["n"], a["l"], a["last"], a["bits"], a["payload_off"], a["payload"])
        c.load_patterns(v, keep)
        tri, _ = c.all2all_dense()
        assert np.array_equal(tri, want)
        for r0, r1 in ((0, N), (0, 1), (1, 2), (17, 200), (N - 1, N)):
            text, off = c.csv_dense_rows(r0, r1)
            for s in range(r0, r1):
                row = want[ou.tri_cells(s):ou.tri_cells(s) + s]
                expect = "".join(f"{int(x)}," for x in row).encode()
                assert text[int(off[s - r0]):int(off[s - r0 + 1])] == expect, (r0, r1, s)
        assert {len(str(int(x))) for x in want} >= {1, 10}   # the input does hold one- and ten-digit numbers
        rows, _ = c.all2all_dense_rows(100, 300)
        text, off = c.csv_dense_rows(150, 160)
        assert text[:int(off[1])] == "".join(f"{int(x)}," for x in want[ou.tri_cells(150):ou.tri_cells(150) + 150]).encode()
        with pytest.raises(libs.KdbxError, match="not in the resident block"):
            c.csv_dense_rows(50, 120)
    db, dense, _ = golden_dbs["virus.k18"]
    cli(tmp_path, "all2all", db, tmp_path / "dev.csv")
    cli(tmp_path, "all2all", "-host-csv", db, tmp_path / "host.csv")
    assert ou.read_bytes(tmp_path / "dev.csv") == ou.read_bytes(dense) == ou.read_bytes(tmp_path / "host.csv")


def test_distance_on_the_device(libs, oracle, cli, ref_fixtures, tmp_path):
    """kdbx_distance_dense_rows / `distance -device`: jaccard, min, max, cosine and num-kmers with six decimals from the
    matrix in HBM — the reference's golden files for the virus table, the host formatter on a generated one, and the
    arithmetic restated with numpy (IEEE division and square root are correctly rounded everywhere)."""
    for m in ("jaccard", "min", "max", "cosine"):
        cli(ref_fixtures, "distance", m, "-device", "test/virus/k18.csv", tmp_path / f"dev.{m}")
        assert ou.read_bytes(tmp_path / f"dev.{m}") == ou.read_bytes(ref_fixtures / f"test/virus/k18.csv.{m}")
    r = cli(ref_fixtures, "distance", "mash", "-device", "test/virus/k18.csv", tmp_path / "x", check=False)
    assert r.returncode != 0 and "logarithm-based measures run on the host" in r.stderr
    r = cli(ref_fixtures, "distance", "jaccard", "-device", "test/virus/k18.n2a.csv", tmp_path / "x", check=False)
    assert r.returncode != 0 and "dense triangular table" in r.stderr
    t = libs.Trie.synth(num_samples=300, num_clusters=3, genome_kmers=50000, seed=17)
    N = 300
    cnt = np.array(t.sample_kmer_counts(), np.uint32)
    want, _ = ou.oracle_all2all(oracle, N, t.arrays())

    def f6(v):
        if v == 0:
            return "0"
        x = int(v * 1000000.0 + 0.5)
        return f"{x // 1000000}.{x % 1000000:06d}"
    with libs.Context(device=0) as c:
        c.load_patterns(t)
        c.all2all_dense()
        for m in ("jaccard", "min", "max", "cosine", "num-kmers"):
            text, off = c.distance_dense_rows(m, cnt, 0, N)
            for s in (0, 1, 2, 57, 299):
                row = want[ou.tri_cells(s):ou.tri_cells(s) + s].astype(np.float64)
                a, b = np.float64(cnt[s]), cnt[:s].astype(np.float64)
                with np.errstate(invalid="ignore", divide="ignore"):
                    v = {"jaccard": row / ((cnt[s] + cnt[:s] - want[ou.tri_cells(s):ou.tri_cells(s) + s]).astype(np.uint32)).astype(np.float64),
                         "min": row / np.minimum(a, b), "max": row / np.maximum(a, b),
                         "cosine": row / np.sqrt((cnt[s] * cnt[:s]).astype(np.uint32).astype(np.float64)), "num-kmers": row}[m]
                assert text[int(off[s]):int(off[s + 1])].decode() == "".join(f6(float(x)) + "," for x in v), (m, s)
        c.stage_matrix(want, N)
        text2, off2 = c.distance_dense_rows("jaccard", cnt, 10, 20)
        text1, off1 = c.distance_dense_rows("jaccard", cnt, 0, N)
        assert text2 == text1[int(off1[10]):int(off1[20])]
        bad = cnt.copy(); bad[5] = 0
        with pytest.raises(libs.KdbxError, match="has no k-mers"):
            c.distance_dense_rows("jaccard", bad, 0, N)


# ---- all2all-parts / db2db_sp (src/console_all2all_parts.cpp, src/similarity_calculator.cpp:1225-1540) ----
def _split_virus_lists(ref_fixtures, tmp_path, cuts):
    lines = (ref_fixtures / "test/virus/seqs.list").read_text().split()
    out = []
    for i, (b, e) in enumerate(zip([0] + cuts, cuts + [len(lines)])):
        f = tmp_path / f"slice{i}.list"
        f.write_text("\n".join(lines[b:e]) + "\n")
        out.append(f)
    return out


@pytest.mark.parametrize("cuts", [[100], [60, 120], [1, 164]])
def test_cli_all2all_parts(cli, ref_fixtures, tmp_path, cuts):
    """The reference's CI check for the mode (self-hosted.yml:357-363): parts built separately, all2all-parts over the list
    == k18.sparse.csv; with output filters == what the reference binary wrote for the same three parts (tests/golden)."""
    dbs = []
    for i, lst in enumerate(_split_virus_lists(ref_fixtures, tmp_path, cuts)):
        dbs.append(tmp_path / f"part{i}.db")
        cli(ref_fixtures, "build", lst, dbs[-1])
    if cuts == [100]:
        dbs[0] = ou.ROOT / "tests" / "golden" / "virus.k18.part1.db"   # as the reference built it
    (tmp_path / "db.list").write_text("\n".join(map(str, dbs)) + "\n")
    r = cli(ref_fixtures, "all2all-parts", tmp_path / "db.list", tmp_path / "parts.csv")
    assert ou.read_bytes(tmp_path / "parts.csv") == ou.read_bytes(ref_fixtures / "test/virus/k18.sparse.csv")
    assert "No. saved pairs: 13530" in r.stderr
    # -buffer <mb> bounds the parts kept on the device: with 9 MB none stays, every cell stages its column part again
    cli(ref_fixtures, "all2all-parts", "-buffer", "9", tmp_path / "db.list", tmp_path / "parts.small.csv")
    assert ou.read_bytes(tmp_path / "parts.small.csv") == ou.read_bytes(ref_fixtures / "test/virus/k18.sparse.csv")
    cli(ref_fixtures, "all2all-parts", "-min", "jaccard:0.985", "-max", "num-kmers:29700", "-min", "ani:0.9995", tmp_path / "db.list", tmp_path / "f.csv")
    assert ou.read_bytes(tmp_path / "f.csv") == ou.read_bytes(ou.ROOT / "tests" / "golden" / "virus.k18.parts.filtered.csv")


def test_cli_all2all_parts_on_two_gpus(libs, cli, ref_fixtures, tmp_path):
    """-gpus 2: the grid rows are dealt to two devices and written in order."""
    k, _ = libs.load()
    if k.kdbx_device_count() < 2:
        pytest.skip("needs two GPUs")
    dbs = []
    for i, lst in enumerate(_split_virus_lists(ref_fixtures, tmp_path, [40, 80, 120])):
        dbs.append(tmp_path / f"part{i}.db")
        cli(ref_fixtures, "build", lst, dbs[-1])
    (tmp_path / "db.list").write_text("\n".join(map(str, dbs)) + "\n")
    cli(ref_fixtures, "all2all-parts", "-gpus", "2", tmp_path / "db.list", tmp_path / "parts.csv")
    assert ou.read_bytes(tmp_path / "parts.csv") == ou.read_bytes(ref_fixtures / "test/virus/k18.sparse.csv")


def test_db2db_cells_against_oracle(libs, oracle, cli, ref_fixtures, tmp_path):
    """kdbx_db2db_sparse through the C ABI against the oracle's db2db restatement, cell for cell: both orientations of
    two databases of different sizes, one row block and many, with device-side filters."""
    oracle.oracle_db2db.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    oracle.oracle_db2db.restype = C.c_uint64
    oracle.oracle_db_read_full.restype = C.c_void_p
    oracle.oracle_db_read_full.argtypes = [C.c_char_p]
    oracle.oracle_db_free.argtypes = [C.c_void_p]
    p1 = ou.ROOT / "tests" / "golden" / "virus.k18.part1.db"
    p2 = tmp_path / "part2.db"
    cli(ref_fixtures, "build", "test/virus/seqs.part2.list", p2)
    t = {1: libs.Trie.read_db_full(p1), 2: libs.Trie.read_db_full(p2)}
    o = {1: oracle.oracle_db_read_full(str(p1).encode()), 2: oracle.oracle_db_read_full(str(p2).encode())}
    assert o[1] and o[2]
    for r, c in ((2, 1), (1, 2)):
        N1, N2 = t[r].num_samples, t[c].num_samples
        want = np.zeros(N1 * N2, np.uint32)
        upd = oracle.oracle_db2db(o[r], o[c], want.ctypes.data)
        assert upd != 2**64 - 1
        want = want.reshape(N1, N2)
        for cfg in (dict(), dict(sparse_block_cells=7 * N2)):
            with libs.Context(device=0, **cfg) as rows, libs.Context(device=0) as cols:
                rows.load_patterns(t[r]); rows.load_hashtables(t[r])
                cols.load_patterns(t[c]); cols.load_hashtables(t[c])
                for _ in range(2):   # (the second call finds both databases prepared)
                    rp, col, val, st = rows.db2db_sparse(cols)
                    got = np.zeros((N1, N2), np.uint32)
                    for s in range(N1):
                        b, e = int(rp[s]), int(rp[s + 1])
                        assert np.all(np.diff(col[b:e].astype(np.int64)) > 0)
                        got[s, col[b:e]] = val[b:e]
                    assert np.array_equal(got, want)
                    assert int(rp[-1]) == int(np.count_nonzero(want)) and st.updates == upd and st.hits > 0 and st.probes >= st.hits
                # filters evaluated on the device: k-mer count bounds and a jaccard bound with both databases' counts
                ca, cb = t[r].sample_kmer_counts().astype(np.uint32), t[c].sample_kmer_counts().astype(np.uint32)
                rp, col, val, st = rows.db2db_sparse(cols, min_common=20000, max_common=29700, metric_bounds=[("jaccard", 0.9, 0.99)],
                                                     row_kmers=ca, col_kmers=cb)
                keep = (want >= 20000) & (want <= 29700)
                with np.errstate(divide="ignore", invalid="ignore"):
                    jac = want.astype(np.float64) / (ca[:, None] + cb[None, :] - want).astype(np.uint32).astype(np.float64)
                keep &= (jac >= 0.9) & (jac <= 0.99) & (want != 0)
                got = np.zeros((N1, N2), bool)
                for s in range(N1):
                    got[s, col[int(rp[s]):int(rp[s + 1])]] = True
                assert np.array_equal(got, keep) and keep.any() and not keep.all()
    # misuse
    with libs.Context(device=0) as rows, libs.Context(device=0) as cols:
        rows.load_patterns(t[1]); rows.load_hashtables(t[1])
        cols.load_patterns(t[2])
        with pytest.raises(libs.KdbxError, match="no k-mer tables loaded"):
            rows.db2db_sparse(cols)
        with pytest.raises(libs.KdbxError, match="two contexts"):
            rows.db2db_sparse(rows)
    for d in o.values():
        oracle.oracle_db_free(d)


def test_cli_sample_rows(cli, ref_fixtures, golden_dbs, tmp_path):
    """-sample-rows <criterion>:<count> in all2all-sp and all2all-parts (host/csv_out.h::RowSampler over the rows the device
    delivers): the bytes the reference binary wrote for the same options (tests/golden/make_golden.sh); the parts mode
    writes the same table for the genomes split into parts.  (The selection itself is also pinned on CPU:
    tests/test_host.py::test_sparse_table_with_sample_rows_equals_reference_output.)"""
    G = ou.ROOT / "tests" / "golden"
    db = golden_dbs["virus.k18"][0]
    dbs = []
    for i, lst in enumerate(_split_virus_lists(ref_fixtures, tmp_path, [60, 120])):
        dbs.append(tmp_path / f"part{i}.db")
        cli(ref_fixtures, "build", lst, dbs[-1])
    (tmp_path / "db.list").write_text("\n".join(map(str, dbs)) + "\n")
    for words, golden in ((["-sample-rows", "jaccard:3"], "virus.k18.sampled.jaccard_3.csv"),
                          (["-sample-rows", "mash-query:4"], "virus.k18.sampled.mash-query_4.csv"),
                          (["-min", "jaccard:0.99", "-max", "num-kmers:29800", "-sample-rows", "cosine:2"], "virus.k18.sampled.filtered.cosine_2.csv")):
        want = ou.read_bytes(G / golden)
        pairs = sum(len([x for x in ln.split(b",")[2:] if x]) for ln in want.splitlines()[2:])
        r = cli(ref_fixtures, "all2all-sp", *words, db, tmp_path / "sp.csv")
        assert ou.read_bytes(tmp_path / "sp.csv") == want, words
        assert f"No. saved pairs: {pairs}" in r.stderr
        r = cli(ref_fixtures, "all2all-parts", *words, tmp_path / "db.list", tmp_path / "parts.csv")
        assert ou.read_bytes(tmp_path / "parts.csv") == want, words
        assert f"No. saved pairs: {pairs}" in r.stderr


def test_cli_one2all(cli, ref_fixtures, tmp_path):
    """The reference's CI step for the mode (.github/workflows/main.yml:156-160): build -k 25 -f 0.1 from the first 100 genomes,
    one2all with one more genome given without its extension, cmp with test/virus/MT159713.csv."""
    cli(ref_fixtures, "build", "-k", "25", "-f", "0.1", "test/virus/seqs.part1.list", tmp_path / "k25.db")
    cli(ref_fixtures, "one2all", tmp_path / "k25.db", "./test/virus/data/MT159713", tmp_path / "MT159713.csv")
    assert ou.read_bytes(tmp_path / "MT159713.csv") == ou.read_bytes(ref_fixtures / "test/virus/MT159713.csv")
    r = cli(ref_fixtures, "one2all", tmp_path / "k25.db", "./test/virus/data/no-such-genome", tmp_path / "x.csv", check=False)
    assert r.returncode != 0 and "Cannot open sample file" in r.stderr


def test_cli_minhash_inputs(cli, ref_fixtures, tmp_path):
    """-from-minhash through the device paths: the CI step `minhash -f 0.1; build -from-minhash; all2all` == k18.frac.csv
    (.github/workflows/main.yml:143-148) with the k-mer sets going to kdbx_builder_add_kmers; new2all / one2all
    -from-minhash (kdbx_new2all_batch) give the tables of the same queries read from FASTA against that database."""
    import shutil
    work = tmp_path / "w"
    shutil.copytree(ref_fixtures / "test" / "virus", work / "test" / "virus")
    cli(work, "minhash", "-f", "0.1", "test/virus/seqs.list")
    cli(work, "build", "-from-minhash", "test/virus/seqs.list", tmp_path / "mh.db")
    cli(work, "all2all", tmp_path / "mh.db", tmp_path / "mh.csv")
    assert ou.read_bytes(tmp_path / "mh.csv") == ou.read_bytes(ref_fixtures / "test/virus/k18.frac.csv")
    cli(work, "build", "-from-minhash", "test/virus/seqs.part1.list", tmp_path / "p1.db")
    for extra in ([], ["-sparse"]):
        cli(work, "new2all", *extra, tmp_path / "p1.db", "test/virus/seqs.part2.list", tmp_path / "fasta.csv")
        cli(work, "new2all", "-from-minhash", *extra, tmp_path / "p1.db", "test/virus/seqs.part2.list", tmp_path / "minhash.csv")
        assert ou.read_bytes(tmp_path / "minhash.csv") == ou.read_bytes(tmp_path / "fasta.csv")
    cli(work, "one2all", tmp_path / "p1.db", "./test/virus/data/MT159713", tmp_path / "o.fasta.csv")
    cli(work, "one2all", "-from-minhash", tmp_path / "p1.db", "./test/virus/data/MT159713", tmp_path / "o.minhash.csv")
    assert ou.read_bytes(tmp_path / "o.minhash.csv") == ou.read_bytes(tmp_path / "o.fasta.csv")
