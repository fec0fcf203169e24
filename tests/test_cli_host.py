"""Host-side modes of the CLI (build -host-build, distance) and the host builder against the reference's own
CI vectors (.github/workflows/main.yml:45-165, self-hosted.yml:91-426).  CPU only: databases we
build are checked by running the ORACLE's all2all / new2all on them and comparing with the
reference's golden CSVs."""
import ctypes as C
import shutil

import numpy as np
import pytest

import oracle_util as ou


def _oracle_all2all_csv(oracle, db, out, sparse=False):
    assert oracle.oracle_all2all_file(str(db).encode(), str(out).encode(), 1 if sparse else 0) != 2**64 - 1
    return ou.read_bytes(out)


def _oracle_new2all_csv(oracle, db, lst, out, multi=False, sparse=False, cwd=None):
    import os
    old = os.getcwd()
    os.chdir(cwd)
    try:
        n = oracle.oracle_new2all_file(str(db).encode(), str(lst).encode(), 1 if multi else 0, str(out).encode(), 1 if sparse else 0)
    finally:
        os.chdir(old)
    assert n >= 0
    return ou.read_bytes(out)


@pytest.mark.parametrize("args,golden", [
    (["test/virus/seqs.list"], "test/virus/k18.csv"),
    (["-multisample-fasta", "test/virus/multi.list"], "test/virus/k18.csv"),
    (["-multisample-fasta", "test/virus/multi.split.list"], "test/virus/k18.csv"),
    (["-f", "0.1", "test/virus/seqs.list"], "test/virus/k18.frac.csv"),
    (["-k", "24", "test/virus/seqs.list"], "test/virus/k24.csv"),
    (["-t", "1", "test/virus/seqs.list"], "test/virus/k18.csv"),
    (["-multisample-fasta", "-k", "21", "test/synth/synth.list"], "test/synth/a2a"),
])
def test_build_then_oracle_all2all_reproduces_reference_csv(cli, oracle, ref_fixtures, tmp_path, args, golden):
    cli(ref_fixtures, "build", "-host-build", *args, tmp_path / "x.db")
    assert _oracle_all2all_csv(oracle, tmp_path / "x.db", tmp_path / "x.csv") == ou.read_bytes(ref_fixtures / golden)


def test_build_extend(cli, oracle, ref_fixtures, tmp_path):
    """build part1, then -extend with part2 (k given on the command line is ignored, main.yml:132-135)."""
    db = tmp_path / "parts.db"
    cli(ref_fixtures, "build", "-host-build", "test/virus/seqs.part1.list", db)
    cli(ref_fixtures, "build", "-host-build", "-extend", "-k", "25", "test/virus/seqs.part2.list", db)
    assert _oracle_all2all_csv(oracle, db, tmp_path / "x.csv") == ou.read_bytes(ref_fixtures / "test/virus/k18.csv")
    assert _oracle_all2all_csv(oracle, db, tmp_path / "s.csv", sparse=True) == ou.read_bytes(ref_fixtures / "test/virus/k18.sparse.csv")


@pytest.mark.parametrize("alphabet", ["aa", "aa11_diamond", "aa12_mmseqs", "aa6_dayhoff"])
def test_build_amino_alphabets(cli, oracle, ref_fixtures, tmp_path, alphabet):
    cli(ref_fixtures, "build", "-host-build", "-k", "8", "-multisample-fasta", "-alphabet", alphabet, "test/protein/aa_100x1000.fasta", tmp_path / "a.db")
    assert _oracle_all2all_csv(oracle, tmp_path / "a.db", tmp_path / "a.csv") == ou.read_bytes(ref_fixtures / f"test/protein/{alphabet}.a2a")


def test_built_kmer_tables_serve_queries(cli, oracle, ref_fixtures, golden_dbs, tmp_path):
    """The k-mer tables our build writes must answer the reference's new2all vectors; the same
    oracle on the database the REFERENCE built (tests/golden) pins the oracle itself."""
    ours = tmp_path / "p1.db"
    cli(ref_fixtures, "build", "-host-build", "test/virus/seqs.part1.list", ours)
    theirs = ou.ROOT / "tests" / "golden" / "virus.k18.part1.db"
    for db in (ours, theirs):
        got = _oracle_new2all_csv(oracle, db, "test/virus/seqs.part2.list", tmp_path / "n.csv", cwd=ref_fixtures)
        assert got == ou.read_bytes(ref_fixtures / "test/virus/k18.n2a.csv")
        got = _oracle_new2all_csv(oracle, db, "test/virus/seqs.part2.list", tmp_path / "n.csv", sparse=True, cwd=ref_fixtures)
        assert got == ou.read_bytes(ref_fixtures / "test/virus/k18.n2a.sparse.csv")
    got = _oracle_new2all_csv(oracle, golden_dbs["virus.k18"][0], "test/virus/seqs.list", tmp_path / "i.csv", cwd=ref_fixtures)
    assert got == ou.read_bytes(ref_fixtures / "test/virus/k18.n2a.itself.csv")
    got = _oracle_new2all_csv(oracle, golden_dbs["synth.k21"][0], "test/synth/synth.list", tmp_path / "s.csv", multi=True, cwd=ref_fixtures)
    assert got == ou.read_bytes(ref_fixtures / "test/synth/n2a")
    got = _oracle_new2all_csv(oracle, golden_dbs["synth.k21"][0], "test/synth/synth.list", tmp_path / "s.csv", multi=True, sparse=True,
                              cwd=ref_fixtures)
    assert got == ou.read_bytes(ref_fixtures / "test/synth/n2a-sparse")


def test_reference_binary_accepts_our_database(cli, ref_bin, ref_fixtures, tmp_path):
    if ref_bin is None:
        pytest.skip("reference binary not built (oracle/_ref)")
    import subprocess
    cli(ref_fixtures, "build", "-host-build", "test/virus/seqs.part1.list", tmp_path / "p1.db")
    subprocess.run([str(ref_bin), "new2all", str(tmp_path / "p1.db"), "test/virus/seqs.part2.list", str(tmp_path / "n.csv")],
                   cwd=str(ref_fixtures), check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    assert ou.read_bytes(tmp_path / "n.csv") == ou.read_bytes(ref_fixtures / "test/virus/k18.n2a.csv")
    subprocess.run([str(ref_bin), "build", "-extend", "test/virus/seqs.part2.list", str(tmp_path / "p1.db")],
                   cwd=str(ref_fixtures), check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    subprocess.run([str(ref_bin), "all2all", str(tmp_path / "p1.db"), str(tmp_path / "a.csv")],
                   cwd=str(ref_fixtures), check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    assert ou.read_bytes(tmp_path / "a.csv") == ou.read_bytes(ref_fixtures / "test/virus/k18.csv")


DISTANCE_CASES = [  # (arguments, input table, golden output) — main.yml:99-110, self-hosted.yml:146-260
    (["jaccard"], "test/virus/k18.csv", "test/virus/k18.csv.jaccard"),
    (["min"], "test/virus/k18.csv", "test/virus/k18.csv.min"),
    (["max"], "test/virus/k18.csv", "test/virus/k18.csv.max"),
    (["cosine"], "test/virus/k18.csv", "test/virus/k18.csv.cosine"),
    (["mash"], "test/virus/k18.csv", "test/virus/k18.csv.mash"),
    (["mash"], "test/synth/a2a", "test/synth/a2a.mash"),
    (["ani"], "test/synth/a2a", "test/synth/a2a.ani"),
    (["-sparse", "ani"], "test/synth/a2a", "test/synth/a2a-sparse.ani"),
    (["-sparse", "-max", "1.0", "-min", "-1.0", "mash"], "test/synth/a2a", "test/synth/a2a-sparse.mash"),
    (["mash"], "test/synth/a2a-sparse", "test/synth/a2a-sparse.mash"),
    (["ani"], "test/synth/a2a-sparse", "test/synth/a2a-sparse.ani"),
    (["-sparse", "mash", "-min", "0.03", "-max", "mash:1.0"], "test/synth/a2a-sparse", "test/synth/a2a.mash.above-below"),
    (["-sparse", "-min", "0.03", "-max", "mash:1.0", "-min", "num-kmers:36", "mash"], "test/synth/a2a", "test/synth/a2a.mash-sparse-min2max"),
    (["mash"], "test/synth/n2a", "test/synth/n2a.mash"),
    (["ani"], "test/synth/n2a", "test/synth/n2a.ani"),
    (["-sparse", "ani"], "test/synth/n2a", "test/synth/n2a-sparse.ani"),
    (["-sparse", "-max", "1.0", "-min", "-1.0", "mash"], "test/synth/n2a", "test/synth/n2a-sparse.mash"),
    (["mash"], "test/synth/n2a-sparse", "test/synth/n2a-sparse.mash"),
    (["ani"], "test/synth/n2a-sparse", "test/synth/n2a-sparse.ani"),
]


@pytest.mark.parametrize("args,table,golden", DISTANCE_CASES)
def test_distance_bytes(cli, ref_fixtures, tmp_path, args, table, golden):
    cli(ref_fixtures, "distance", *args, table, tmp_path / "d.csv")
    assert ou.read_bytes(tmp_path / "d.csv") == ou.read_bytes(ref_fixtures / golden)


def test_cli_errors(cli, ref_fixtures, tmp_path):
    r = cli(ref_fixtures, "distance", "nonsense", "test/synth/a2a", tmp_path / "d", check=False)
    assert r.returncode != 0 and "ERROR" in r.stderr
    r = cli(ref_fixtures, "distance", "-min", "foo:1", "mash", "test/synth/a2a", tmp_path / "d", check=False)
    assert r.returncode != 0 and "unknown metric" in r.stderr
    r = cli(ref_fixtures, "build", "-k", "40", "test/virus/seqs.list", tmp_path / "x.db", check=False)
    assert r.returncode != 0 and "cannot exceed" in r.stderr
    r = cli(ref_fixtures, "build", "-host-build", "missing.list", tmp_path / "x.db", check=False)
    assert r.returncode != 0 and "Unable to open input file" in r.stderr
    r = cli(ref_fixtures, "build", "only-one-file", check=False)
    assert r.returncode != 0
    # -sample-rows: the options are parsed before any device is opened (src/params.cpp:533-556)
    for words, msg in ((["-sample-rows", "3"], "random selection"), (["-sample-rows", "nosuch:3"], "unknown measure"),
                       (["-sample-rows", "jaccard:x"], "unable to parse numerical value")):
        for mode in ("all2all-sp", "all2all-parts"):
            r = cli(ref_fixtures, mode, *words, "in", tmp_path / "o.csv", check=False)
            assert r.returncode != 0 and msg in r.stderr, (mode, words, r.stderr)
    assert cli(ref_fixtures, "-version").stdout.startswith("kmer-db-b200")


@pytest.mark.parametrize("seed", range(4))
def test_builder_matches_set_intersections(libs, oracle, seed):
    """Semantic ground truth: M[s][t] = |K_s ∩ K_t| for k-mer sets K_s — checked
    on databases the host builder makes from random overlapping sets, through the oracle's
    all2all and its one2all."""
    rng = np.random.default_rng(seed)
    N = int(rng.integers(2, 40))
    universe = np.unique(rng.integers(0, 1 << 40, size=3000, dtype=np.uint64))
    sets = []
    for s in range(N):
        base = sets[int(rng.integers(0, s))] if s and rng.random() < 0.7 else universe[rng.random(universe.size) < 0.2]
        keep = base[rng.random(base.size) < 0.9]
        extra = universe[rng.random(universe.size) < 0.02]
        sets.append(np.unique(np.concatenate([keep, extra])))
    if seed == 1:
        sets[3 % N] = np.zeros(0, np.uint64)  # an empty sample is registered but adds nothing
    t = libs.Trie.build([(f"s{i}", k) for i, k in enumerate(sets)], k=20)
    t.validate()
    assert t.sample_kmer_counts().tolist() == [len(k) for k in sets]
    tri, _ = ou.oracle_all2all(oracle, N, t.arrays())
    want = np.zeros(ou.tri_cells(N), np.uint32)
    for a in range(1, N):
        for b in range(a):
            want[a * (a - 1) // 2 + b] = np.intersect1d(sets[a], sets[b], assume_unique=True).size
    assert np.array_equal(tri, want)


def test_sequence_stream_carries_what_sample_stream_extracts(libs, ref_fixtures, tmp_path):
    """The device builder and new2all take SequenceStream's symbols (records separated by NUL); extracting k-mers
    from those symbols on the host must give exactly SampleStream's k-mers, for list files, multi-sample FASTA and
    gzip input alike (the device-side extraction is compared with the host's in tests/test_gpu_build.py)."""
    import subprocess
    src = tmp_path / "seqcheck.cpp"
    src.write_text(r'''
#include <algorithm>
#include <cstdio>
#include "ingest.h"
using namespace kdbx;
int main(int argc, char** argv) {
    const std::string list = argv[1]; const bool multi = std::string(argv[2]) == "1"; const uint32_t k = (uint32_t)std::atoi(argv[3]);
    const Alphabet al = Alphabet::make(kNt); const MinHash f(1.0, 0.0, k);
    SampleStream a(list, al, f, k, multi, 2);
    SequenceStream b(list, multi, 2);
    SampleKmers sa; SampleSeq sb; size_t n = 0;
    for (;;) {
        const bool ma = a.next(sa), mb = b.next(sb);
        if (ma != mb) { std::printf("streams differ in length\n"); return 1; }
        if (!ma) break;
        if (sa.name != sb.name) { std::printf("names differ: %s %s\n", sa.name.c_str(), sb.name.c_str()); return 1; }
        std::vector<uint64_t> km;
        extract_kmers(sb.symbols.data(), sb.symbols.size(), k, al, f, km);   // NUL is outside the alphabet
        std::sort(km.begin(), km.end()); km.erase(std::unique(km.begin(), km.end()), km.end());
        if (km != sa.kmers) { std::printf("k-mers differ for %s\n", sa.name.c_str()); return 1; }
        ++n;
    }
    std::printf("%zu\n", n);
    return 0;
}
''')
    exe = tmp_path / "seqcheck"
    host = ou.ROOT / "kmer-db_b200" / "host"
    lib = ou.ROOT / "kmer-db_b200" / "lib"
    subprocess.run(["g++", "-O1", "-std=c++17", "-pthread", "-I", str(host), str(src), "-o", str(exe), "-L", str(lib), "-lkdbx_host", "-lkdbx",
                    f"-Wl,-rpath,{lib}"], check=True)
    for lst, multi, k, want in (("test/virus/seqs.list", "0", 18, None), ("test/virus/multi.list", "1", 18, None),
                                ("test/synth/synth.list", "1", 21, None), ("test/virus/seqs.part2.list", "0", 24, None)):
        r = subprocess.run([str(exe), lst, multi, str(k)], cwd=str(ref_fixtures), capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        assert int(r.stdout.strip()) > 0


def _split_virus_lists(ref_fixtures, tmp_path, cuts):
    """List files holding consecutive slices of test/virus/seqs.list (cuts = slice boundaries)."""
    lines = (ref_fixtures / "test/virus/seqs.list").read_text().split()
    out = []
    for i, (b, e) in enumerate(zip([0] + cuts, cuts + [len(lines)])):
        f = tmp_path / f"slice{i}.list"
        f.write_text("\n".join(lines[b:e]) + "\n")
        out.append(f)
    return out


@pytest.mark.parametrize("cuts", [[100], [60, 120], [1, 164]])
def test_oracle_all2all_parts_reproduces_reference_csv(cli, oracle, ref_fixtures, tmp_path, cuts):
    """all2all-parts over the virus genomes split into partial databases gives the all2all-sp table of the whole
    collection — the reference's own CI check (.github/workflows/self-hosted.yml:357-363: parts1 + parts2 -> k18.sparse.csv).
    Pins the oracle's restatement of db2db_sp and of the grid driver; the parts are built by our host builder, and in
    the CI's own split part 1 is also taken as the REFERENCE built it (tests/golden/virus.k18.part1.db)."""
    oracle.oracle_all2all_parts_file.argtypes = [C.c_char_p, C.c_char_p]
    oracle.oracle_all2all_parts_file.restype = C.c_int64
    dbs = []
    for i, lst in enumerate(_split_virus_lists(ref_fixtures, tmp_path, cuts)):
        dbs.append(tmp_path / f"part{i}.db")
        cli(ref_fixtures, "build", "-host-build", lst, dbs[-1])
    variants = [dbs]
    if cuts == [100]:
        variants.append([ou.ROOT / "tests" / "golden" / "virus.k18.part1.db", dbs[1]])
    for v in variants:
        (tmp_path / "db.list").write_text("\n".join(map(str, v)) + "\n")
        pairs = oracle.oracle_all2all_parts_file(str(tmp_path / "db.list").encode(), str(tmp_path / "parts.csv").encode())
        assert pairs == 13530   # "No. saved pairs" of the reference run
        assert ou.read_bytes(tmp_path / "parts.csv") == ou.read_bytes(ref_fixtures / "test/virus/k18.sparse.csv")


def test_one2all_table_reproduces_reference_csv(cli, libs, oracle, ref_fixtures, tmp_path):
    """The reference's CI step `build -k 25 -f 0.1 seqs.part1.list` + `one2all k25.db data/MT159713` == test/virus/MT159713.csv
    (.github/workflows/main.yml:156-160).  CPU: our host builder writes the database, the oracle's new2all supplies the
    similarity vector of that one query, the product's one2all emitter (host/csv_out.cpp::write_one2all_csv) formats it."""
    db = tmp_path / "k25.db"
    cli(ref_fixtures, "build", "-host-build", "-k", "25", "-f", "0.1", "test/virus/seqs.part1.list", db)
    (tmp_path / "q.list").write_text("./test/virus/data/MT159713\n")
    n2a = _oracle_new2all_csv(oracle, db, tmp_path / "q.list", tmp_path / "n2a.csv", cwd=ref_fixtures).decode().splitlines()
    assert len(n2a) == 3
    row = n2a[2].split(",")
    assert row[0] == "MT159713" and row[-1] == ""
    t = libs.Trie.read_db(db)
    t.write_one2all_csv("./test/virus/data/MT159713", int(row[1]), np.array(row[2:-1], dtype=np.uint32), tmp_path / "o.csv")
    assert ou.read_bytes(tmp_path / "o.csv") == ou.read_bytes(ref_fixtures / "test/virus/MT159713.csv")


def _virus_copy(ref_fixtures, tmp_path):
    """a private copy of the virus inputs (the minhash mode writes next to the samples)"""
    work = tmp_path / "w"
    shutil.copytree(ref_fixtures / "test" / "virus", work / "test" / "virus")
    return work


def test_minhash_mode_and_build_from_minhash(cli, oracle, ref_bin, ref_fixtures, tmp_path):
    """The reference's CI step `minhash -f 0.1 seqs.list; build -from-minhash seqs.list db; all2all` == test/virus/k18.frac.csv
    (.github/workflows/main.yml:143-148).  The .minhash files must be the reference's byte for byte (digests of what the
    reference binary wrote: tests/golden/virus.minhash.f01.sha256); CPU: the host builder reads them back, the oracle runs
    all2all."""
    import hashlib
    work = _virus_copy(ref_fixtures, tmp_path)
    cli(work, "minhash", "-f", "0.1", "test/virus/seqs.list")
    want = dict(reversed(ln.split()) for ln in (ou.ROOT / "tests" / "golden" / "virus.minhash.f01.sha256").read_text().splitlines())
    got = {f.name: hashlib.sha256(f.read_bytes()).hexdigest() for f in (work / "test/virus/data").glob("*.minhash")}
    assert got == want and len(got) == 165
    cli(work, "build", "-host-build", "-from-minhash", "test/virus/seqs.list", tmp_path / "mh.db")
    assert _oracle_all2all_csv(oracle, tmp_path / "mh.db", tmp_path / "mh.csv") == ou.read_bytes(ref_fixtures / "test/virus/k18.frac.csv")
    if ref_bin is not None:   # the unmodified reference builds the same database from OUR files
        import subprocess
        subprocess.run([str(ref_bin), "build", "-from-minhash", "test/virus/seqs.list", str(tmp_path / "ref.db")], cwd=work, check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        assert _oracle_all2all_csv(oracle, tmp_path / "ref.db", tmp_path / "ref.csv") == ou.read_bytes(ref_fixtures / "test/virus/k18.frac.csv")
    # default fraction of the mode is 0.01 (src/params.cpp:130-133); k and fraction travel in the files
    cli(work, "minhash", "-k", "20", "test/virus/seqs.part2.list")
    raw = (work / "test/virus/data/MT159713.minhash").read_bytes()
    n = int.from_bytes(raw[4:12], "little")
    assert raw[:4] == bytes.fromhex("98badcfe") and len(raw) == 4 + 8 + 8 * n + 4 + 8
    assert int.from_bytes(raw[12 + 8 * n:16 + 8 * n], "little") == 20 and np.frombuffer(raw[16 + 8 * n:], np.float64)[0] == 0.01
    # a database cannot take samples of another k-mer length or fraction; a missing or damaged file is an error
    r = cli(work, "build", "-host-build", "-from-minhash", "test/virus/seqs.list", tmp_path / "bad.db", check=False)
    assert r.returncode != 0
    (work / "test/virus/data/NC_045512.minhash").write_bytes(b"\x98\xba\xdc\xfe" + (5).to_bytes(8, "little") + b"\0" * 8)
    r = cli(work, "build", "-host-build", "-from-minhash", "test/virus/seqs.part1.list", tmp_path / "bad.db", check=False)
    assert "failed:./test/virus/data/NC_045512" in r.stderr
    r = cli(work, "minhash", "-from-kmers", "test/virus/seqs.list", check=False)
    assert r.returncode != 0 and "not supported" in r.stderr
