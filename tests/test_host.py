"""Host logic and the C-ABI surface; no GPU needed."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

import oracle_util as ou

ROOT = Path(__file__).resolve().parent.parent


def _declared(header):
    text = (ROOT / "include" / header).read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kdbxh?_[a-z0-9_]+)\s*\(", text)))


def test_libraries_export_every_declared_symbol(libs):
    k, h = libs.load()
    for name in _declared("kdbx.h"):
        assert hasattr(k, name), name
    for name in _declared("kdbx_host.h"):
        assert hasattr(h, name), name
    assert set(_declared("kdbx.h")) == set(libs.KDBX_SYMBOLS)
    assert set(_declared("kdbx_host.h")) == set(libs.KDBXH_SYMBOLS)
    assert k.kdbx_abi_version() == 4


def test_struct_sizes_match_header(libs, tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "kdbx_host.h"\nint main(){printf("%zu %zu %zu %zu %zu",sizeof(kdbx_config),'
                   'sizeof(kdbx_trie_view),sizeof(kdbx_stats),sizeof(kdbxh_synth_params),sizeof(kdbxh_totals));'
                   'printf(" %zu %zu %zu",sizeof(kdbx_filter),sizeof(kdbx_csr),sizeof(kdbx_tables_view));'
                   'printf(" %zu %zu %zu",sizeof(kdbx_build_params),sizeof(kdbx_build_result),sizeof(kdbx_build_arrays));return 0;}')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", str(ROOT / "include"), str(src), "-o", str(exe)], check=True)
    sizes = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [C.sizeof(libs.Config), C.sizeof(libs.TrieView), C.sizeof(libs.Stats), C.sizeof(libs.SynthParams),
                     C.sizeof(libs.Totals), C.sizeof(libs.Filter), C.sizeof(libs.Csr), C.sizeof(libs.TablesView),
                     C.sizeof(libs.BuildParams), C.sizeof(libs.BuildResult), C.sizeof(libs.BuildArrays)]


def test_no_gpu_means_loud_failure_not_fallback(libs):
    k, _ = libs.load()
    if k.kdbx_device_count() > 0:
        pytest.skip("a B200 is present")
    with pytest.raises(libs.KdbxError, match="no CPU fallback|sm_100a"):
        libs.Context()


def test_product_never_touches_the_oracle():
    for p in (ROOT / "kmer-db_b200").rglob("*"):
        if p.suffix in {".cpp", ".h", ".cu", ".py"} or p.name == "Makefile":
            assert "oracle" not in p.read_text(), p


@pytest.mark.parametrize("name", ["virus.k18", "virus.k18.f01", "virus.k24", "synth.k21"])
def test_reader_agrees_with_oracle_reader_and_csv_writer_is_byte_exact(libs, oracle, golden_dbs, tmp_path, name):
    db, dense, sparse = golden_dbs[name]
    t = libs.Trie.read_db(db)
    t.validate()
    a = t.arrays()
    tri, U = ou.oracle_all2all(oracle, t.num_samples, a)   # the checker computes the matrix ...
    assert U == t.totals().updates
    out = tmp_path / "h.csv"
    t.write_all2all_csv(tri, out)                           # ... the product formats it
    assert ou.read_bytes(out) == ou.read_bytes(dense)
    if sparse is not None:
        t.write_all2all_csv(tri, out, sparse=True)
        assert ou.read_bytes(out) == ou.read_bytes(sparse)


def test_db_writer_round_trip_and_reference_accepts_it(libs, golden_dbs, ref_bin, tmp_path):
    db, dense, _ = golden_dbs["virus.k18"]
    t = libs.Trie.read_db(db)
    out = tmp_path / "copy.db"
    t.write_db(out)
    t2 = libs.Trie.read_db(out)
    a, b = t.arrays(), t2.arrays()
    for key in a:
        assert np.array_equal(a[key], b[key]), key
    assert t.sample_names() == t2.sample_names()
    if ref_bin is not None:  # the unmodified reference computes the golden CSV from OUR file
        subprocess.run([str(ref_bin), "all2all", "-t", "2", str(out), str(tmp_path / "r.csv")], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        assert ou.read_bytes(tmp_path / "r.csv") == ou.read_bytes(dense)


def test_missing_db_is_an_error(libs, tmp_path):
    with pytest.raises(libs.KdbxError, match="Cannot open k-mer database"):
        libs.Trie.read_db(tmp_path / "nope.db")
    (tmp_path / "short.db").write_bytes(b"\x01\x00\x00")
    with pytest.raises(libs.KdbxError):
        libs.Trie.read_db(tmp_path / "short.db")


@pytest.mark.parametrize("interleaved", [False, True])
def test_generator_makes_valid_tries_with_exact_kmer_accounting(libs, oracle, interleaved):
    L = 20000
    t = libs.Trie.synth(num_samples=40, num_clusters=3, genome_kmers=L, seed=3, interleaved=interleaved)
    t.validate()
    a = t.arrays()
    N = t.num_samples
    # every sample's k-mers are partitioned over the patterns whose full list contains it
    per_sample = np.zeros(N, np.int64)
    P = len(a["n"])
    for p in range(1, P):
        q = p
        while q >= 0:
            ids = np.zeros(int(a["l"][q]), np.uint32)
            oracle.oracle_decode_local(a["payload"][int(a["payload_off"][q]):].ctypes.data, int(a["l"][q]), int(a["last"][q]),
                                       ids.ctypes.data)
            per_sample[ids] += int(a["num_kmers"][p])
            q = int(a["parent_id"][q])
    assert (per_sample == L).all()
    # shared k-mers never exceed a genome, members of different clusters share nothing
    tri, _ = ou.oracle_all2all(oracle, N, a)
    assert tri.max() <= L
    cl = (lambda s: s % 3) if interleaved else (lambda s: s * 3 // N)
    for s in range(N):
        for c in range(s):
            if cl(s) != cl(c):
                assert tri[s * (s - 1) // 2 + c] == 0
    # determinism
    t2 = libs.Trie.synth(num_samples=40, num_clusters=3, genome_kmers=L, seed=3, interleaved=interleaved, threads=1)
    for key, v in t2.arrays().items():
        assert np.array_equal(v, a[key]), key


def test_shard_rows_by_work(libs):
    w = np.array([0, 1, 2, 3, 4, 5, 6, 7, 8, 9] * 10, dtype=np.uint64)
    for world in (1, 2, 4, 8):
        b = libs.shard_rows_by_work(w, world)
        assert b[0] == 0 and b[-1] == len(w) and all(x <= y for x, y in zip(b, b[1:]))
        loads = [int(w[b[i]:b[i + 1]].sum()) for i in range(world)]
        assert max(loads) - min(loads) <= 2 * int(w.max())


def test_prefix_cuts_a_cluster_exactly(libs, oracle):
    t = libs.Trie.synth(num_samples=60, num_clusters=3, genome_kmers=15000, seed=4)
    sub = t.prefix(20)
    sub.validate()
    full, _ = ou.oracle_all2all(oracle, 60, t.arrays())
    part, _ = ou.oracle_all2all(oracle, 20, sub.arrays())
    assert np.array_equal(part, full[:ou.tri_cells(20)])
    with pytest.raises(libs.KdbxError, match="prefix"):
        t.prefix(30)  # inside a cluster
    ti = libs.Trie.synth(num_samples=60, num_clusters=3, genome_kmers=15000, seed=4, interleaved=True)
    with pytest.raises(libs.KdbxError, match="prefix"):
        ti.prefix(20)


@pytest.mark.parametrize("source", ["synth", "synth-interleaved", "virus.k18", "synth.k21"])
@pytest.mark.parametrize("parts", [1, 2, 3, 7])
def test_partition_parts_sum_to_the_whole_matrix(libs, oracle, golden_dbs, source, parts):
    """Sharding for multi-GPU runs: every part is a valid trie, the parts own disjoint patterns that
    cover the trie, and their matrices add up (uint32) to the matrix of the whole database."""
    if source.startswith("synth") and "." not in source:
        t = libs.Trie.synth(num_samples=48, num_clusters=3, genome_kmers=12000, seed=9, interleaved=source.endswith("interleaved"))
    else:
        t = libs.Trie.read_db(golden_dbs[source][0])
    N = t.num_samples
    a = t.arrays()
    full, U = ou.oracle_all2all(oracle, N, a)
    acc = np.zeros_like(full)
    owned_u, owned_patterns, kmers = 0, 0, 0
    costs = []
    for r in range(parts):
        sub, u = t.partition(parts, r)
        sub.validate()
        assert sub.num_samples == N and sub.sample_names() == t.sample_names()
        b = sub.arrays()
        tri, _ = ou.oracle_all2all(oracle, N, b)
        acc += tri  # uint32, wraps like the device all-reduce
        owned_u += u
        owned_patterns += int((b["num_kmers"] != 0).sum())
        kmers += int(b["num_kmers"].sum())
        costs.append(int((b["l"].astype(np.int64) * (2 * b["n"].astype(np.int64) - b["l"] - 1) // 2 + 34 * b["n"].astype(np.int64) + 200).sum()))
    assert np.array_equal(acc, full)
    assert owned_u == U
    assert kmers == int(a["num_kmers"].sum())
    assert owned_patterns == int((a["num_kmers"] != 0).sum())
    if parts > 1 and len(a["n"]) > 2000:  # balanced on the cost model (one pattern of slack + the replicated chain)
        assert max(costs) <= 1.25 * (sum(costs) / parts)


def test_partition_rejects_bad_arguments(libs):
    t = libs.Trie.synth(num_samples=8, num_clusters=2, genome_kmers=500, seed=1)
    with pytest.raises(libs.KdbxError, match="partition"):
        t.partition(0, 0)
    with pytest.raises(libs.KdbxError, match="partition"):
        t.partition(2, 2)


def test_relabel_moves_the_matrix_block(libs, oracle):
    t = libs.Trie.synth(num_samples=20, num_clusters=2, genome_kmers=3000, seed=5)
    base, U = ou.oracle_all2all(oracle, 20, t.arrays())
    t.relabel(30, 64)
    t.validate()
    assert t.num_samples == 64 and t.totals().updates == U
    moved, U2 = ou.oracle_all2all(oracle, 64, t.arrays())
    assert U2 == U
    want = np.zeros(ou.tri_cells(64), np.uint32)
    for s in range(1, 20):
        src = base[ou.tri_cells(s):ou.tri_cells(s) + s]
        o = ou.tri_cells(s + 30) + 30
        want[o:o + s] = src
    assert np.array_equal(moved, want)
    with pytest.raises(libs.KdbxError, match="relabel"):
        t.relabel(1, 64)


def test_partition_into_more_parts_than_patterns(libs, oracle):
    """Empty parts are valid one-pattern tries (the sentinel) whose matrix is zero."""
    t = libs.Trie.synth(num_samples=6, num_clusters=1, genome_kmers=40, seed=2, mutation_rate=0.05)
    N = t.num_samples
    full, U = ou.oracle_all2all(oracle, N, t.arrays())
    P = t.num_patterns
    parts = P + 5
    acc = np.zeros_like(full)
    owned, empty = 0, 0
    for r in range(parts):
        sub, u = t.partition(parts, r)
        sub.validate()
        tri, _ = ou.oracle_all2all(oracle, N, sub.arrays())
        acc += tri
        owned += u
        empty += int(sub.num_patterns == 1)
    assert np.array_equal(acc, full) and owned == U and empty >= 5


def test_column_window_arithmetic():
    """The diagonal-relative column windows of the scatter kernel (win_hi / win_col0 / make_plan in
    csrc/kdbx.cu), restated: for every row block the windows are disjoint, cover all columns below the block's
    rows, hold at most tile_cols columns each and number at most T; window 0 ends at the block's 32-aligned end."""
    rng = np.random.default_rng(1)
    for _ in range(300):
        N = int(rng.integers(2, 6000))
        tr = int(2 ** rng.integers(0, 6))
        tc = 32 * int(rng.integers(1, 80))
        RB = (N + tr - 1) // tr
        hi_max = (RB * tr + 31) & ~31
        T = max(1, (hi_max + tc - 1) // tc)
        for rb in {0, 1, RB // 2, RB - 1} & set(range(RB)):
            hi = ((rb + 1) * tr + 31) & ~31
            assert (rb + 1) * tr <= hi <= hi_max
            covered = 0
            prev_lo = hi
            for t in range(T):
                col0 = max(0, hi - (t + 1) * tc)
                upper = hi - t * tc
                if upper <= 0:
                    break
                assert upper == prev_lo and 0 < upper - col0 <= tc
                covered += upper - col0
                prev_lo = col0
                if col0 == 0:
                    break
            assert prev_lo == 0 and covered == hi     # [0, hi) is tiled exactly
            # the window of a column: t = (hi - 1 - c) // tile_cols, as emit_run computes it
            for c in rng.integers(0, hi, size=5):
                t = (hi - 1 - int(c)) // tc
                assert t < T and max(0, hi - (t + 1) * tc) <= c < hi - t * tc


def test_reader_builds_32_bit_header_mirrors(libs, golden_dbs):
    """kdbx_trie_view::parent_id32 / num_kmers32: the reader and the partitioner keep 32-bit mirrors of the two 64-bit
    header arrays (24 instead of 40 bytes per pattern to upload); they must equal the 64-bit arrays."""
    import ctypes as C
    t = libs.Trie.read_db(golden_dbs["virus.k18"][0])
    v = t.view()
    P = int(v.num_patterns)
    assert v.parent_id32 and v.num_kmers32
    a = t.arrays()
    p32 = np.ctypeslib.as_array(C.cast(v.parent_id32, C.POINTER(C.c_int32)), (P,))
    k32 = np.ctypeslib.as_array(C.cast(v.num_kmers32, C.POINTER(C.c_uint32)), (P,))
    assert np.array_equal(p32.astype(np.int64), a["parent_id"]) and np.array_equal(k32.astype(np.int64), a["num_kmers"])
    for part, _, _ in t.partition_all(3):
        pv = part.view()
        pa = part.arrays()
        q32 = np.ctypeslib.as_array(C.cast(pv.parent_id32, C.POINTER(C.c_int32)), (int(pv.num_patterns),))
        assert np.array_equal(q32.astype(np.int64), pa["parent_id"])


def test_pattern_level_generator_matches_a_fasta_level_build_of_the_same_model(libs, tmp_path):
    """host/synth.cpp simulates `build` on k-mer runs instead of sequences.  The same model at the
    FASTA level — first genome of a cluster random, every later one a copy of a uniformly chosen earlier member with
    i.i.d. substitutions — pushed through the real ingest + builder must give the same database STATISTICS.  The two
    draw different random copy trees, so the comparison is between means over seeds (within 12 %; the number of distinct
    k-mers, which does not depend on the tree, within 3 %)."""
    def fasta_db(N, C, L, mu, seed, k=18):
        rng = np.random.default_rng(seed)
        acgt = np.frombuffer(b"ACGT", np.uint8)
        members = [[] for _ in range(C)]
        path = tmp_path / f"s{seed}.fa"
        with open(path, "w") as f:
            for g in range(N):
                c = (g * C) // N
                if not members[c]:
                    code = rng.integers(0, 4, size=L + k - 1)
                else:
                    code = members[c][int(rng.integers(0, len(members[c])))].copy()
                    pos = np.flatnonzero(rng.random(L + k - 1) < mu)
                    code[pos] = (code[pos] + rng.integers(1, 4, size=pos.size)) % 4   # a different base
                members[c].append(code)
                f.write(f">g{g:04d}\n{acgt[code].tobytes().decode()}\n")
        return libs.Trie.build(libs.load_samples(path, k=k, multisample=True, threads=4), k=k, threads=4)
    N, C, L, seeds = 64, 2, 20000, range(1, 17)   # per-seed spread: 3 % (patterns) .. 7 % (updates); 16 seeds -> ~2 % on the means
    fields = ("num_patterns", "sum_n", "sum_l", "updates", "kmers_count")
    fa, sy = dict.fromkeys(fields, 0), dict.fromkeys(fields, 0)
    for seed in seeds:
        a = fasta_db(N, C, L, 0.005, 100 + seed).totals()
        b = libs.Trie.synth(num_samples=N, num_clusters=C, genome_kmers=L, mutation_rate=0.005, seed=seed).totals()
        for f in fields:
            fa[f] += getattr(a, f); sy[f] += getattr(b, f)
    for f in fields:
        assert abs(sy[f] - fa[f]) <= 0.12 * fa[f], (f, fa[f] / len(seeds), sy[f] / len(seeds))
    assert abs(sy["kmers_count"] - fa["kmers_count"]) <= 0.03 * fa["kmers_count"]


# ---- the decoder's per-pattern code (kmer-db_b200/csrc/gamma_tokens.cuh), compiled for the host ----------------------------
@pytest.fixture(scope="module")
def gamma_tokens(tmp_path_factory):
    so = tmp_path_factory.mktemp("native") / "libgamma_tokens_host.so"
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", str(so), str(ROOT / "tests" / "native" / "gamma_tokens_host.cpp")], check=True)
    lib = C.CDLL(str(so))
    lib.kdbx_test_decode_list.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    return lib


def _encode_gamma(deltas):
    """(words with two guard words, number of bits) of the Elias-gamma stream of `deltas` (MSB first; 1 -> '0')."""
    bits = []
    for d in deltas:
        b = int(d).bit_length()
        bits += [1] * (b - 1) + [0] + [(int(d) >> k) & 1 for k in range(b - 2, -1, -1)]
    nb = len(bits)
    words = max(2, ((nb + 127) // 128) * 2) + 2
    w = np.zeros(words, np.uint64)
    for i, bit in enumerate(bits):
        if bit:
            w[i >> 6] |= np.uint64(1) << np.uint64(63 - (i & 63))
    return w, nb


def _decode(lib, w, nb, l, last, sh=None):
    out = np.zeros(max(1, l), np.uint32)
    closes = np.zeros(3 * max(1, l), np.uint32)
    n, runs = C.c_uint32(0), C.c_uint32(0)
    rc = lib.kdbx_test_decode_list(w.ctypes.data, nb, l, last, 0 if sh is None else 1, sh or 0, out.ctypes.data, closes.ctypes.data, C.byref(n), C.byref(runs))
    return rc, out[:l], closes[:3 * n.value].reshape(-1, 3), runs.value


def test_gamma_tokens_against_oracle_on_random_lists(gamma_tokens, oracle):
    """Token parser + both second passes == the oracle's decoder (oracle_decode_local) on random ascending lists with long runs
    of consecutive ids, lone ids, huge gaps and runs that straddle 64-bit words; the row-block stretches equal a per-id scan."""
    rng = np.random.default_rng(5)
    for trial in range(400):
        l = int(rng.integers(2, 400))
        style = trial % 4
        if style == 0:
            deltas = np.ones(l - 1, np.int64)
        elif style == 1:
            deltas = rng.integers(1, 3, size=l - 1)
        elif style == 2:
            deltas = np.where(rng.random(l - 1) < 0.85, 1, rng.integers(2, 5000, size=l - 1))
        else:
            deltas = rng.integers(1, 2**20, size=l - 1)
        first = int(rng.integers(0, 1000))
        ids = first + np.concatenate([[0], np.cumsum(deltas)])
        last = int(ids[-1])
        w, nb = _encode_gamma(deltas)
        want = np.zeros(l, np.uint32)
        oracle.oracle_decode_local(w.ctypes.data, l, last, want.ctypes.data)
        assert np.array_equal(want, ids.astype(np.uint32))
        rc, got, _, runs = _decode(gamma_tokens, w, nb, l, last)
        assert rc == 0 and np.array_equal(got, want) and runs == 1 + int((deltas >= 2).sum())
        for sh in (0, 3, 5):
            rc, got, closes, _ = _decode(gamma_tokens, w, nb, l, last, sh)
            assert rc == 0 and np.array_equal(got, want)
            rb = ids >> sh
            starts = np.concatenate([[0], np.nonzero(np.diff(rb))[0] + 1])
            ends = np.concatenate([starts[1:], [l]])
            assert np.array_equal(closes, np.stack([rb[starts], starts, ends - starts], axis=1).astype(np.uint32))


def test_gamma_tokens_rejects_malformed_streams(gamma_tokens):
    deltas = [1, 1, 5, 1, 300, 1, 1, 1, 2]
    w, nb = _encode_gamma(deltas)
    l, last = len(deltas) + 1, 1000
    assert _decode(gamma_tokens, w, nb, l, last)[0] == 0
    assert _decode(gamma_tokens, w, nb - 1, l, last)[0] == 1        # a bit short
    assert _decode(gamma_tokens, w, nb + 1, l, last)[0] == 1        # a bit long
    assert _decode(gamma_tokens, w, nb, l + 1, last)[0] == 1        # one id too many
    assert _decode(gamma_tokens, w, nb, l - 1, last)[0] == 1        # one too few: bits left over
    assert _decode(gamma_tokens, w, nb, l, 100)[0] == 3             # deltas sum to more than the last id
    ones = np.full(4, np.uint64(0xFFFFFFFFFFFFFFFF))
    assert _decode(gamma_tokens, ones, 128, 3, 10)[0] == 1           # a unary prefix longer than any code
    w31, nb31 = _encode_gamma([2**31 + 5])
    assert _decode(gamma_tokens, w31, nb31, 2, 2**32 - 1)[0] == 2    # a delta that no sample id can reach


def test_gamma_tokens_on_the_reference_built_databases(gamma_tokens, oracle, libs, golden_dbs):
    """Every pattern of the databases the reference built (tests/golden), decoded in place from the packed payload (the next
    pattern's words follow immediately): same ids as the oracle."""
    checked = 0
    for name in ("virus.k18", "synth.k21", "virus.k18.f01"):
        t = libs.Trie.read_db(golden_dbs[name][0])   # (the arrays are views into it)
        a = t.arrays()
        pay = np.concatenate([a["payload"], np.zeros(2, np.uint64)])
        for p in range(len(a["n"])):
            l = int(a["l"][p])
            if l < 2:
                continue
            w = pay[int(a["payload_off"][p]):]
            want = np.zeros(l, np.uint32)
            oracle.oracle_decode_local(w.ctypes.data, l, int(a["last"][p]), want.ctypes.data)
            rc, got, closes, _ = _decode(gamma_tokens, w, int(a["bits"][p]), l, int(a["last"][p]), 5)
            assert rc == 0 and np.array_equal(got, want) and int(closes[:, 2].sum()) == l
            checked += 1
    assert checked > 50


def _read_both_ways(libs, path, monkeypatch):
    """(arrays, view facts) of the mapped parallel reader and of the streaming reader (KDBX_DB_READER=stream), or the error"""
    out = []
    for mode in ("mapped", "stream"):
        if mode == "stream":
            monkeypatch.setenv("KDBX_DB_READER", "stream")
        else:
            monkeypatch.delenv("KDBX_DB_READER", raising=False)
        try:
            t = libs.Trie.read_db(path)
        except libs.KdbxError as e:
            out.append(("error", str(e)))
            continue
        v = t.view()
        out.append(({k: np.array(x, copy=True) for k, x in t.arrays().items()}, bool(v.parent_id32), bool(v.payload_off), int(v.payload_words),
                    t.sample_names()))
        t.close()
    monkeypatch.delenv("KDBX_DB_READER", raising=False)
    return out


def test_parallel_mapped_reader_equals_streaming_reader(libs, golden_dbs, tmp_path, monkeypatch):
    """host/db_io.cpp reads the pattern blocks of a mapped file in parallel (count, prefix sum, fill); the sequential
    reader stays for inputs that cannot be mapped.  Same arrays, same 32-bit mirrors / dense-payload facts, same errors."""
    # (a) the reference-built fixtures
    for name in ("virus.k18", "virus.k24", "synth.k21"):
        a, b = _read_both_ways(libs, golden_dbs[name][0], monkeypatch)
        assert a[1:] == b[1:]
        for k in a[0]:
            assert np.array_equal(a[0][k], b[0][k]), (name, k)
    # (b) a file of several pattern blocks (> 64 MB of patterns), written by our writer
    t = libs.Trie.synth(num_samples=200, num_clusters=4, genome_kmers=1_000_000, seed=5, mutation_rate=0.02)
    big = tmp_path / "big.db"
    t.write_db(big)
    assert big.stat().st_size > 2 * (64 << 20)
    a, b = _read_both_ways(libs, big, monkeypatch)
    assert a[1:] == b[1:] and a[1] is True and a[2] is False
    want = t.arrays()
    for k in want:
        assert np.array_equal(a[0][k], want[k]) and np.array_equal(b[0][k], want[k]), k
    # (c) truncated and corrupted pattern sections: both readers refuse
    raw = big.read_bytes()
    cut = tmp_path / "cut.db"
    cut.write_bytes(raw[:len(raw) - 1000])
    a, b = _read_both_ways(libs, cut, monkeypatch)
    assert a[0] == b[0] == "error" and "truncated" in a[1] and "truncated" in b[1]
    small = golden_dbs["virus.k18"][0].read_bytes()
    tv = libs.Trie.read_db(golden_dbs["virus.k18"][0])
    P = tv.num_patterns
    # the one-block fixture ends with its pattern section: [P u64][block bytes u64][P packed patterns]
    block_bytes = 40 * P + 8 * int(tv.view().payload_words)
    at = len(small) - 16 - block_bytes
    assert small[at:at + 8] == np.uint64(P).tobytes() and small[at + 8:at + 16] == np.uint64(block_bytes).tobytes()
    for (off, val), what in (((at, np.uint64(P + 1).tobytes()), "more patterns announced than stored"),
                             ((at, np.uint64(P - 1).tobytes()), "fewer patterns announced than the block holds"),
                             ((at + 8, np.uint64((64 << 20) + 1).tobytes()), "block larger than the reader's limit"),
                             ((at + 16 + 28, np.uint32(0x7FFFFFF0).tobytes()), "num_bits of the first pattern runs over the block")):
        bad = bytearray(small)
        bad[off:off + len(val)] = val
        f = tmp_path / "bad.db"
        f.write_bytes(bytes(bad))
        a, b = _read_both_ways(libs, f, monkeypatch)
        assert a[0] == b[0] == "error", what


def _csr_of_block(tri, r0, r1, c0, c1):
    """rows [r0, r1) x columns [c0, c1) of the packed lower-triangular matrix as (row_ptr, col, val): non-zero cells below
    the diagonal, columns ascending and relative to c0 — what kdbx_all2all_sparse / kdbx_db2db_sparse hand to the host"""
    ptr, cols, vals = [0], [], []
    for r in range(r0, r1):
        row = tri[ou.tri_cells(r):ou.tri_cells(r) + r]
        hi = min(c1, r)
        if hi > c0:
            seg = row[c0:hi]
            nz = np.nonzero(seg)[0]
            cols.append(nz.astype(np.uint32)); vals.append(seg[nz])
        ptr.append(ptr[-1] + (len(cols[-1]) if hi > c0 else 0))
    cat = lambda xs: np.concatenate(xs) if xs else np.zeros(0, np.uint32)
    return np.array(ptr, np.uint64), cat(cols), cat(vals)


@pytest.mark.parametrize("golden,filters,sample_rows", [
    ("virus.k18.sparse.csv", None, None),
    ("virus.k18.sampled.jaccard_3.csv", None, "jaccard:3"),
    ("virus.k18.sampled.num-kmers_5.csv", None, "num-kmers:5"),
    ("virus.k18.sampled.ani_2.csv", None, "ani:2"),
    ("virus.k18.sampled.mash-query_4.csv", None, "mash-query:4"),
    ("virus.k18.sampled.filtered.cosine_2.csv", "-min jaccard:0.99 -max num-kmers:29800", "cosine:2"),
])
def test_sparse_table_with_sample_rows_equals_reference_output(libs, oracle, golden_dbs, tmp_path, golden, filters, sample_rows):
    """-sample-rows <criterion>:<count> (host/csv_out.h::RowSampler): the oracle computes the matrix, the product filters,
    samples and formats it — as one matrix (all2all-sp) and as the grid of cells of all2all-parts with the samples split
    60 / 60 / 44 and 1 / 163 — and the bytes must be what the reference binary wrote (tests/golden/make_golden.sh)."""
    t = libs.Trie.read_db(golden_dbs["virus.k18"][0])
    N = t.num_samples
    tri, _ = ou.oracle_all2all(oracle, N, t.arrays())
    want = ou.read_bytes(ou.ROOT / "tests" / "golden" / golden)
    out = tmp_path / "o.csv"
    saved = t.write_sparse_csv([(0, 0, *_csr_of_block(tri, 0, N, 0, N))], out, filters=filters, sample_rows=sample_rows)
    assert ou.read_bytes(out) == want
    assert saved == sum(len([x for x in ln.split(b",")[2:] if x]) for ln in want.splitlines()[2:])
    if sample_rows is None:
        return
    for cuts in ([60, 120], [1]):
        edges = [0, *cuts, N]
        cells = [(edges[i], edges[j], *_csr_of_block(tri, edges[i], edges[i + 1], edges[j], edges[j + 1]))
                 for i in range(len(edges) - 1) for j in range(i + 1)]
        # (any order of the cells gives the same rows: the selection is a total order on (criterion, sample id))
        t.write_sparse_csv(cells[::-1], out, filters=filters, sample_rows=sample_rows)
        assert ou.read_bytes(out) == want, cuts


def test_sample_rows_argument_errors(libs, golden_dbs, tmp_path):
    t = libs.Trie.read_db(golden_dbs["virus.k18"][0])
    empty = [(0, 0, np.zeros(t.num_samples + 1, np.uint64), np.zeros(0, np.uint32), np.zeros(0, np.uint32))]
    for bad in ("3", "nosuch:3", "jaccard:x", "jaccard:0"):   # no criterion = the reference's random selection: not offered
        with pytest.raises(libs.KdbxError):
            t.write_sparse_csv(empty, tmp_path / "e.csv", sample_rows=bad)
    with pytest.raises(libs.KdbxError):   # a grid of cells is only meaningful with a sampler
        t.write_sparse_csv(empty + empty, tmp_path / "e.csv")


def _tables_of(t):
    """(slot offsets, slots) of a trie read with its k-mer tables, as numpy copies"""
    v = t.tables_view()
    T = int(v.num_tables)
    off = np.ctypeslib.as_array(C.cast(v.slot_off, C.POINTER(C.c_uint64)), shape=(T + 1,)).copy()
    slots = np.ctypeslib.as_array(C.cast(v.slots, C.POINTER(C.c_uint64)), shape=(int(off[-1]),)).copy()
    return off, slots


def test_parallel_mapped_table_reader_equals_streaming_reader(libs, cli, golden_dbs, ref_fixtures, tmp_path, monkeypatch):
    """The raw k-mer tables of a database (new2all, all2all-parts, build -extend) are expanded to slot arrays in parallel
    from a mapping of the file; same slots, same patterns after them and same refusals as the sequential reader — on the
    reference-built fixtures (256 and 65 536 tables) and on a k = 25 database of ours (262 144 tables)."""
    ours = tmp_path / "k25.db"
    cli(ref_fixtures, "build", "-host-build", "-k", "25", "-f", "0.1", "test/virus/seqs.part1.list", ours)
    for db in (golden_dbs["virus.k18"][0], golden_dbs["virus.k24"][0], golden_dbs["synth.k21"][0], ours):
        got = {}
        for mode in ("mapped", "stream"):
            if mode == "stream":
                monkeypatch.setenv("KDBX_DB_READER", "stream")
            else:
                monkeypatch.delenv("KDBX_DB_READER", raising=False)
            t = libs.Trie.read_db_full(db)
            got[mode] = (_tables_of(t), {k: np.array(x, copy=True) for k, x in t.arrays().items()})
            t.close()
        monkeypatch.delenv("KDBX_DB_READER", raising=False)
        (oa, sa), pa = got["mapped"]
        (ob, sb), pb = got["stream"]
        assert np.array_equal(oa, ob) and np.array_equal(sa, sb), db
        assert int((sa >> np.uint64(32) != np.uint64(0x7FFFFFFF)).sum()) > 0
        for k in pa:
            assert np.array_equal(pa[k], pb[k]), (db, k)
    # a table section cut short or with an impossible header: both readers refuse
    raw = golden_dbs["virus.k18"][0].read_bytes()
    probe = libs.Trie.read_db(golden_dbs["virus.k18"][0])
    names = sum(16 + len(n) for n in probe.sample_names())
    first_table = 8 + 4 + 8 + 8 + 4 + 1 + 8 + 8 + names + 8      # header fields, sample table, table count
    assert raw[first_table - 8:first_table] == np.uint64(256).tobytes()
    for what, data in (("cut inside the tables", raw[:first_table + 1000]),
                       ("allocated is not a power of two", raw[:first_table + 16] + np.uint64(48).tobytes() + raw[first_table + 24:])):
        f = tmp_path / "bad.db"
        f.write_bytes(data)
        for mode in ("mapped", "stream"):
            if mode == "stream":
                monkeypatch.setenv("KDBX_DB_READER", "stream")
            else:
                monkeypatch.delenv("KDBX_DB_READER", raising=False)
            with pytest.raises(libs.KdbxError):
                libs.Trie.read_db_full(f)
        monkeypatch.delenv("KDBX_DB_READER", raising=False)
