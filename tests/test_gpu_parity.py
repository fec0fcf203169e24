"""Parity of the CUDA path (through the C ABI) with the oracle.  Needs a B200: -m gpu."""
import subprocess

import numpy as np
import pytest

import oracle_util as ou

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(libs):
    c = libs.Context(device=0)
    yield c
    c.close()


def _run(libs, N, a, **cfg):
    with libs.Context(device=0, **cfg) as c:
        v, keep = libs.view_from_arrays(N, a["num_kmers"], a["parent_id"], a["n"], a["l"], a["last"], a["bits"], a["payload_off"],
                                        a["payload"])
        c.load_patterns(v, keep)
        out, st = c.all2all_dense()
        return out, st


@pytest.mark.parametrize("name", ["virus.k18", "virus.k18.f01", "virus.k24", "synth.k21"])
def test_golden_databases_csv_bytes(libs, ctx, golden_dbs, tmp_path, name):
    db, dense, sparse = golden_dbs[name]
    t = libs.Trie.read_db(db)
    ctx.load_patterns(t)
    tri, st = ctx.all2all_dense()
    assert st.updates == t.totals().updates
    out = tmp_path / "g.csv"
    t.write_all2all_csv(tri, out)
    assert ou.read_bytes(out) == ou.read_bytes(dense)
    if sparse is not None:
        t.write_all2all_csv(tri, out, sparse=True)
        assert ou.read_bytes(out) == ou.read_bytes(sparse)


def test_cli_reproduces_golden_csv(libs, golden_dbs, tmp_path):
    db, dense, sparse = golden_dbs["virus.k18"]
    exe = ou.ROOT / "kmer-db_b200" / "bin" / "kmer-db-b200"
    subprocess.run([str(exe), "all2all", str(db), str(tmp_path / "a.csv")], check=True, stderr=subprocess.DEVNULL)
    assert ou.read_bytes(tmp_path / "a.csv") == ou.read_bytes(dense)
    subprocess.run([str(exe), "all2all", "-sparse", str(db), str(tmp_path / "s.csv")], check=True, stderr=subprocess.DEVNULL)
    assert ou.read_bytes(tmp_path / "s.csv") == ou.read_bytes(sparse)
    r = subprocess.run([str(exe), "all2all", str(tmp_path / "missing.db"), str(tmp_path / "x.csv")], capture_output=True, text=True)
    assert r.returncode != 0 and "ERROR: Cannot open k-mer database" in r.stderr


@pytest.mark.parametrize("seed", range(8))
def test_random_tries_bit_exact(libs, oracle, seed):
    rng = np.random.default_rng(100 + seed)
    N = int(rng.integers(2, 400))
    a, _ = ou.random_trie(rng, N, int(rng.integers(2, 1500)), max_local=int(rng.integers(1, 30)),
                          big_weights=(seed % 2 == 0), dense_lists=(seed % 3 == 0))
    want, U = ou.oracle_all2all(oracle, N, a)
    got, st = _run(libs, N, a)
    assert st.updates == U
    assert np.array_equal(got, want)


@pytest.mark.parametrize("cfg", [dict(tile_cols=32), dict(tile_cols=64, chunk_ids=4096), dict(tile_cols=128, unit_updates=64),
                                 dict(chunk_ids=5000, unit_updates=1), dict(tile_cols=1024, unit_updates=1 << 30),
                                 dict(tile_rows=1), dict(tile_rows=8, tile_cols=64), dict(tile_rows=16, scatter_threads=128),
                                 dict(tile_rows=2, tile_cols=160, scatter_threads=1024, chunk_ids=9000),
                                 dict(flags=1), dict(flags=1, chunk_ids=4096, tile_cols=96), dict(flags=1, tile_rows=4, unit_updates=50),
                                 dict(flags=2), dict(flags=3, chunk_ids=4096),
                                 dict(flags=4), dict(flags=8), dict(flags=8, unit_updates=1), dict(flags=8, tile_rows=8),
                                 dict(flags=8, scatter_threads=256), dict(flags=10, unit_updates=3000)])
def test_schedule_knobs_do_not_change_results(libs, oracle, cfg):
    """column tiles (T>1), row blocks of 1..32 rows, many chunks, tiny / huge work units: same bits."""
    rng = np.random.default_rng(7)
    N = 900
    a, _ = ou.random_trie(rng, N, 4000, max_local=25, big_weights=True)
    want, U = ou.oracle_all2all(oracle, N, a)
    got, st = _run(libs, N, a, **cfg)
    assert st.updates == U
    assert np.array_equal(got, want)
    if "chunk_ids" in cfg and cfg.get("flags"):
        assert st.chunks > 1
    # flags=1 (KDBX_FLAG_CHUNKED_LISTS) forces the chunked parent-chain expansion; the default here is
    # the resident level-ordered one; 4 = id lists, 8 = run-boundary lists (N <= 1536: one column window)
    if cfg.get("flags", 0) & 8 and "tile_rows" not in cfg:   # (8-row blocks: 113 row blocks, more than the decoder counts)
        assert st.list_form == 1 and st.physical_updates > 0
    if cfg.get("flags", 0) & 4:
        assert st.list_form == 0 and st.physical_updates == U


@pytest.mark.parametrize("seed", range(10))
def test_boundary_lists_bit_exact(libs, oracle, seed):
    """Run-boundary lists (csrc/diff.cuh) against the oracle AND the oracle's restatement of that form: same
    bits, the same number of difference updates; repeated calls on one context (the no-host-round-trip path:
    sizes and levels cached from the first call) give the same result."""
    rng = np.random.default_rng(500 + seed)
    N = int(rng.integers(2, 1500))
    max_local = int(rng.integers(1, 60))
    a, _ = ou.random_trie(rng, N, int(rng.integers(2, 3000)), max_local=max_local,
                          big_weights=(seed % 2 == 0), dense_lists=(seed % 3 != 0))
    want, U = ou.oracle_all2all(oracle, N, a)
    _, _, phys = ou.oracle_boundary(oracle, N, a)
    with libs.Context(device=0, flags=libs.FLAG_BOUNDARY_LISTS) as c:
        v, keep = libs.view_from_arrays(N, a["num_kmers"], a["parent_id"], a["n"], a["l"], a["last"], a["bits"], a["payload_off"],
                                        a["payload"])
        c.load_patterns(v, keep)
        for rep in range(3):
            got, st = c.all2all_dense()
            assert st.updates == U and st.list_form == 1
            assert st.physical_updates == phys
            assert np.array_equal(got, want), f"call {rep}"
    # the default picks the form from the decoded lists: consecutive ids -> boundaries
    got, st = _run(libs, N, a)
    assert np.array_equal(got, want) and st.updates == U
    if seed % 3 != 0 and N > 64 and max_local >= 12:   # (lists of one or two ids cost more entries as boundaries)
        assert st.list_form == 1 and st.physical_updates < U


def test_cached_metadata_follows_the_staged_trie(libs, oracle):
    """Sizes, levels and the list form are remembered per staged trie: loading another trie into the same
    context, or changing the row range, must not reuse them."""
    rng = np.random.default_rng(77)
    tries = []
    for N, P, dense in ((700, 2500, True), (300, 900, False), (1200, 4000, True)):
        a, _ = ou.random_trie(rng, N, P, max_local=30, big_weights=True, dense_lists=dense)
        tries.append((N, a, ou.oracle_all2all(oracle, N, a)))
    with libs.Context(device=0) as c:
        for rnd in range(2):
            for N, a, (want, U) in tries:
                v, keep = libs.view_from_arrays(N, a["num_kmers"], a["parent_id"], a["n"], a["l"], a["last"], a["bits"],
                                                a["payload_off"], a["payload"])
                c.load_patterns(v, keep)
                for _ in range(2):
                    got, st = c.all2all_dense()
                    assert st.updates == U and np.array_equal(got, want)
                half, _ = c.all2all_dense_rows(N // 3, N)
                assert np.array_equal(half, want[ou.tri_cells(N // 3):])
                got, st = c.all2all_dense()
                assert np.array_equal(got, want)


@pytest.mark.parametrize("lo,width,total", [(640, 300, 2000), (0, 1536, 4000), (2048, 1000, 3048), (100, 40, 200)])
def test_sample_window(libs, oracle, lo, width, total):
    """kdbx_set_sample_window: a trie whose samples lie in [lo, lo + width) of a larger sample table is planned like
    a database of `width` samples (one column window, boundary lists) and fills exactly its block of the matrix."""
    rng = np.random.default_rng(lo + width)
    a, _ = ou.random_trie(rng, width, 2000, max_local=25, big_weights=True, dense_lists=True)
    small, U = ou.oracle_all2all(oracle, width, a)
    b = dict(a)
    b["last"] = np.where(a["l"] > 0, a["last"] + lo, a["last"]).astype(np.uint32)
    want, U2 = ou.oracle_all2all(oracle, total, b)
    assert U2 == U
    with libs.Context(device=0) as c:
        v, keep = libs.view_from_arrays(total, b["num_kmers"], b["parent_id"], b["n"], b["l"], b["last"], b["bits"], b["payload_off"],
                                        b["payload"])
        c.load_patterns(v, keep)
        plain, st0 = c.all2all_dense()
        assert np.array_equal(plain, want) and st0.updates == U
        c.set_sample_window(lo, lo + width)
        for _ in range(2):
            got, st = c.all2all_dense()
            assert st.updates == U and np.array_equal(got, want)
            assert st.list_form == 1
        rows, _ = c.all2all_dense_rows(lo + 5, min(total, lo + width + 7))
        assert np.array_equal(rows, want[ou.tri_cells(lo + 5):ou.tri_cells(min(total, lo + width + 7))])
        if lo >= 32:
            c.set_sample_window(lo + 32, lo + width)
            with pytest.raises(libs.KdbxError, match="outside the declared sample window"):
                c.all2all_dense()
        with pytest.raises(libs.KdbxError, match="bad sample window"):
            c.set_sample_window(5, total + 1)


def test_edge_cases(libs, oracle):
    z = np.zeros
    # only the sentinel pattern; N = 0, 1, 2
    for N in (0, 1, 2):
        a = {"num_kmers": z(1, np.int64), "parent_id": np.full(1, -1, np.int64), "n": z(1, np.uint32), "l": z(1, np.uint32),
             "last": z(1, np.uint32), "bits": z(1, np.uint32), "payload_off": z(1, np.uint64), "payload": z(2, np.uint64)}
        got, st = _run(libs, N, a)
        assert got.size == ou.tri_cells(N) and not got.any() and st.updates == 0
    # one pattern holding every sample (maximum list length), N not a multiple of 32
    N = 1000
    w, nb = ou.gamma_encode([1] * (N - 1))
    a = {"num_kmers": np.array([0, 2**32 + 3], np.int64), "parent_id": np.array([-1, -1], np.int64),
         "n": np.array([0, N], np.uint32), "l": np.array([0, N], np.uint32), "last": np.array([0, N - 1], np.uint32),
         "bits": np.array([0, nb], np.uint32), "payload_off": z(2, np.uint64), "payload": np.concatenate([w, z(2, np.uint64)])}
    got, st = _run(libs, N, a)
    assert st.updates == N * (N - 1) // 2 and (got == 3).all()
    # dense payload (payload_off = NULL) gives the same result
    rng = np.random.default_rng(3)
    a, _ = ou.random_trie(rng, 300, 800, max_local=20)
    want, _ = ou.oracle_all2all(oracle, 300, a)
    with libs.Context(device=0) as c:
        v, keep = libs.view_from_arrays(300, a["num_kmers"], a["parent_id"], a["n"], a["l"], a["last"], a["bits"], None, a["payload"])
        c.load_patterns(v, keep)
        got, _ = c.all2all_dense()
    assert np.array_equal(got, want)


def test_malformed_and_misuse_are_errors(libs):
    with libs.Context(device=0) as c:
        with pytest.raises(libs.KdbxError, match="no patterns loaded"):
            c.num_samples = 4
            c.all2all_dense()
    rng = np.random.default_rng(1)
    a, _ = ou.random_trie(rng, 50, 60, max_local=8)
    bad = dict(a)
    bad["bits"] = a["bits"].copy()
    idx = int(np.argmax(a["l"] > 1))
    bad["bits"][idx] += 1
    with pytest.raises(libs.KdbxError, match="malformed trie: Elias-gamma"):
        _run(libs, 50, bad)
    bad = dict(a)
    bad["parent_id"] = a["parent_id"].copy()
    bad["parent_id"][5] = 40  # parent after child
    with pytest.raises(libs.KdbxError, match="malformed trie: parent_id"):
        _run(libs, 50, bad)
    bad = dict(a)
    bad["n"] = a["n"].copy()
    bad["n"][7] += 1
    with pytest.raises(libs.KdbxError, match="malformed trie: parent_id"):
        _run(libs, 50, bad)
    bad = dict(a)
    bad["last"] = a["last"].copy()
    child = int(np.argmax(a["parent_id"] > 0))
    bad["last"][child] = a["last"][int(a["parent_id"][child])]  # child list no longer after the parent's
    with pytest.raises(libs.KdbxError, match="malformed trie"):
        _run(libs, 50, bad)
    bad = dict(a)
    bad["payload_off"] = a["payload_off"].copy()
    bad["payload_off"][idx] = 10**9
    with pytest.raises(libs.KdbxError, match="malformed trie: payload offset"):
        _run(libs, 50, bad)
    with libs.Context(device=0) as c:
        v, keep = libs.view_from_arrays(50, a["num_kmers"], a["parent_id"], a["n"], a["l"], a["last"], a["bits"], a["payload_off"], a["payload"])
        c.load_patterns(v, keep)
        with pytest.raises(libs.KdbxError, match="bad row range"):
            c.all2all_dense_rows(10, 60)


def test_w_accumulation_and_decode_taps(libs, oracle):
    rng = np.random.default_rng(9)
    N = 200
    a, lists = ou.random_trie(rng, N, 3000, max_local=12, big_weights=True)
    with libs.Context(device=0) as c:
        v, keep = libs.view_from_arrays(N, a["num_kmers"], a["parent_id"], a["n"], a["l"], a["last"], a["bits"], a["payload_off"], a["payload"])
        c.load_patterns(v, keep)
        c.all2all_dense()
        W = c.debug_fetch(0, len(a["n"]), np.uint32)
        loff = c.debug_fetch(2, len(a["n"]) + 1, np.uint64)
        loc = c.debug_fetch(1, int(a["l"].sum()), np.uint32)
    assert np.array_equal(W, ou.host_w(a))
    assert np.array_equal(loff, np.concatenate([[0], np.cumsum(a["l"].astype(np.uint64))]))
    assert loc.tolist() == [x for ids in lists for x in ids]


def test_row_sharding_matches_full_matrix(libs, oracle):
    t = libs.Trie.synth(num_samples=300, num_clusters=3, genome_kmers=40000, seed=5, interleaved=True)
    N = t.num_samples
    want, U = ou.oracle_all2all(oracle, N, t.arrays())
    with libs.Context(device=0) as c:
        c.load_patterns(t)
        upd = c.row_updates()
        assert int(upd.sum()) == U
        for world in (2, 3, 8):
            b = libs.shard_rows_by_work(upd, world)
            parts, total = [], 0
            for r in range(world):
                out, st = c.all2all_dense_rows(b[r], b[r + 1])
                assert st.updates == int(upd[b[r]:b[r + 1]].sum())
                parts.append(out)
                total += st.updates
            assert total == U
            assert np.array_equal(np.concatenate(parts), want)


@pytest.mark.parametrize("flags", [0, 1])
def test_pattern_parts_sum_to_full_matrix(libs, oracle, flags):
    """kdbx_all2all_dense_part_device: the partial matrices of all parts add up (uint32) to the full one."""
    import torch
    t = libs.Trie.synth(num_samples=300, num_clusters=3, genome_kmers=40000, seed=6)
    N = t.num_samples
    want, U = ou.oracle_all2all(oracle, N, t.arrays())
    with libs.Context(device=0, chunk_ids=200000, flags=flags) as c:
        c.load_patterns(t)
        for parts in (1, 2, 5):
            acc = torch.zeros(ou.tri_cells(N), dtype=torch.int64, device="cuda:0")
            buf = torch.zeros(ou.tri_cells(N), dtype=torch.int32, device="cuda:0")
            total = 0
            for part in range(parts):
                st = c.all2all_dense_part_device(part, parts, buf.data_ptr())
                torch.cuda.synchronize()
                acc += buf.to(torch.int64) & 0xFFFFFFFF
                total += st.updates
                if parts > 1:
                    assert st.chunks > parts  # pattern parts always stream chunk by chunk
            assert total == U
            assert np.array_equal((acc & 0xFFFFFFFF).cpu().numpy().astype(np.uint32), want)
        with pytest.raises(libs.KdbxError, match="bad part"):
            c.all2all_dense_part_device(3, 3, buf.data_ptr())


@pytest.mark.parametrize("parts", [2, 3, 8])
def test_sub_tries_sum_to_full_matrix(libs, oracle, parts):
    """The multi-GPU sharding of bench.py on one device: every sub-trie of kdbxh_partition runs through the
    whole pipeline; the partial matrices add up (uint32) to the oracle's matrix of the whole database.
    The relabelled shard (weak scaling) fills exactly its diagonal block of the larger matrix."""
    t = libs.Trie.synth(num_samples=300, num_clusters=3, genome_kmers=40000, seed=6)
    N = t.num_samples
    want, U = ou.oracle_all2all(oracle, N, t.arrays())
    acc = np.zeros_like(want)
    owned_total = 0
    with libs.Context(device=0) as c:
        for r in range(parts):
            sub, owned = t.partition(parts, r)
            c.load_patterns(sub)
            tri, st = c.all2all_dense()
            assert st.updates == sub.totals().updates
            acc += tri
            owned_total += owned
        assert owned_total == U and np.array_equal(acc, want)
        off, total = 300 * (parts - 1), 300 * parts
        t.relabel(off, total)
        c.load_patterns(t)
        big, st = c.all2all_dense()
        assert st.updates == U
        for s in (1, 2, 150, 299):
            o = ou.tri_cells(s + off)
            assert np.array_equal(big[o + off:o + off + s], want[ou.tri_cells(s):ou.tri_cells(s) + s])
            assert not big[o:o + off].any()
        assert int(big.astype(np.int64).sum()) == int(want.astype(np.int64).sum())


def test_generated_db_matches_reference_binary(libs, ref_bin, tmp_path):
    """Same .db through the unmodified reference (all host cores) and through the GPU path: cmp."""
    if ref_bin is None:
        pytest.skip("reference binary not built (oracle/_ref)")
    t = libs.Trie.synth(num_samples=400, num_clusters=4, genome_kmers=200000, seed=21)
    db = tmp_path / "g.db"
    t.write_db(db)
    subprocess.run([str(ref_bin), "all2all", str(db), str(tmp_path / "ref.csv")], check=True, stdout=subprocess.DEVNULL,
                   stderr=subprocess.DEVNULL)
    with libs.Context(device=0) as c:
        c.load_patterns(t)
        tri, st = c.all2all_dense()
    assert st.updates == t.totals().updates
    t.write_all2all_csv(tri, tmp_path / "gpu.csv")
    assert ou.read_bytes(tmp_path / "gpu.csv") == ou.read_bytes(tmp_path / "ref.csv")


@pytest.mark.parametrize("tile_cols", [0, 2048])
def test_large_n(libs, oracle, tile_cols):
    """N = 5000: one 5024-column tile per row by default; 2048-column tiles force the
    (row, tile) decomposition and the global-atomics bucketing path (15000 keys)."""
    t = libs.Trie.synth(num_samples=5000, num_clusters=10, genome_kmers=3000, seed=8, mutation_rate=0.01)
    N = t.num_samples
    want, U = ou.oracle_all2all(oracle, N, t.arrays())
    with libs.Context(device=0, tile_cols=tile_cols) as c:
        c.load_patterns(t)
        got, st = c.all2all_dense()
    assert st.updates == U
    assert np.array_equal(got, want)


def test_full_size_properties(libs):
    """A BASELINE-shaped cluster (250 genomes x 5 Mbp, one of config 2's four): results must not
    depend on the schedule (chunking / unit size), every cell is bounded by the genome size,
    row shards tile the matrix, and the update count equals U from the trie."""
    t = libs.Trie.synth(num_samples=250, num_clusters=1, genome_kmers=5_000_000, seed=2, pinned=False)
    tot = t.totals()
    with libs.Context(device=0) as c:
        c.load_patterns(t)
        a, st = c.all2all_dense()
        assert st.updates == tot.updates
        upd = c.row_updates()
        assert int(upd.sum()) == tot.updates
        b = libs.shard_rows_by_work(upd, 4)
        parts = [c.all2all_dense_rows(b[r], b[r + 1])[0] for r in range(4)]
    assert np.array_equal(np.concatenate(parts), a)
    assert int(a.max()) <= 5_000_000
    with libs.Context(device=0, chunk_ids=1 << 26, unit_updates=4096) as c:
        c.load_patterns(t)
        b2, _ = c.all2all_dense()
    assert np.array_equal(a, b2)
    # checksum of checksums: sum of the matrix == sum over jobs of W*i, computed from the device taps
    # (W, decoded locals) with numpy only


def test_reduce_scatter_entry_points_one_rank(libs, oracle):
    """kdbx_all2all_dense_reduce_scatter[_device] on a context without a communicator = one rank owning every cell."""
    import torch
    t = libs.Trie.synth(num_samples=200, num_clusters=2, genome_kmers=30000, seed=4)
    want, U = ou.oracle_all2all(oracle, 200, t.arrays())
    with libs.Context(device=0) as c:
        c.load_patterns(t)
        out = np.zeros(ou.tri_cells(200), np.uint32)
        first, count, st = c.all2all_dense_reduce_scatter(out)
        assert (first, count) == (0, ou.tri_cells(200)) and st.updates == U and np.array_equal(out, want)
        d = torch.zeros(ou.tri_cells(200), dtype=torch.int32, device="cuda:0")
        first, count, st = c.all2all_dense_reduce_scatter_device(d.data_ptr())
        torch.cuda.synchronize()
        assert np.array_equal(d.cpu().numpy().view(np.uint32), want)


def test_reduce_scatter_two_gpus_one_process(libs, oracle):
    """kdbx_comm_init_all + one host thread per device: the CLI's -gpus path.  Needs two B200s (skipped on one)."""
    import threading
    import ctypes as C
    k, _ = libs.load()
    if k.kdbx_device_count() < 2:
        pytest.skip("needs two GPUs")
    t = libs.Trie.synth(num_samples=300, num_clusters=5, genome_kmers=30000, seed=9, cluster_skew=0.4)
    N = 300
    want, U = ou.oracle_all2all(oracle, N, t.arrays())
    ctxs = [libs.Context(device=g) for g in range(2)]
    arr = (C.c_void_p * 2)(*[c._p for c in ctxs])
    assert k.kdbx_comm_init_all(arr, 2) == 0
    parts = list(t.partition_all(2))
    B = (ou.tri_cells(N) + 1) // 2
    outs, errs = [np.zeros(B, np.uint32) for _ in range(2)], []

    def work(g):
        try:
            ctxs[g].comm_nranks, ctxs[g].comm_rank = 2, g
            ctxs[g].load_patterns(parts[g][0])
            ctxs[g].set_sample_window(*parts[g][2])
            for _ in range(2):
                outs[g][:] = 0
                first, count, st = ctxs[g].all2all_dense_reduce_scatter(outs[g])
                assert first == g * B and st.updates == parts[g][0].totals().updates
        except Exception as e:  # noqa: BLE001
            errs.append(e)
    th = [threading.Thread(target=work, args=(g,)) for g in range(2)]
    [x.start() for x in th]
    [x.join() for x in th]
    assert not errs, errs
    got = np.concatenate(outs)[:ou.tri_cells(N)]
    assert np.array_equal(got, want)
    for c in ctxs:
        c.close()


def test_cli_all2all_on_two_gpus(libs, golden_dbs, tmp_path):
    """kmer-db-b200 all2all -gpus 2: sub-tries + one reduce-scatter, same CSV bytes as the reference's golden file."""
    k, _ = libs.load()
    if k.kdbx_device_count() < 2:
        pytest.skip("needs two GPUs")
    db, dense, _ = golden_dbs["virus.k18"]
    exe = ou.ROOT / "kmer-db_b200" / "bin" / "kmer-db-b200"
    subprocess.run([str(exe), "all2all", "-gpus", "2", str(db), str(tmp_path / "a.csv")], check=True, stderr=subprocess.DEVNULL)
    assert ou.read_bytes(tmp_path / "a.csv") == ou.read_bytes(dense)


@pytest.mark.parametrize("N", [1025, 1728, 1729, 2049, 3072, 3073])
def test_sample_counts_around_the_window_sizes(libs, oracle, N):
    """N around the tile limit: up to 1728 samples a tile holds whole rows (boundary lists, decoder job counts); beyond,
    these random tries (lists reaching anywhere) fail the locality test of the sliding window and run the id form."""
    rng = np.random.default_rng(N)
    a, _ = ou.random_trie(rng, N, 2500, max_local=40, big_weights=True, dense_lists=(N % 2 == 1))
    want, U = ou.oracle_all2all(oracle, N, a)
    got, st = _run(libs, N, a)
    assert st.updates == U and np.array_equal(got, want)
    got, st = _run(libs, N, a, flags=libs.FLAG_BOUNDARY_LISTS)
    assert st.updates == U and np.array_equal(got, want)
    if N <= 1728:
        assert st.list_form == 1
    elif N >= 2049:   # (just above 1728 hardly any list reaches below the sliding window)
        assert st.list_form == 0


def test_pattern_count_guard(libs):
    """num_patterns must stay below 2^31 (pattern ids are int32 in the node records): refused before anything is copied."""
    z = np.zeros
    a = {"num_kmers": z(1, np.int64), "parent_id": np.full(1, -1, np.int64), "n": z(1, np.uint32), "l": z(1, np.uint32),
         "last": z(1, np.uint32), "bits": z(1, np.uint32), "payload_off": z(1, np.uint64), "payload": z(2, np.uint64)}
    with libs.Context(device=0) as c:
        v, keep = libs.view_from_arrays(4, a["num_kmers"], a["parent_id"], a["n"], a["l"], a["last"], a["bits"], a["payload_off"], a["payload"])
        for bad in (0, 2**31, 2**40):
            v.num_patterns = bad
            with pytest.raises(libs.KdbxError, match=r"num_patterns must be in \[1, 2\^31\)"):
                c.load_patterns(v, keep)


def test_unaligned_device_output_takes_the_id_form(libs, oracle):
    """The boundary form flushes its tiles with 16-byte bulk reductions; an output pointer that is not 16-byte aligned
    gets the id form instead (same bits), also when the aligned call before it cached the boundary form."""
    import torch
    rng = np.random.default_rng(91)
    N = 500
    a, _ = ou.random_trie(rng, N, 3000, max_local=30, big_weights=True, dense_lists=True)
    want, U = ou.oracle_all2all(oracle, N, a)
    cells = ou.tri_cells(N)
    with libs.Context(device=0) as c:
        v, keep = libs.view_from_arrays(N, a["num_kmers"], a["parent_id"], a["n"], a["l"], a["last"], a["bits"], a["payload_off"], a["payload"])
        c.load_patterns(v, keep)
        buf = torch.zeros(cells + 8, dtype=torch.int32, device="cuda:0")
        for shift, form in ((0, 1), (0, 1), (1, 0), (3, 0), (4, 1)):
            st = c.all2all_dense_rows_device(0, N, buf[shift:].data_ptr())
            torch.cuda.synchronize()
            assert st.list_form == form and st.updates == U
            assert np.array_equal(buf[shift:shift + cells].cpu().numpy().view(np.uint32), want)


@pytest.mark.parametrize("N,clusters,mu,flags,form", [(4000, 16, 0.002, 0, 1), (6000, 5, 0.002, 0, 1), (4000, 2, 0.002, 0, 0),
                                                   (9000, 30, 0.002, 0, 1), (4000, 16, 0.01, 0, 0), (4000, 16, 0.01, "boundary", 1)])
def test_sliding_column_window(libs, oracle, N, clusters, mu, flags, form):
    """Beyond 1728 samples: one column window per row block that slides with the diagonal (make_plan).  Databases whose
    lists stay within 1728 ids below their rows (clusters up to that size) keep the headline path — boundary lists, decoder
    job counts — at any N up to 16384; a database with wider lists (two clusters of 2000) is re-planned with the id form,
    and so is one whose lists are too gappy for the boundary form to pay (mu = 0.01: 1.15 boundary entries per id,
    computed on the host from the decoded lists) unless the caller insists on it."""
    t = libs.Trie.synth(num_samples=N, num_clusters=clusters, genome_kmers=2500, seed=N + clusters, mutation_rate=mu)
    want, U = ou.oracle_all2all(oracle, N, t.arrays())
    cfg = {"flags": libs.FLAG_BOUNDARY_LISTS} if flags == "boundary" else {}
    with libs.Context(device=0, **cfg) as c:
        c.load_patterns(t)
        for _ in range(2):   # second call: cached sizes, no host round trips
            got, st = c.all2all_dense()
            assert st.updates == U and st.list_form == form
            assert np.array_equal(got, want)


@pytest.mark.parametrize("chunk_bytes", [0, 4096, 100_000])
def test_asynchronous_chunked_upload(libs, oracle, chunk_bytes):
    """KDBX_FLAG_ASYNC_UPLOAD: kdbx_load_patterns returns with the copies in flight — headers first, then the payload in up
    to eight chunks — and the decoder is launched once per chunk on the blocks whose payload that chunk completes.  Same
    bits as the synchronous path, for every chunk size, on re-staged tries of different sizes (buffers and events reused)."""
    with libs.Context(device=0, flags=libs.FLAG_ASYNC_UPLOAD, upload_chunk_bytes=chunk_bytes) as c:
        for N, gk, seed in ((300, 4000, 5), (700, 1500, 6), (300, 4000, 5)):
            t = libs.Trie.synth(num_samples=N, num_clusters=3, genome_kmers=gk, seed=seed, mutation_rate=0.004, pinned=True)
            want, U = ou.oracle_all2all(oracle, N, t.arrays())
            c.load_patterns(t)
            got, st = c.all2all_dense()
            assert st.updates == U and np.array_equal(got, want)
            got, st = c.all2all_dense()      # (resident: nothing in flight, one decoder launch)
            assert st.updates == U and np.array_equal(got, want)
            c.load_patterns(t)               # staged again while nothing else runs
            rp, col, val, st = c.all2all_sparse()
            assert int(rp[-1]) == int(np.count_nonzero(want))
