"""The device-side `build` (kdbx_builder_*, csrc/build.cuh) against the host builder — itself pinned on the
reference's CI vectors by tests/test_cli_host.py — and against plain set arithmetic.  Needs a B200: -m gpu."""
import ctypes as C

import numpy as np
import pytest

import oracle_util as ou

pytestmark = pytest.mark.gpu


def _random_sets(rng, N, universe_size=3000, bits=40):
    universe = np.unique(rng.integers(0, 1 << bits, size=universe_size, dtype=np.uint64))
    sets = []
    for s in range(N):
        base = sets[int(rng.integers(0, s))] if s and rng.random() < 0.7 else universe[rng.random(universe.size) < 0.2]
        keep = base[rng.random(base.size) < 0.9]
        extra = universe[rng.random(universe.size) < 0.02]
        sets.append(np.unique(np.concatenate([keep, extra])))
    return sets


def _table_dict(slot_off, slots):
    d = {}
    for t in range(len(slot_off) - 1):
        seg = slots[int(slot_off[t]):int(slot_off[t + 1])]
        size = len(seg)
        assert size >= 16 and size & (size - 1) == 0
        used = seg[(seg >> np.uint64(32)) != np.uint64(0x7FFFFFFF)]
        assert len(used) <= 0.8 * size
        for s in used.tolist():
            d[(t << 32) | (s & 0xFFFFFFFF)] = s >> 32
    return d


def _host_tables(libs, trie):
    v = trie.tables_view()
    T = int(v.num_tables)
    off = libs._np_from(v.slot_off, T + 1, np.uint64)
    return off, libs._np_from(v.slots, int(off[-1]), np.uint64)


def _assert_same_trie(got, want):
    for key in ("num_kmers", "parent_id", "n", "l", "last", "bits", "payload_off"):
        assert np.array_equal(got[key], want[key]), key
    words = len(want["payload"])
    assert np.array_equal(got["payload"][:words], want["payload"])


@pytest.mark.parametrize("seed", range(5))
def test_device_builder_equals_host_builder(libs, oracle, seed):
    rng = np.random.default_rng(40 + seed)
    N = int(rng.integers(2, 60))
    sets = _random_sets(rng, N)
    if seed == 1:
        sets[1] = np.zeros(0, np.uint64)   # an empty sample is registered but adds nothing
    host = libs.Trie.build([(f"s{i}", k) for i, k in enumerate(sets)], k=20)
    with libs.Context(device=0) as ctx, libs.DeviceBuilder(ctx, k=20) as b:
        for k in sets:
            b.add_kmers(k)
        a, slot_off, slots, filled, r = b.finish()
        assert r.num_samples == N and r.kmers_count == len(np.unique(np.concatenate(sets)))
        _assert_same_trie(a, host.arrays())
        assert int(filled.sum()) == r.kmers_count
        # every k-mer points at the pattern whose sample list is exactly the set of samples holding it
        hoff, hslots = _host_tables(libs, host)
        assert _table_dict(slot_off, slots) == _table_dict(hoff, hslots)
        # the exported tables answer queries through the probe kernel (same hash, linear probing)
        v, keep = libs.view_from_arrays(N, a["num_kmers"], a["parent_id"], a["n"], a["l"], a["last"], a["bits"], a["payload_off"], a["payload"])
        ctx.load_patterns(v, keep)
        tv = libs.TablesView(len(slot_off) - 1, slot_off.ctypes.data, slots.ctypes.data)
        ctx.load_hashtables(tv)
        out, _ = ctx.new2all_batch(sets[:8])
        for q in range(min(8, N)):
            for s in range(N):
                assert out[q, s] == np.intersect1d(sets[q], sets[s], assume_unique=True).size
        tri, _ = ctx.all2all_dense()
    want, _ = ou.oracle_all2all(oracle, N, host.arrays())
    assert np.array_equal(tri, want)


def _write_fasta(path, records, width=70):
    with open(path, "w") as f:
        for name, seq in records:
            f.write(f">{name} some description\n")
            for i in range(0, len(seq), width):
                f.write(seq[i:i + width] + "\n")


@pytest.mark.parametrize("k,fraction", [(18, 1.0), (18, 0.3), (24, 1.0), (15, 1.0), (25, 0.5)])
def test_device_extraction_equals_host_extraction(libs, tmp_path, k, fraction):
    """Sequences with repeats, lower case, Ns and several records per sample: the database built from raw
    symbols on the device equals the one the host builder makes from the host's k-mer extraction."""
    rng = np.random.default_rng(k * 100 + int(fraction * 10))
    base = "".join(rng.choice(list("ACGT"), size=6000))
    seqs = []
    for s in range(12):
        x = list(base if s % 3 else "".join(rng.choice(list("ACGT"), size=5000)))
        for pos in rng.integers(0, len(x), size=40):
            x[pos] = "ACGT"[int(rng.integers(0, 4))]
        for pos in rng.integers(0, len(x), size=3):
            x[pos] = "N"
        x = "".join(x)
        if s % 2:
            x = x[:1000].lower() + x[1000:]
        recs = [(f"r{s}_0", x[:2500]), (f"r{s}_1", x[2500:] + x[:300]), (f"r{s}_2", "ACG")]   # the last one is shorter than k
        seqs.append(recs)
        _write_fasta(tmp_path / f"s{s}.fa", recs)
    (tmp_path / "list.txt").write_text("\n".join(str(tmp_path / f"s{s}.fa") for s in range(12)) + "\n")
    host_samples = libs.load_samples(tmp_path / "list.txt", k=k, fraction=fraction)
    host = libs.Trie.build(host_samples, k=k, fraction=fraction)
    with libs.Context(device=0) as ctx, libs.DeviceBuilder(ctx, k=k, fraction=fraction) as b:
        for s, recs in enumerate(seqs):
            symbols = b"".join(seq.encode() + b"\0" for _, seq in recs)
            assert b.add_sequence(symbols) == len(host_samples[s][1])
        a, slot_off, slots, filled, r = b.finish()
    _assert_same_trie(a, host.arrays())
    hoff, hslots = _host_tables(libs, host)
    assert _table_dict(slot_off, slots) == _table_dict(hoff, hslots)


def test_device_builder_grows_its_table_and_adopts_a_database(libs, oracle):
    rng = np.random.default_rng(5)
    universe = np.unique(rng.integers(0, 1 << 36, size=400000, dtype=np.uint64))
    sets = [universe[rng.random(universe.size) < 0.25] for _ in range(10)]
    host_all = libs.Trie.build([(f"s{i}", k) for i, k in enumerate(sets)], k=18)
    with libs.Context(device=0) as ctx:
        with libs.DeviceBuilder(ctx, k=18) as b:
            for k in sets:
                b.add_kmers(k)
            a, slot_off, slots, filled, r = b.finish()
        assert r.table_growths >= 2
        _assert_same_trie(a, host_all.arrays())
        # -extend: stage the database of the first 6 samples, adopt it, add the rest
        host_part = libs.Trie.build([(f"s{i}", k) for i, k in enumerate(sets[:6])], k=18)
        ctx.load_patterns(host_part)
        ctx.load_hashtables(host_part)
        with libs.DeviceBuilder(ctx, k=18) as b:
            b.adopt()
            for k in sets[6:]:
                b.add_kmers(k)
            a2, slot_off2, slots2, _, r2 = b.finish()
        assert r2.num_samples == 10
        _assert_same_trie(a2, host_all.arrays())
        assert _table_dict(slot_off2, slots2) == _table_dict(slot_off, slots)


def test_device_builder_misuse(libs):
    with libs.Context(device=0) as ctx:
        with libs.DeviceBuilder(ctx, k=18) as b:
            with pytest.raises(libs.KdbxError, match="ascend"):
                b.add_kmers(np.array([5, 3, 9], np.uint64))
            with pytest.raises(libs.KdbxError, match="ascend"):
                b.add_kmers(np.array([1, 1 << 62], np.uint64))   # wider than an 18-mer
            b.add_kmers(np.array([3, 5, 9], np.uint64))
            assert b.add_sequence(b"ACGTACGT") == 0               # shorter than k: an empty sample
            a, *_rest, r = b.finish()
            assert r.num_samples == 2 and r.num_patterns == 2 and r.kmers_count == 3
            with pytest.raises(libs.KdbxError, match="finished"):
                b.add_kmers(np.array([4], np.uint64))
        with pytest.raises(libs.KdbxError, match="kdbx_builder_open"):
            libs.DeviceBuilder(ctx, k=40)
        with libs.DeviceBuilder(ctx, k=18) as b:
            with pytest.raises(libs.KdbxError, match="stage the database"):
                b.adopt()


@pytest.mark.parametrize("k,fraction", [(18, 1.0), (21, 0.4)])
def test_new2all_from_sequences_equals_new2all_from_host_kmers(libs, tmp_path, k, fraction):
    """kdbx_new2all_sequences (extraction, minhash, sort, unique on the device) gives the rows and the k-mer counts
    of kdbx_new2all_batch fed with the host's extraction of the same files; sub-batching does not change them."""
    rng = np.random.default_rng(77 + k)
    base = "".join(rng.choice(list("ACGT"), size=8000))
    files, seqs = [], []
    for s_ in range(14):
        x = list(base)
        for pos in rng.integers(0, len(x), size=60 * (s_ + 1)):
            x[pos] = "ACGT"[int(rng.integers(0, 4))]
        x = "".join(x)
        recs = [(f"q{s_}a", x[:5000]), (f"q{s_}b", x[5000:] + "NN" + x[:200].lower())]
        _write_fasta(tmp_path / f"q{s_}.fa", recs)
        files.append(str(tmp_path / f"q{s_}.fa"))
        seqs.append(b"".join(r.encode() + b"\0" for _, r in recs))
    seqs.append(b"ACGT")     # shorter than k: no k-mers
    (tmp_path / "list.txt").write_text("\n".join(files) + "\n")
    samples = libs.load_samples(tmp_path / "list.txt", k=k, fraction=fraction)
    db = libs.Trie.build(samples[:10], k=k, fraction=fraction)
    want_sets = [s[1] for s in samples] + [np.zeros(0, np.uint64)]
    for batch in (0, 4000):
        with libs.Context(device=0, query_batch_kmers=batch) as ctx:
            ctx.load_patterns(db)
            ctx.load_hashtables(db)
            want, _ = ctx.new2all_batch(want_sets)
            got, uniq, st = ctx.new2all_sequences(seqs, k=k, fraction=fraction)
        assert uniq.tolist() == [len(s) for s in want_sets]
        assert np.array_equal(got, want)
        assert st.probes == sum(len(s) for s in want_sets)
    with libs.Context(device=0) as ctx:
        ctx.load_patterns(db)
        ctx.load_hashtables(db)
        with pytest.raises(libs.KdbxError, match="k-mer tables"):
            ctx.new2all_sequences(seqs, k=22 if k != 22 else 23)     # another number of prefix tables than the database has
