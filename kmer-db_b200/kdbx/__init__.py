"""ctypes binding of the two C-ABI libraries of this repository (include/kdbx.h, include/kdbx_host.h).

This is the Python-side mirror that bench.py and the tests use; the product itself is the C++
host code in kmer-db_b200/host plus the CUDA library.  Nothing here computes anything: every
function forwards to libkdbx.so (CUDA, sm_100a) or libkdbx_host.so and raises ``KdbxError`` on a
non-zero return code — exactly as the C++ mirror rethrows ``std::runtime_error``
(kmer-db_b200/host/similarity_calculator.h; reference behaviour: src/main.cpp:56-59).
There is no CPU fallback: without the built libraries ``load()`` raises, and without a B200
``Context()`` raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_PKG = Path(__file__).resolve().parent.parent
LIB_DIR = _PKG / "lib"


class KdbxError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("flags", C.c_uint32), ("chunk_ids", C.c_uint64),
                ("tile_cols", C.c_uint32), ("unit_updates", C.c_uint32), ("sparse_block_cells", C.c_uint64),
                ("query_batch_kmers", C.c_uint64), ("tile_rows", C.c_uint32), ("scatter_threads", C.c_uint32),
                ("upload_chunk_bytes", C.c_uint64)]


class TrieView(C.Structure):
    _fields_ = [("num_patterns", C.c_uint64), ("num_samples", C.c_uint32), ("_pad", C.c_uint32),
                ("num_kmers", C.c_void_p), ("parent_id", C.c_void_p), ("num_samples_full", C.c_void_p),
                ("num_local_samples", C.c_void_p), ("last_sample_id", C.c_void_p), ("num_bits", C.c_void_p),
                ("payload_off", C.c_void_p), ("payload", C.c_void_p), ("payload_words", C.c_uint64),
                ("parent_id32", C.c_void_p), ("num_kmers32", C.c_void_p)]


class Stats(C.Structure):
    _fields_ = [("updates", C.c_uint64), ("jobs", C.c_uint64), ("flat_ids", C.c_uint64), ("local_ids", C.c_uint64),
                ("units", C.c_uint64), ("chunks", C.c_uint32), ("kernel_launches", C.c_uint32),
                ("ms_upload", C.c_float), ("ms_prepare", C.c_float), ("ms_expand", C.c_float), ("ms_bucket", C.c_float),
                ("ms_scatter", C.c_float), ("ms_total", C.c_float), ("ms_download", C.c_float),
                ("scatter_launches", C.c_uint32), ("_pad", C.c_uint32), ("probes", C.c_uint64), ("hits", C.c_uint64),
                ("ms_probe", C.c_float), ("ms_compact", C.c_float), ("physical_updates", C.c_uint64),
                ("list_form", C.c_uint32), ("ms_collective", C.c_float)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k not in ("_pad", "_pad2", "reserved")}


class MetricBound(C.Structure):
    _fields_ = [("metric", C.c_int32), ("_pad", C.c_int32), ("lo", C.c_double), ("hi", C.c_double)]


class Filter(C.Structure):
    _fields_ = [("min_common", C.c_uint32), ("max_common", C.c_uint32), ("num_metric_bounds", C.c_uint32), ("_pad", C.c_uint32),
                ("metric_bounds", MetricBound * 4), ("sample_kmers", C.c_void_p)]


class Csr(C.Structure):
    _fields_ = [("num_rows", C.c_uint32), ("_pad", C.c_uint32), ("nnz", C.c_uint64), ("row_ptr", C.c_void_p),
                ("col", C.c_void_p), ("val", C.c_void_p)]


class TablesView(C.Structure):
    _fields_ = [("num_tables", C.c_uint64), ("slot_off", C.c_void_p), ("slots", C.c_void_p)]


class BuildParams(C.Structure):
    _fields_ = [("kmer_length", C.c_uint32), ("bits_per_symbol", C.c_uint32), ("alphabet_size", C.c_uint32),
                ("preserve_strand", C.c_uint32), ("fraction", C.c_double), ("fraction_start", C.c_double),
                ("symbol_map", C.c_int8 * 256), ("table_capacity_hint", C.c_uint64)]


class BuildResult(C.Structure):
    _fields_ = [("num_patterns", C.c_uint64), ("payload_words", C.c_uint64), ("num_tables", C.c_uint64),
                ("total_slots", C.c_uint64), ("kmers_count", C.c_uint64), ("sum_local_samples", C.c_uint64),
                ("table_capacity", C.c_uint64), ("num_samples", C.c_uint32), ("kernel_launches", C.c_uint32),
                ("table_growths", C.c_uint32), ("ms_finish", C.c_float), ("reserved", C.c_uint64 * 2)]


class BuildArrays(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("num_kmers", "parent_id", "num_samples_full", "num_local_samples", "last_sample_id",
                                          "num_bits", "payload_off", "payload", "slot_off", "slots", "table_filled")]


METRICS = {"jaccard": 0, "min": 1, "max": 2, "cosine": 3}
FLAG_CHUNKED_LISTS = 1
FLAG_ASYNC_UPLOAD = 2
FLAG_ID_LISTS = 4
FLAG_BOUNDARY_LISTS = 8


class SynthParams(C.Structure):
    _fields_ = [("num_samples", C.c_uint32), ("num_clusters", C.c_uint32), ("genome_kmers", C.c_uint64),
                ("k", C.c_uint32), ("interleaved", C.c_int32), ("mutation_rate", C.c_double), ("seed", C.c_uint64),
                ("threads", C.c_int32), ("_pad", C.c_int32), ("cluster_skew", C.c_double)]


class Totals(C.Structure):
    _fields_ = [("num_patterns", C.c_uint64), ("num_samples", C.c_uint64), ("updates", C.c_uint64),
                ("sum_n", C.c_uint64), ("sum_l", C.c_uint64), ("payload_bytes", C.c_uint64), ("kmers_count", C.c_uint64),
                ("kmer_length", C.c_uint32), ("_pad", C.c_uint32), ("fraction", C.c_double)]


# every symbol include/kdbx.h declares (tests check that the library exports all of them)
KDBX_SYMBOLS = ["kdbx_abi_version", "kdbx_device_count", "kdbx_open", "kdbx_close", "kdbx_last_error",
                "kdbx_host_alloc", "kdbx_host_free", "kdbx_load_patterns", "kdbx_set_sample_window", "kdbx_row_updates", "kdbx_all2all_dense",
                "kdbx_all2all_dense_rows", "kdbx_all2all_dense_rows_device", "kdbx_all2all_dense_part_device", "kdbx_all2all_sparse", "kdbx_all2all_sparse_rows", "kdbx_db2db_sparse", "kdbx_csv_dense_rows", "kdbx_stage_matrix", "kdbx_distance_dense_rows", "kdbx_free_csr",
                "kdbx_load_hashtables", "kdbx_new2all_batch", "kdbx_debug_fetch", "kdbx_comm_unique_id", "kdbx_comm_init_rank",
                "kdbx_comm_init_all", "kdbx_comm_destroy", "kdbx_all2all_dense_reduce_scatter_device", "kdbx_all2all_dense_reduce_scatter",
                "kdbx_builder_open", "kdbx_builder_close", "kdbx_builder_adopt", "kdbx_builder_add_sequence", "kdbx_builder_add_kmers",
                "kdbx_builder_finish", "kdbx_builder_export", "kdbx_new2all_sequences"]
KDBXH_SYMBOLS = ["kdbxh_last_error", "kdbxh_trie_new", "kdbxh_trie_free", "kdbxh_read_db", "kdbxh_write_db",
                 "kdbxh_synth", "kdbxh_validate", "kdbxh_prefix", "kdbxh_partition", "kdbxh_partitioner_new", "kdbxh_partitioner_free",
                 "kdbxh_partitioner_part", "kdbxh_partition_write_all", "kdbxh_relabel", "kdbxh_view", "kdbxh_totals_of", "kdbxh_sample_name",
                 "kdbxh_sample_kmers", "kdbxh_write_all2all_csv", "kdbxh_write_sparse_csv", "kdbxh_write_one2all_csv", "kdbxh_read_db_full", "kdbxh_tables_view",
                 "kdbxh_builder_new", "kdbxh_builder_free", "kdbxh_builder_add_sample", "kdbxh_builder_finish",
                 "kdbxh_samples_load", "kdbxh_samples_free", "kdbxh_samples_count", "kdbxh_samples_name", "kdbxh_samples_kmers"]

_libs = None


def load():
    """dlopen libkdbx.so and libkdbx_host.so from kmer-db_b200/lib (built by `make`)."""
    global _libs
    if _libs is not None:
        return _libs
    so, hso = LIB_DIR / "libkdbx.so", LIB_DIR / "libkdbx_host.so"
    if not so.exists() or not hso.exists():
        raise KdbxError(f"{so} / {hso} not built: run `make -C {_PKG}` (or __graft_entry__.build()); "
                        "there is no fallback implementation")
    k = C.CDLL(str(so), mode=C.RTLD_GLOBAL)
    h = C.CDLL(str(hso), mode=C.RTLD_GLOBAL)
    P = C.POINTER
    k.kdbx_abi_version.restype = C.c_int
    k.kdbx_device_count.restype = C.c_int
    k.kdbx_open.argtypes = [P(Config), P(C.c_void_p)]
    k.kdbx_close.argtypes = [C.c_void_p]
    k.kdbx_close.restype = None
    k.kdbx_last_error.argtypes = [C.c_void_p]
    k.kdbx_last_error.restype = C.c_char_p
    k.kdbx_host_alloc.argtypes = [P(C.c_void_p), C.c_size_t]
    k.kdbx_host_free.argtypes = [C.c_void_p]
    k.kdbx_host_free.restype = None
    k.kdbx_load_patterns.argtypes = [C.c_void_p, P(TrieView)]
    k.kdbx_set_sample_window.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
    k.kdbx_row_updates.argtypes = [C.c_void_p, C.c_void_p]
    k.kdbx_all2all_dense.argtypes = [C.c_void_p, C.c_void_p, P(Stats)]
    k.kdbx_all2all_dense_rows.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, P(Stats)]
    k.kdbx_all2all_dense_rows_device.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, P(Stats)]
    k.kdbx_all2all_dense_part_device.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, P(Stats)]
    k.kdbx_csv_dense_rows.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p, P(C.c_uint64)]
    k.kdbx_stage_matrix.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
    k.kdbx_distance_dense_rows.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p, P(C.c_uint64)]
    k.kdbx_comm_unique_id.argtypes = [C.c_void_p]
    k.kdbx_comm_init_rank.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    k.kdbx_comm_init_all.argtypes = [P(C.c_void_p), C.c_int]
    k.kdbx_comm_destroy.argtypes = [C.c_void_p]
    k.kdbx_comm_destroy.restype = None
    k.kdbx_all2all_dense_reduce_scatter_device.argtypes = [C.c_void_p, C.c_void_p, P(C.c_uint64), P(C.c_uint64), P(Stats)]
    k.kdbx_all2all_dense_reduce_scatter.argtypes = [C.c_void_p, C.c_void_p, P(C.c_uint64), P(C.c_uint64), P(Stats)]
    k.kdbx_all2all_sparse.argtypes = [C.c_void_p, P(Filter), P(Csr), P(Stats)]
    k.kdbx_all2all_sparse_rows.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, P(Filter), P(Csr), P(Stats)]
    k.kdbx_db2db_sparse.argtypes = [C.c_void_p, C.c_void_p, P(Filter), C.c_void_p, P(Csr), P(Stats)]
    k.kdbx_free_csr.argtypes = [P(Csr)]
    k.kdbx_free_csr.restype = None
    k.kdbx_load_hashtables.argtypes = [C.c_void_p, P(TablesView)]
    k.kdbx_new2all_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, P(Stats)]
    k.kdbx_builder_open.argtypes = [C.c_void_p, P(BuildParams), P(C.c_void_p)]
    k.kdbx_builder_close.argtypes = [C.c_void_p]
    k.kdbx_builder_close.restype = None
    k.kdbx_builder_adopt.argtypes = [C.c_void_p]
    k.kdbx_builder_add_sequence.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, P(C.c_uint64)]
    k.kdbx_builder_add_kmers.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    k.kdbx_builder_finish.argtypes = [C.c_void_p, P(BuildResult)]
    k.kdbx_builder_export.argtypes = [C.c_void_p, P(BuildArrays)]
    k.kdbx_new2all_sequences.argtypes = [C.c_void_p, P(BuildParams), C.c_char_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, P(Stats)]
    k.kdbx_debug_fetch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64]
    k.kdbx_debug_fetch.restype = C.c_int64
    h.kdbxh_last_error.restype = C.c_char_p
    h.kdbxh_trie_new.argtypes = [C.c_int]
    h.kdbxh_trie_new.restype = C.c_void_p
    h.kdbxh_trie_free.argtypes = [C.c_void_p]
    h.kdbxh_trie_free.restype = None
    h.kdbxh_read_db.argtypes = [C.c_void_p, C.c_char_p]
    h.kdbxh_write_db.argtypes = [C.c_void_p, C.c_char_p]
    h.kdbxh_synth.argtypes = [C.c_void_p, P(SynthParams)]
    h.kdbxh_validate.argtypes = [C.c_void_p]
    h.kdbxh_prefix.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
    h.kdbxh_partition.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, P(C.c_uint64)]
    h.kdbxh_partitioner_new.argtypes = [C.c_void_p, C.c_uint32]
    h.kdbxh_partitioner_new.restype = C.c_void_p
    h.kdbxh_partitioner_free.argtypes = [C.c_void_p]
    h.kdbxh_partitioner_free.restype = None
    h.kdbxh_partitioner_part.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, P(C.c_uint64), P(C.c_uint32 * 2)]
    h.kdbxh_partition_write_all.argtypes = [C.c_void_p, C.c_uint32, C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    h.kdbxh_relabel.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
    h.kdbxh_view.argtypes = [C.c_void_p, P(TrieView)]
    h.kdbxh_totals_of.argtypes = [C.c_void_p, P(Totals)]
    h.kdbxh_sample_name.argtypes = [C.c_void_p, C.c_uint32]
    h.kdbxh_sample_name.restype = C.c_char_p
    h.kdbxh_sample_kmers.argtypes = [C.c_void_p, C.c_uint32]
    h.kdbxh_sample_kmers.restype = C.c_uint64
    h.kdbxh_write_all2all_csv.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p, C.c_int]
    h.kdbxh_write_one2all_csv.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_void_p, C.c_char_p]
    h.kdbxh_write_sparse_csv.argtypes = [C.c_void_p, P(Csr), C.c_void_p, C.c_void_p, C.c_uint32, C.c_char_p, C.c_char_p, C.c_char_p, P(C.c_uint64)]
    h.kdbxh_read_db_full.argtypes = [C.c_void_p, C.c_char_p]
    h.kdbxh_tables_view.argtypes = [C.c_void_p, P(TablesView)]
    h.kdbxh_builder_new.argtypes = [C.c_int]
    h.kdbxh_builder_new.restype = C.c_void_p
    h.kdbxh_builder_free.argtypes = [C.c_void_p]
    h.kdbxh_builder_free.restype = None
    h.kdbxh_builder_add_sample.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_double]
    h.kdbxh_builder_finish.argtypes = [C.c_void_p, C.c_void_p]
    h.kdbxh_samples_load.argtypes = [C.c_char_p, C.c_uint32, C.c_double, C.c_double, C.c_int32, C.c_int, C.c_int]
    h.kdbxh_samples_load.restype = C.c_void_p
    h.kdbxh_samples_free.argtypes = [C.c_void_p]
    h.kdbxh_samples_free.restype = None
    h.kdbxh_samples_count.argtypes = [C.c_void_p]
    h.kdbxh_samples_count.restype = C.c_uint32
    h.kdbxh_samples_name.argtypes = [C.c_void_p, C.c_uint32]
    h.kdbxh_samples_name.restype = C.c_char_p
    h.kdbxh_samples_kmers.argtypes = [C.c_void_p, C.c_uint32, P(C.c_uint64)]
    h.kdbxh_samples_kmers.restype = C.c_void_p
    _libs = (k, h)
    return _libs


def tri_cells(n: int) -> int:
    return n * (n - 1) // 2 if n > 0 else 0


def _np_from(ptr, count, dtype):
    if count == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_char * (count * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=count)


class Trie:
    """A kmer-db database as the all2all path needs it (owned by libkdbx_host.so)."""

    def __init__(self, pinned: bool = False):
        _, h = load()
        self._h = h
        self._p = h.kdbxh_trie_new(1 if pinned else 0)
        if not self._p:
            raise KdbxError("kdbxh_trie_new failed")
        self._keep = None

    def close(self):
        if self._p:
            self._h.kdbxh_trie_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise KdbxError(self._h.kdbxh_last_error().decode())

    @classmethod
    def read_db(cls, path, pinned=False):
        t = cls(pinned)
        t._check(t._h.kdbxh_read_db(t._p, os.fsencode(str(path))))
        return t

    @classmethod
    def read_db_full(cls, path, pinned=False):
        """With the k-mer tables (what new2all needs)."""
        t = cls(pinned)
        t._check(t._h.kdbxh_read_db_full(t._p, os.fsencode(str(path))))
        return t

    @classmethod
    def build(cls, samples, k=18, fraction=1.0, threads=1, pinned=False):
        """Database from [(name, sorted unique uint64 k-mers)] through the host builder."""
        _, h = load()
        b = h.kdbxh_builder_new(threads)
        t = cls(pinned)
        try:
            for name, kmers in samples:
                kmers = np.ascontiguousarray(kmers, np.uint64)
                t._check(h.kdbxh_builder_add_sample(b, name.encode(), kmers.ctypes.data if kmers.size else None, kmers.size, k, fraction))
            t._check(h.kdbxh_builder_finish(b, t._p))
        finally:
            h.kdbxh_builder_free(b)
        return t

    def tables_view(self) -> TablesView:
        v = TablesView()
        self._check(self._h.kdbxh_tables_view(self._p, C.byref(v)))
        return v

    def sample_kmer_counts(self):
        return np.array([self._h.kdbxh_sample_kmers(self._p, i) for i in range(self.num_samples)], dtype=np.uint64)

    @classmethod
    def synth(cls, num_samples, num_clusters=4, genome_kmers=5_000_000, k=18, mutation_rate=0.005, seed=2,
              interleaved=False, threads=0, pinned=False, cluster_skew=0.0):
        t = cls(pinned)
        sp = SynthParams(num_samples, num_clusters, genome_kmers, k, 1 if interleaved else 0, mutation_rate, seed, threads, 0, cluster_skew)
        t._check(t._h.kdbxh_synth(t._p, C.byref(sp)))
        return t

    def prefix(self, num_samples, pinned=False):
        """Sub-database of the first `num_samples` samples (cluster boundary of a generated db)."""
        t = Trie(pinned)
        self._check(self._h.kdbxh_prefix(self._p, num_samples, t._p))
        return t

    def partition(self, num_parts, part, pinned=False):
        """(sub-database of part `part`, U of the patterns it owns): see kdbxh_partition."""
        t = Trie(pinned)
        u = C.c_uint64()
        self._check(self._h.kdbxh_partition(self._p, num_parts, part, t._p, C.byref(u)))
        return t, int(u.value)

    def partition_all(self, num_parts, pinned=False):
        """Yields (part trie, owned updates, (lo, hi) sample window) for every part of one cut (kdbxh_partitioner_*)."""
        pt = self._h.kdbxh_partitioner_new(self._p, num_parts)
        if not pt:
            raise KdbxError(self._h.kdbxh_last_error().decode())
        try:
            for part in range(num_parts):
                out = Trie(pinned)
                owned = C.c_uint64(0)
                win = (C.c_uint32 * 2)(0, 0)
                self._check(self._h.kdbxh_partitioner_part(pt, part, out._p, C.byref(owned), C.byref(win)))
                yield out, int(owned.value), (int(win[0]), int(win[1]))
        finally:
            self._h.kdbxh_partitioner_free(pt)

    def partition_write_all(self, num_parts, prefix):
        """Cuts the database into num_parts parts and writes them to <prefix><part>of<num_parts>.db, one host thread per part;
        returns [(owned updates, (lo, hi) window, updates of the part, patterns of the part)]."""
        owned = np.zeros(num_parts, np.uint64); win = np.zeros(2 * num_parts, np.uint32)
        upd = np.zeros(num_parts, np.uint64); pats = np.zeros(num_parts, np.uint64)
        self._check(self._h.kdbxh_partition_write_all(self._p, num_parts, os.fsencode(str(prefix)), owned.ctypes.data, win.ctypes.data,
                                                      upd.ctypes.data, pats.ctypes.data))
        return [(int(owned[g]), (int(win[2 * g]), int(win[2 * g + 1])), int(upd[g]), int(pats[g])) for g in range(num_parts)]

    def relabel(self, offset, new_total):
        """Shift all sample ids by `offset` inside a table of `new_total` samples (in place)."""
        self._check(self._h.kdbxh_relabel(self._p, offset, new_total))

    def write_db(self, path):
        self._check(self._h.kdbxh_write_db(self._p, os.fsencode(str(path))))

    def validate(self):
        self._check(self._h.kdbxh_validate(self._p))

    def view(self) -> TrieView:
        v = TrieView()
        self._check(self._h.kdbxh_view(self._p, C.byref(v)))
        return v

    def totals(self) -> Totals:
        t = Totals()
        self._check(self._h.kdbxh_totals_of(self._p, C.byref(t)))
        return t

    @property
    def num_samples(self):
        return int(self.view().num_samples)

    @property
    def num_patterns(self):
        return int(self.view().num_patterns)

    def arrays(self):
        """Zero-copy numpy views of the SoA arrays (valid while the Trie lives)."""
        v = self.view()
        P = int(v.num_patterns)
        bits = _np_from(v.num_bits, P, np.uint32)
        if v.payload_off:
            payload_off = _np_from(v.payload_off, P, np.uint64)
        else:   # densely packed in pattern order (the view says so with NULL): ceil(bits / 128) * 2 words per pattern
            words = ((bits.astype(np.uint64) + np.uint64(127)) // np.uint64(128)) * np.uint64(2)
            payload_off = np.zeros(P, np.uint64)
            np.cumsum(words[:-1], out=payload_off[1:])
        return {
            "num_kmers": _np_from(v.num_kmers, P, np.int64), "parent_id": _np_from(v.parent_id, P, np.int64),
            "n": _np_from(v.num_samples_full, P, np.uint32), "l": _np_from(v.num_local_samples, P, np.uint32),
            "last": _np_from(v.last_sample_id, P, np.uint32), "bits": bits,
            "payload_off": payload_off,
            "payload": _np_from(v.payload, int(v.payload_words), np.uint64),
        }

    def sample_names(self):
        return [self._h.kdbxh_sample_name(self._p, i).decode() for i in range(self.num_samples)]

    def write_all2all_csv(self, tri: np.ndarray, path, sparse=False):
        tri = np.ascontiguousarray(tri, dtype=np.uint32)
        if tri.size != tri_cells(self.num_samples):
            raise KdbxError("matrix size does not match the sample count")
        self._check(self._h.kdbxh_write_all2all_csv(self._p, tri.ctypes.data, os.fsencode(str(path)), 1 if sparse else 0))

    def write_one2all_csv(self, sample, kmers, sims, path):
        sims = np.ascontiguousarray(sims, dtype=np.uint32)
        if sims.size != self.num_samples:
            raise KdbxError("one similarity per database sample is needed")
        self._check(self._h.kdbxh_write_one2all_csv(self._p, os.fsencode(str(sample)), int(kmers), sims.ctypes.data, os.fsencode(str(path))))

    def write_sparse_csv(self, cells, path, filters=None, sample_rows=None):
        """The all2all-sp / all2all-parts table from sparse rows (kdbxh_write_sparse_csv).  cells: [(row_shift, col_shift,
        row_ptr u64[rows+1], col u32[], val u32[])] — one cell with both shifts 0 is the whole matrix of all2all-sp;
        filters: '-min jaccard:0.9 -max 100'; sample_rows: '<criterion>:<count>'.  Returns the number of pairs written."""
        keep, arr = [], (Csr * len(cells))()
        rs = np.array([c[0] for c in cells], np.uint32); cs = np.array([c[1] for c in cells], np.uint32)
        for i, (_, _, row_ptr, col, val) in enumerate(cells):
            row_ptr = np.ascontiguousarray(row_ptr, np.uint64); col = np.ascontiguousarray(col, np.uint32); val = np.ascontiguousarray(val, np.uint32)
            keep.append((row_ptr, col, val))
            arr[i].num_rows = len(row_ptr) - 1; arr[i].nnz = int(row_ptr[-1])
            arr[i].row_ptr = row_ptr.ctypes.data; arr[i].col = col.ctypes.data; arr[i].val = val.ctypes.data
        saved = C.c_uint64(0)
        self._check(self._h.kdbxh_write_sparse_csv(self._p, arr, rs.ctypes.data, cs.ctypes.data, len(cells),
                                                   filters.encode() if filters else None, sample_rows.encode() if sample_rows else None,
                                                   os.fsencode(str(path)), C.byref(saved)))
        return int(saved.value)


def view_from_arrays(num_samples, num_kmers, parent_id, n, l, last, bits, payload_off, payload):
    """TrieView over caller-owned numpy arrays (kept alive by the returned tuple)."""
    arrs = [np.ascontiguousarray(num_kmers, np.int64), np.ascontiguousarray(parent_id, np.int64),
            np.ascontiguousarray(n, np.uint32), np.ascontiguousarray(l, np.uint32),
            np.ascontiguousarray(last, np.uint32), np.ascontiguousarray(bits, np.uint32),
            None if payload_off is None else np.ascontiguousarray(payload_off, np.uint64),
            np.ascontiguousarray(payload, np.uint64)]
    v = TrieView()
    v.num_patterns = len(arrs[0])
    v.num_samples = num_samples
    v.num_kmers, v.parent_id = arrs[0].ctypes.data, arrs[1].ctypes.data
    v.num_samples_full, v.num_local_samples = arrs[2].ctypes.data, arrs[3].ctypes.data
    v.last_sample_id, v.num_bits = arrs[4].ctypes.data, arrs[5].ctypes.data
    v.payload_off = None if arrs[6] is None else arrs[6].ctypes.data
    v.payload = arrs[7].ctypes.data if arrs[7].size else None
    v.payload_words = arrs[7].size
    return v, arrs


def load_samples(list_arg, k=18, fraction=1.0, fraction_start=0.0, alphabet_id=0, multisample=False, threads=4):
    """[(name, sorted unique k-mers)] of a sample list / FASTA file through the host ingest."""
    _, h = load()
    p = h.kdbxh_samples_load(os.fsencode(str(list_arg)), k, fraction, fraction_start, alphabet_id, 1 if multisample else 0, threads)
    if not p:
        raise KdbxError(h.kdbxh_last_error().decode())
    try:
        out = []
        for i in range(h.kdbxh_samples_count(p)):
            n = C.c_uint64()
            ptr = h.kdbxh_samples_kmers(p, i, C.byref(n))
            out.append((h.kdbxh_samples_name(p, i).decode(), _np_from(ptr, n.value, np.uint64).copy()))
        return out
    finally:
        h.kdbxh_samples_free(p)


class Context:
    """One GPU context of libkdbx.so (mirror of SimilarityCalculator's lifetime)."""

    def __init__(self, device: int = -1, chunk_ids: int = 0, tile_cols: int = 0, unit_updates: int = 0,
                 sparse_block_cells: int = 0, query_batch_kmers: int = 0, tile_rows: int = 0, scatter_threads: int = 0,
                 flags: int = 0, upload_chunk_bytes: int = 0):
        k, _ = load()
        self._k = k
        cfg = Config(device, flags, chunk_ids, tile_cols, unit_updates, sparse_block_cells, query_batch_kmers, tile_rows, scatter_threads, upload_chunk_bytes)
        p = C.c_void_p()
        rc = k.kdbx_open(C.byref(cfg), C.byref(p))
        if rc != 0:
            raise KdbxError(k.kdbx_last_error(None).decode())
        self._p = p
        self.num_samples = 0
        self.num_patterns = 0
        self._keep = None

    def close(self):
        if getattr(self, "_p", None):
            self._k.kdbx_close(self._p)
            self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != 0:
            raise KdbxError(self._k.kdbx_last_error(self._p).decode())

    def load_patterns(self, trie_or_view, keep=None):
        v = trie_or_view.view() if isinstance(trie_or_view, Trie) else trie_or_view
        self._keep = (trie_or_view, keep)
        self._check(self._k.kdbx_load_patterns(self._p, C.byref(v)))
        self.num_samples, self.num_patterns = int(v.num_samples), int(v.num_patterns)

    def set_sample_window(self, lo, hi):
        """Declare that every sample id of the staged trie lies in [lo, hi) (kdbx_set_sample_window)."""
        self._check(self._k.kdbx_set_sample_window(self._p, lo, hi))

    def all2all_dense(self, out: np.ndarray | None = None):
        n = self.num_samples
        if out is None:
            out = np.empty(tri_cells(n), dtype=np.uint32)
        st = Stats()
        self._check(self._k.kdbx_all2all_dense(self._p, out.ctypes.data if out.size else None, C.byref(st)))
        return out, st

    def all2all_dense_rows(self, row_begin, row_end, out: np.ndarray | None = None):
        cells = tri_cells(row_end) - tri_cells(row_begin)
        if out is None:
            out = np.empty(cells, dtype=np.uint32)
        st = Stats()
        self._check(self._k.kdbx_all2all_dense_rows(self._p, row_begin, row_end, out.ctypes.data if out.size else None, C.byref(st)))
        return out, st

    def all2all_dense_rows_device(self, row_begin, row_end, device_ptr: int):
        st = Stats()
        self._check(self._k.kdbx_all2all_dense_rows_device(self._p, row_begin, row_end, C.c_void_p(device_ptr), C.byref(st)))
        return st

    def all2all_dense_part_device(self, part, num_parts, device_ptr: int):
        st = Stats()
        self._check(self._k.kdbx_all2all_dense_part_device(self._p, part, num_parts, C.c_void_p(device_ptr), C.byref(st)))
        return st

    def csv_dense_rows(self, row_begin, row_end):
        """(text bytes, row offsets) of the dense table's cells for the rows of the last host-output all2all call."""
        rows = row_end - row_begin
        off = np.zeros(rows + 1, np.uint64)
        total = C.c_uint64(0)
        self._check(self._k.kdbx_csv_dense_rows(self._p, row_begin, row_end, None, 0, off.ctypes.data, C.byref(total)))
        text = np.zeros(max(1, total.value), np.uint8)
        self._check(self._k.kdbx_csv_dense_rows(self._p, row_begin, row_end, text.ctypes.data, total.value, off.ctypes.data, C.byref(total)))
        return text[:total.value].tobytes(), off

    def stage_matrix(self, tri: np.ndarray, num_samples: int):
        tri = np.ascontiguousarray(tri, np.uint32)
        self._check(self._k.kdbx_stage_matrix(self._p, tri.ctypes.data if tri.size else None, num_samples))

    def distance_dense_rows(self, metric: str, sample_kmers, row_begin, row_end):
        """(text bytes, row offsets) of the six-decimal measure table for rows of the resident matrix."""
        cnt = np.ascontiguousarray(sample_kmers, np.uint32)
        rows = row_end - row_begin
        off = np.zeros(rows + 1, np.uint64)
        total = C.c_uint64(0)
        m = {"jaccard": 0, "min": 1, "max": 2, "cosine": 3, "num-kmers": 4}[metric]
        self._check(self._k.kdbx_distance_dense_rows(self._p, m, cnt.ctypes.data, row_begin, row_end, None, 0, off.ctypes.data, C.byref(total)))
        text = np.zeros(max(1, total.value), np.uint8)
        self._check(self._k.kdbx_distance_dense_rows(self._p, m, cnt.ctypes.data, row_begin, row_end, text.ctypes.data, total.value, off.ctypes.data, C.byref(total)))
        return text[:total.value].tobytes(), off

    # ---- several GPUs (kdbx.h: kdbx_comm_*) ----
    @staticmethod
    def comm_unique_id() -> bytes:
        k, _ = load()
        buf = C.create_string_buffer(128)
        if k.kdbx_comm_unique_id(buf) != 0:
            raise KdbxError(k.kdbx_last_error(None).decode())
        return buf.raw

    def comm_init_rank(self, nranks: int, rank: int, unique_id: bytes):
        self._check(self._k.kdbx_comm_init_rank(self._p, nranks, rank, C.create_string_buffer(unique_id, 128)))
        self.comm_nranks, self.comm_rank = nranks, rank

    def block_cells(self) -> int:
        """cells of the packed triangle per rank after the reduce-scatter"""
        n = getattr(self, "comm_nranks", 1)
        return (tri_cells(self.num_samples) + n - 1) // n

    def all2all_dense_reduce_scatter_device(self, device_ptr: int):
        st = Stats()
        first, count = C.c_uint64(0), C.c_uint64(0)
        self._check(self._k.kdbx_all2all_dense_reduce_scatter_device(self._p, C.c_void_p(device_ptr), C.byref(first), C.byref(count), C.byref(st)))
        return int(first.value), int(count.value), st

    def all2all_dense_reduce_scatter(self, out: np.ndarray):
        st = Stats()
        first, count = C.c_uint64(0), C.c_uint64(0)
        self._check(self._k.kdbx_all2all_dense_reduce_scatter(self._p, out.ctypes.data if out.size else None, C.byref(first), C.byref(count), C.byref(st)))
        return int(first.value), int(count.value), st

    def all2all_sparse(self, min_common=0, max_common=0xFFFFFFFF, metric_bounds=(), sample_kmers=None, rows=None):
        """Sparse rows (row_ptr, col, val) as numpy copies; metric_bounds = [(name, lo, hi)]; rows = (begin, end) computes
        only that block of rows (kdbx_all2all_sparse_rows)."""
        f = Filter()
        f.min_common, f.max_common = min_common, max_common
        f.num_metric_bounds = len(metric_bounds)
        for i, (name, lo, hi) in enumerate(metric_bounds):
            f.metric_bounds[i].metric, f.metric_bounds[i].lo, f.metric_bounds[i].hi = METRICS[name], lo, hi
        keep = None
        if sample_kmers is not None:
            keep = np.ascontiguousarray(sample_kmers, np.uint32)
            f.sample_kmers = keep.ctypes.data
        csr, st = Csr(), Stats()
        if rows is None:
            self._check(self._k.kdbx_all2all_sparse(self._p, C.byref(f), C.byref(csr), C.byref(st)))
        else:
            self._check(self._k.kdbx_all2all_sparse_rows(self._p, rows[0], rows[1], C.byref(f), C.byref(csr), C.byref(st)))
        try:
            row_ptr = _np_from(csr.row_ptr, csr.num_rows + 1, np.uint64).copy()
            col = _np_from(csr.col, csr.nnz, np.uint32).copy()
            val = _np_from(csr.val, csr.nnz, np.uint32).copy()
        finally:
            self._k.kdbx_free_csr(C.byref(csr))
        return row_ptr, col, val, st

    def db2db_sparse(self, cols_db: "Context", min_common=0, max_common=0xFFFFFFFF, metric_bounds=(), row_kmers=None, col_kmers=None):
        """One cell of the all2all-parts grid (kdbx_db2db_sparse): this context holds the ROW database, cols_db the COLUMN
        database (patterns and k-mer tables staged on both).  Returns (row_ptr, col, val, stats); column ids are local to
        the column database."""
        f = Filter()
        f.min_common, f.max_common = min_common, max_common
        f.num_metric_bounds = len(metric_bounds)
        for i, (name, lo, hi) in enumerate(metric_bounds):
            f.metric_bounds[i].metric, f.metric_bounds[i].lo, f.metric_bounds[i].hi = METRICS[name], lo, hi
        keep_r = keep_c = None
        if row_kmers is not None:
            keep_r = np.ascontiguousarray(row_kmers, np.uint32)
            f.sample_kmers = keep_r.ctypes.data
        if col_kmers is not None:
            keep_c = np.ascontiguousarray(col_kmers, np.uint32)
        csr, st = Csr(), Stats()
        self._check(self._k.kdbx_db2db_sparse(self._p, cols_db._p, C.byref(f), keep_c.ctypes.data if keep_c is not None else None,
                                              C.byref(csr), C.byref(st)))
        try:
            row_ptr = _np_from(csr.row_ptr, csr.num_rows + 1, np.uint64).copy()
            col = _np_from(csr.col, csr.nnz, np.uint32).copy()
            val = _np_from(csr.val, csr.nnz, np.uint32).copy()
        finally:
            self._k.kdbx_free_csr(C.byref(csr))
        return row_ptr, col, val, st

    def load_hashtables(self, trie_or_view):
        v = trie_or_view.tables_view() if isinstance(trie_or_view, Trie) else trie_or_view
        self._keep_tables = trie_or_view
        self._check(self._k.kdbx_load_hashtables(self._p, C.byref(v)))

    def new2all_batch(self, queries):
        """queries: list of sorted unique uint64 arrays -> (len(queries) x N uint32 matrix, stats)."""
        q_off = np.zeros(len(queries) + 1, np.uint64)
        for i, q in enumerate(queries):
            q_off[i + 1] = q_off[i] + len(q)
        kmers = np.ascontiguousarray(np.concatenate([np.asarray(q, np.uint64) for q in queries]) if queries else np.zeros(0, np.uint64))
        out = np.zeros((len(queries), self.num_samples), np.uint32)
        st = Stats()
        self._check(self._k.kdbx_new2all_batch(self._p, kmers.ctypes.data if kmers.size else None, q_off.ctypes.data, len(queries),
                                              out.ctypes.data if out.size else None, C.byref(st)))
        return out, st

    def new2all_sequences(self, sequences, k=18, fraction=1.0, fraction_start=0.0, alphabet=None, preserve_strand=False):
        """sequences: list of bytes (records separated by a byte outside the alphabet) ->
        (len(sequences) x N uint32 matrix, distinct k-mers per query, stats)."""
        bp = build_params(k, fraction, fraction_start, alphabet, preserve_strand)
        q_off = np.zeros(len(sequences) + 1, np.uint64)
        for i, q in enumerate(sequences):
            q_off[i + 1] = q_off[i] + len(q)
        blob = b"".join(sequences)
        out = np.zeros((len(sequences), self.num_samples), np.uint32)
        uniq = np.zeros(len(sequences), np.uint64)
        st = Stats()
        self._check(self._k.kdbx_new2all_sequences(self._p, C.byref(bp), blob if blob else None, q_off.ctypes.data, len(sequences),
                                                   out.ctypes.data if out.size else None, uniq.ctypes.data if uniq.size else None, C.byref(st)))
        return out, uniq, st

    def row_updates(self):
        out = np.zeros(self.num_samples, dtype=np.uint64)
        self._check(self._k.kdbx_row_updates(self._p, out.ctypes.data if out.size else None))
        return out

    def debug_fetch(self, what: int, count: int, dtype):
        out = np.zeros(count, dtype=dtype)
        got = self._k.kdbx_debug_fetch(self._p, what, out.ctypes.data, count)
        if got < 0:
            raise KdbxError(self._k.kdbx_last_error(self._p).decode())
        return out[:got]


NT_MAP = {"A": 0, "C": 1, "G": 2, "T": 3, "U": 3}


def build_params(k=18, fraction=1.0, fraction_start=0.0, alphabet=None, preserve_strand=False, table_capacity_hint=0):
    """kdbx_build_params for an alphabet given as dict symbol -> code (upper case; lower case is added)."""
    alphabet = alphabet or NT_MAP
    size = max(alphabet.values()) + 1
    bits = max(1, (size - 1).bit_length())
    bp = BuildParams(k, bits, size, 1 if preserve_strand else 0, fraction, fraction_start)
    for i in range(256):
        bp.symbol_map[i] = -1
    for ch, code in alphabet.items():
        bp.symbol_map[ord(ch.upper())] = code
        bp.symbol_map[ord(ch.lower())] = code
    bp.table_capacity_hint = table_capacity_hint
    return bp


class DeviceBuilder:
    """kdbx_builder_* on a Context (the device-side `build`).  alphabet: dict symbol -> code (upper case;
    lower case is added), default nucleotides."""

    def __init__(self, ctx: "Context", k=18, fraction=1.0, fraction_start=0.0, alphabet=None, preserve_strand=False,
                 table_capacity_hint=0):
        self._ctx = ctx
        self._k = ctx._k
        bp = build_params(k, fraction, fraction_start, alphabet, preserve_strand, table_capacity_hint)
        bits = bp.bits_per_symbol
        p = C.c_void_p()
        ctx._check(self._k.kdbx_builder_open(ctx._p, C.byref(bp), C.byref(p)))
        self._p = p
        self.k, self.bits = k, bits

    def close(self):
        if getattr(self, "_p", None):
            self._k.kdbx_builder_close(self._p)
            self._p = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def adopt(self):
        self._ctx._check(self._k.kdbx_builder_adopt(self._p))

    def add_sequence(self, symbols: bytes) -> int:
        n = C.c_uint64()
        self._ctx._check(self._k.kdbx_builder_add_sequence(self._p, symbols, len(symbols), C.byref(n)))
        return int(n.value)

    def add_kmers(self, kmers):
        kmers = np.ascontiguousarray(kmers, np.uint64)
        self._ctx._check(self._k.kdbx_builder_add_kmers(self._p, kmers.ctypes.data if kmers.size else None, kmers.size))

    def finish(self):
        """-> (dict of SoA arrays like Trie.arrays(), slot_off, slots, table_filled, BuildResult)"""
        r = BuildResult()
        self._ctx._check(self._k.kdbx_builder_finish(self._p, C.byref(r)))
        P, T = int(r.num_patterns), int(r.num_tables)
        a = {"num_kmers": np.zeros(P, np.int64), "parent_id": np.zeros(P, np.int64), "n": np.zeros(P, np.uint32),
             "l": np.zeros(P, np.uint32), "last": np.zeros(P, np.uint32), "bits": np.zeros(P, np.uint32),
             "payload_off": np.zeros(P, np.uint64), "payload": np.zeros(int(r.payload_words) + 2, np.uint64)}
        slot_off = np.zeros(T + 1, np.uint64)
        slots = np.zeros(int(r.total_slots), np.uint64)
        filled = np.zeros(T, np.uint64)
        ba = BuildArrays(a["num_kmers"].ctypes.data, a["parent_id"].ctypes.data, a["n"].ctypes.data, a["l"].ctypes.data,
                         a["last"].ctypes.data, a["bits"].ctypes.data, a["payload_off"].ctypes.data, a["payload"].ctypes.data,
                         slot_off.ctypes.data, slots.ctypes.data if slots.size else None, filled.ctypes.data)
        self._ctx._check(self._k.kdbx_builder_export(self._p, C.byref(ba)))
        return a, slot_off, slots, filled, r


def pinned_empty(count: int, dtype=np.uint32):
    """numpy array over page-locked host memory from kdbx_host_alloc (freed with the array)."""
    k, _ = load()
    nbytes = max(1, count * np.dtype(dtype).itemsize)
    p = C.c_void_p()
    if k.kdbx_host_alloc(C.byref(p), nbytes) != 0:
        raise KdbxError("kdbx_host_alloc failed")

    class _Owner:
        def __init__(self, ptr):
            self.ptr = ptr

        def __del__(self):
            k.kdbx_host_free(self.ptr)

    owner = _Owner(p)
    buf = (C.c_char * nbytes).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=count)
    _PINNED_OWNERS[id(buf)] = owner  # keep alive as long as the module; arrays are few and large
    return arr


_PINNED_OWNERS: dict = {}


def shard_rows_by_work(row_updates: np.ndarray, world: int):
    """Contiguous row blocks with (nearly) equal numbers of updates: returns
    world+1 boundaries.  Deterministic, so every rank computes the same split."""
    n = len(row_updates)
    cum = np.concatenate([[0], np.cumsum(row_updates.astype(np.float64))])
    total = cum[-1]
    bounds = [0]
    for r in range(1, world):
        target = total * r / world
        b = int(np.searchsorted(cum, target, side="left"))
        bounds.append(min(max(b, bounds[-1]), n))
    bounds.append(n)
    return bounds
