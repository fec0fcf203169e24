"""Test-side helpers: loading the oracle (oracle/kdb_oracle.c) and making random valid tries.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use the oracle."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"


def load_oracle():
    so = ORACLE_DIR / "_build" / "liboracle.so"
    src = ORACLE_DIR / "kdb_oracle.c"
    if not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(ORACLE_DIR), "_build/liboracle.so", "_build/kdb_oracle"], check=True,
                       stdout=subprocess.DEVNULL)
    lib = C.CDLL(str(so))
    vp = C.c_void_p
    lib.oracle_all2all.argtypes = [C.c_uint64, C.c_uint32] + [vp] * 9
    lib.oracle_all2all.restype = C.c_uint64
    lib.oracle_all2all_bruteforce.argtypes = [C.c_uint64, C.c_uint32] + [vp] * 8
    lib.oracle_all2all_bruteforce.restype = C.c_int
    lib.oracle_all2all_regrouped.argtypes = [C.c_uint64, C.c_uint32] + [vp] * 8
    lib.oracle_all2all_regrouped.restype = C.c_uint64
    lib.oracle_all2all_boundary.argtypes = [C.c_uint64, C.c_uint32] + [vp] * 9
    lib.oracle_all2all_boundary.restype = C.c_uint64
    lib.oracle_all2all_file.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
    lib.oracle_all2all_file.restype = C.c_uint64
    lib.oracle_decode_local.argtypes = [vp, C.c_uint32, C.c_uint32, vp]
    lib.oracle_decode_local.restype = None
    return lib


def tri_cells(n):
    return n * (n - 1) // 2 if n > 0 else 0


def oracle_all2all(lib, N, a):
    """a: dict of SoA numpy arrays (num_kmers,parent_id,n,l,last,bits,payload_off,payload) -> (tri, U)"""
    tri = np.zeros(max(1, tri_cells(N)), dtype=np.uint32)
    pay = a["payload"] if a["payload"].size else np.zeros(2, np.uint64)
    U = lib.oracle_all2all(len(a["n"]), N, a["num_kmers"].ctypes.data, a["parent_id"].ctypes.data, a["n"].ctypes.data,
                           a["l"].ctypes.data, a["last"].ctypes.data, a["bits"].ctypes.data, a["payload_off"].ctypes.data,
                           pay.ctypes.data, tri.ctypes.data)
    assert U != 2**64 - 1
    return tri[:tri_cells(N)], U


def oracle_regrouped(lib, N, a):
    """(tri, operations) by the column-side regrouping of DESIGN.md §8."""
    tri = np.zeros(max(1, tri_cells(N)), dtype=np.uint32)
    pay = a["payload"] if a["payload"].size else np.zeros(2, np.uint64)
    ops = lib.oracle_all2all_regrouped(len(a["n"]), N, a["num_kmers"].ctypes.data, a["parent_id"].ctypes.data, a["n"].ctypes.data,
                                       a["l"].ctypes.data, a["last"].ctypes.data, a["payload_off"].ctypes.data, pay.ctypes.data,
                                       tri.ctypes.data)
    assert ops != 2**64 - 1
    return tri[:tri_cells(N)], ops


def oracle_boundary(lib, N, a):
    """(tri, U, difference updates) by the run-boundary form of kmer-db_b200/csrc/diff.cuh."""
    tri = np.zeros(max(1, tri_cells(N)), dtype=np.uint32)
    pay = a["payload"] if a["payload"].size else np.zeros(2, np.uint64)
    U = C.c_uint64(0)
    phys = lib.oracle_all2all_boundary(len(a["n"]), N, a["num_kmers"].ctypes.data, a["parent_id"].ctypes.data, a["n"].ctypes.data,
                                       a["l"].ctypes.data, a["last"].ctypes.data, a["payload_off"].ctypes.data, pay.ctypes.data,
                                       tri.ctypes.data, C.addressof(U))
    assert phys != 2**64 - 1
    return tri[:tri_cells(N)], U.value, phys


def oracle_bruteforce(lib, N, a):
    tri = np.zeros(max(1, tri_cells(N)), dtype=np.uint32)
    pay = a["payload"] if a["payload"].size else np.zeros(2, np.uint64)
    rc = lib.oracle_all2all_bruteforce(len(a["n"]), N, a["num_kmers"].ctypes.data, a["parent_id"].ctypes.data, a["n"].ctypes.data,
                                       a["l"].ctypes.data, a["last"].ctypes.data, a["payload_off"].ctypes.data, pay.ctypes.data,
                                       tri.ctypes.data)
    assert rc == 0
    return tri[:tri_cells(N)]


def host_w(a):
    """W_p = subtree sums of num_kmers mod 2^32 (reverse sweep, parent_id < id)."""
    W = a["num_kmers"].astype(np.int64).copy()
    par = a["parent_id"]
    for i in range(len(W) - 1, 0, -1):
        if par[i] >= 0:
            W[par[i]] += W[i]
    return (W & 0xFFFFFFFF).astype(np.uint32)


# ---- Elias-gamma writer (format: SURVEY.md §A.1) -------------------------------------------
def gamma_encode(deltas):
    bits = []
    for v in deltas:
        assert v >= 1
        b = int(v).bit_length()
        bits.extend([1] * (b - 1))
        bits.append(0)
        bits.extend((int(v) >> s) & 1 for s in range(b - 2, -1, -1))
    nbits = len(bits)
    nwords = ((nbits + 127) // 128) * 2 if nbits else 0
    words = np.zeros(nwords, dtype=np.uint64)
    for i, bit in enumerate(bits):
        if bit:
            words[i >> 6] |= np.uint64(1) << np.uint64(63 - (i & 63))
    return words, nbits


def random_trie(rng, N, P, max_local=6, big_weights=False, dense_lists=False):
    """A random VALID kmer-db trie with P patterns over N samples (pattern 0 = sentinel).
    Every node picks a parent among earlier nodes (or none) and a strictly ascending local list
    whose first id exceeds the parent's last id."""
    num_kmers = [0]; parent = [-1]; n = [0]; l = [0]; last = [0]; bits = [0]; off = [0]
    payload = []
    words_total = 0
    lists = [[]]
    tries = 0
    while len(n) < P and tries < 50 * P:
        tries += 1
        par = int(rng.integers(-1, len(n)))
        if par == 0:
            par = -1
        lo = 0 if par < 0 else last[par] + 1
        if lo >= N:
            continue
        room = N - lo
        cnt = int(min(room, rng.integers(1, max_local + 1)))
        if dense_lists:
            start = lo + int(rng.integers(0, room - cnt + 1))
            ids = list(range(start, start + cnt))
        else:
            ids = sorted(rng.choice(np.arange(lo, N), size=cnt, replace=False).tolist())
        deltas = [ids[i + 1] - ids[i] for i in range(cnt - 1)]
        w, nb = gamma_encode(deltas)
        if big_weights:
            k = int(rng.choice([0, 1, 7, 2**31, 2**32 - 1, 2**32 + 5, 2**33 + 11, int(rng.integers(0, 2**34))]))
        else:
            k = int(rng.choice([0, 0, 1, 2, 5, int(rng.integers(0, 1000))]))
        num_kmers.append(k); parent.append(par); l.append(cnt); n.append((n[par] if par >= 0 else 0) + cnt)
        last.append(ids[-1]); bits.append(nb); off.append(words_total)
        payload.append(w); words_total += len(w)
        lists.append(ids)
    a = {
        "num_kmers": np.array(num_kmers, np.int64), "parent_id": np.array(parent, np.int64),
        "n": np.array(n, np.uint32), "l": np.array(l, np.uint32), "last": np.array(last, np.uint32),
        "bits": np.array(bits, np.uint32), "payload_off": np.array(off, np.uint64),
        "payload": np.concatenate(payload + [np.zeros(2, np.uint64)]) if payload else np.zeros(2, np.uint64),
    }
    return a, lists


def read_bytes(p):
    with open(p, "rb") as f:
        return f.read()
