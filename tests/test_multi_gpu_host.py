"""Host-side logic of the N>1 paths on CPU: two `gloo` ranks shard the work the way bench.py does
(sub-tries from kdbxh_partition / relabelled shards + one all-reduce of uint32 partial matrices, as
bench.py does; also pattern chunks round-robin + all-reduce and row blocks balanced on per-row updates + gather), with the ORACLE standing in for the device kernels."""
import os
import socket
import sys

import numpy as np
import pytest

import oracle_util as ou

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, str(ou.ROOT / "kmer-db_b200"))
    sys.path.insert(0, str(ou.ROOT / "tests"))
    import ctypes as C
    import kdbx
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        oracle = ou.load_oracle()
        rng = np.random.default_rng(123)  # same trie on every rank (the trie is replicated)
        N = 150
        a, lists = ou.random_trie(rng, N, 2500, max_local=12, big_weights=True)
        full, U = ou.oracle_all2all(oracle, N, a)
        # (1) pattern chunks round-robin + all-reduce; uint32 sums wrap like int32 sums
        part = np.zeros(ou.tri_cells(N), np.uint32)
        oracle.oracle_all2all_part.restype = C.c_uint64
        u = oracle.oracle_all2all_part(C.c_uint64(len(a["n"])), C.c_uint32(N), a["num_kmers"].ctypes.data_as(C.c_void_p),
                                       a["parent_id"].ctypes.data_as(C.c_void_p), a["n"].ctypes.data_as(C.c_void_p),
                                       a["l"].ctypes.data_as(C.c_void_p), a["last"].ctypes.data_as(C.c_void_p),
                                       a["payload_off"].ctypes.data_as(C.c_void_p), a["payload"].ctypes.data_as(C.c_void_p),
                                       C.c_uint64(64), C.c_uint32(rank), C.c_uint32(world), part.ctypes.data_as(C.c_void_p))
        t = torch.from_numpy(part.view(np.int32).copy())
        dist.all_reduce(t)
        ut = torch.tensor([u], dtype=torch.int64)
        dist.all_reduce(ut)
        ok1 = bool(np.array_equal(t.numpy().view(np.uint32), full)) and int(ut.item()) == U
        # (2) row blocks balanced on per-row updates, gathered in rank order
        upd = np.zeros(N, np.uint64)
        nn, ll = a["n"].astype(np.int64), a["l"].astype(np.int64)
        for p, ids in enumerate(lists):
            first = int(nn[p] - ll[p])
            for j, row in enumerate(ids):
                upd[row] += first + j
        assert int(upd.sum()) == U
        b = kdbx.shard_rows_by_work(upd, world)
        mine = full[ou.tri_cells(b[rank]):ou.tri_cells(b[rank + 1])]
        sizes = [ou.tri_cells(b[r + 1]) - ou.tri_cells(b[r]) for r in range(world)]
        bufs = [torch.zeros(sz, dtype=torch.int32) for sz in sizes]
        # gloo all_gather needs equal sizes: pad to the largest share
        m = max(sizes)
        padded = torch.zeros(m, dtype=torch.int32)
        padded[:mine.size] = torch.from_numpy(mine.view(np.int32).copy())
        out = [torch.zeros(m, dtype=torch.int32) for _ in range(world)]
        dist.all_gather(out, padded)
        got = np.concatenate([out[r][:sizes[r]].numpy().view(np.uint32) for r in range(world)])
        ok2 = bool(np.array_equal(got, full)) and b[0] == 0 and b[-1] == N and all(b[i] <= b[i + 1] for i in range(world))
        # shares are balanced: no rank has more than ~1.5x the mean work on this input
        work = [int(upd[b[r]:b[r + 1]].sum()) for r in range(world)]
        ok3 = max(work) <= 1.5 * (U / world) + int(upd.max())
        # max over ranks of a per-rank time, as bench.py reports it
        tt = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ok4 = tt.item() == float(world)
        # (3) what bench.py does: the trie is cut into one sub-trie per rank (kdbxh_partition), each rank runs the
        # whole pipeline on its own part (here: the oracle) and one all-reduce adds the partial matrices
        t = kdbx.Trie.synth(num_samples=64, num_clusters=4, genome_kmers=8000, seed=11)
        whole, U3 = ou.oracle_all2all(oracle, 64, t.arrays())
        sub, owned = t.partition(world, rank)
        mine3, u3 = ou.oracle_all2all(oracle, 64, sub.arrays())
        t3 = torch.from_numpy(mine3.view(np.int32).copy())
        dist.all_reduce(t3)
        ut3 = torch.tensor([owned, u3], dtype=torch.int64)
        dist.all_reduce(ut3)
        ok5 = bool(np.array_equal(t3.numpy().view(np.uint32), whole)) and int(ut3[0]) == U3 and U3 <= int(ut3[1]) <= 1.05 * U3
        # (4) weak scaling: every rank lays its own copy of the database onto its own sample ids (kdbxh_relabel)
        t.relabel(rank * 64, world * 64)
        mine4, u4 = ou.oracle_all2all(oracle, world * 64, t.arrays())
        t4 = torch.from_numpy(mine4.view(np.int32).copy())
        dist.all_reduce(t4)
        got4 = t4.numpy().view(np.uint32)
        ok6 = u4 == U3
        for r in range(world):  # block r of the big matrix is the small matrix; everything else is zero
            for srow in range(1, 64):
                o = ou.tri_cells(srow + r * 64)
                ok6 = ok6 and bool(np.array_equal(got4[o + r * 64:o + r * 64 + srow], whole[ou.tri_cells(srow):ou.tri_cells(srow) + srow]))
                ok6 = ok6 and not got4[o:o + r * 64].any()
        # (5) the product's N > 1 path (bench.py, CLI -gpus): one cut for all ranks (kdbxh_partitioner), every rank knows the
        # band of sample ids its part covers, a REDUCE-SCATTER of the partial matrices leaves rank r with the cells
        # [r B, (r+1) B) of the packed triangle, B = ceil(cells / world) (kdbx_all2all_dense_reduce_scatter)
        t5 = kdbx.Trie.synth(num_samples=96, num_clusters=5, genome_kmers=6000, seed=13, cluster_skew=0.4)
        whole5, U5 = ou.oracle_all2all(oracle, 96, t5.arrays())
        parts = list(t5.partition_all(world))
        sub5, owned5, win5 = parts[rank]
        a5 = sub5.arrays()
        ids_ok = True
        for p5 in range(len(a5["l"])):   # every decoded id of the part lies inside its declared window
            if a5["l"][p5]:
                out = np.zeros(int(a5["l"][p5]), np.uint32)
                oracle.oracle_decode_local(a5["payload"][int(a5["payload_off"][p5]):].ctypes.data, int(a5["l"][p5]), int(a5["last"][p5]), out.ctypes.data)
                ids_ok = ids_ok and win5[0] <= int(out[0]) and int(out[-1]) < win5[1]
        mine5, _ = ou.oracle_all2all(oracle, 96, a5)
        cells5 = ou.tri_cells(96)
        B = (cells5 + world - 1) // world
        padded5 = np.zeros(B * world, np.uint32)
        padded5[:cells5] = mine5
        t5r = torch.from_numpy(padded5.view(np.int32).copy())   # gloo has no reduce_scatter: all_reduce, keep the own block
        dist.all_reduce(t5r)
        outs = t5r[rank * B:(rank + 1) * B]
        lo5, hi5 = min(cells5, rank * B), min(cells5, (rank + 1) * B)
        ok7 = ids_ok and bool(np.array_equal(outs.numpy().view(np.uint32)[:hi5 - lo5], whole5[lo5:hi5]))
        ot = torch.tensor([owned5], dtype=torch.int64)
        dist.all_reduce(ot)
        ok7 = ok7 and int(ot.item()) == U5
        q.put((rank, ok1, ok2, ok3, ok4, ok5, ok6, ok7))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_two_rank_sharding_over_gloo(libs, oracle, world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok1, ok2, ok3, ok4, ok5, ok6, ok7 in sorted(res):
        assert ok1, f"rank {rank}: all-reduced partial matrices differ from the full matrix"
        assert ok2, f"rank {rank}: gathered row blocks differ from the full matrix"
        assert ok3, f"rank {rank}: row shares are unbalanced"
        assert ok4
        assert ok5, f"rank {rank}: all-reduced matrices of the sub-tries differ from the matrix of the whole trie"
        assert ok6, f"rank {rank}: relabelled shards do not form the block-diagonal matrix"
        assert ok7, f"rank {rank}: reduce-scattered block of the partitioner's parts differs from the whole matrix (or an id left its window)"
